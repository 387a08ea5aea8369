// hbfill.cu -- libhbfill.so: C ABI (include/hbfill.h) of the B200-native histbook fill engine.
//
// Host side: NVRTC compile + module load of the generated per-book kernel, float64 arenas for the bin
// contents, the launch loop (device-resident columns in place, host columns through double-buffered
// staging), and a few hand-written utility kernels (arena algebra, L2 flush).  The per-event work is in
// the generated kernel (histbook_b200/codegen.py + csrc/hb_template.cuh).
//
// Built by __graft_entry__.build():
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC hbfill.cu -ldl
// Neither libcuda nor libnvrtc is linked: both are resolved with dlopen (driver at hb_init, NVRTC at the first
// compile), so the library loads -- and can NVRTC-compile -- on a machine without a GPU.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvrtc.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/hbfill.h"
#include "hb_template.cuh"

// NVTX ranges around the host-visible stages (compile, load, staging copies, launches, reads): visible in Nsight
// Systems / any NVTX consumer; a few nanoseconds each without a tool attached.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// ------------------------------------------------------------------------------------------------
// errors

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CUDA_TRY(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(HB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
    } while (0)

extern "C" const char* hb_last_error(void) { return g_err.c_str(); }
extern "C" int hb_version(void) { return 1; }
extern "C" void hb_free(void* p) { free(p); }

// ------------------------------------------------------------------------------------------------
// driver API through dlopen

struct Driver {
    void* lib = nullptr;
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
};

static Driver g_drv;
static std::mutex g_mutex;
static int g_device = -1;
static int g_sm_count = 0;
static int g_max_smem_optin = 0;
static cudaStream_t g_stream = 0;          // compute stream (default: legacy stream 0, ordered with torch's default)
static cudaStream_t g_copy_stream = 0;     // non-blocking stream for host -> device staging
static int64_t g_chunk_events = (int64_t)1 << 24;

static std::string drv_err(CUresult r) {
    const char* s = nullptr;
    if (g_drv.GetErrorString && g_drv.GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
    return "CUresult " + std::to_string((int)r);
}

#define DRV_TRY(call)                                                                          \
    do {                                                                                       \
        CUresult r_ = (call);                                                                  \
        if (r_ != CUDA_SUCCESS) return fail(HB_ERR_CUDA, std::string(#call) + ": " + drv_err(r_)); \
    } while (0)

static int load_driver() {
    if (g_drv.lib) return HB_OK;
    void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(HB_ERR_CUDA, "libcuda.so.1 not found: no NVIDIA driver on this machine (there is no CPU fallback)");
#define SYM(field, name)                                                                       \
    *(void**)(&g_drv.field) = dlsym(lib, name);                                                \
    if (!g_drv.field) return fail(HB_ERR_CUDA, std::string("missing driver symbol ") + name)
    SYM(ModuleLoadData, "cuModuleLoadData");
    SYM(ModuleUnload, "cuModuleUnload");
    SYM(ModuleGetFunction, "cuModuleGetFunction");
    SYM(FuncSetAttribute, "cuFuncSetAttribute");
    SYM(FuncGetAttribute, "cuFuncGetAttribute");
    SYM(LaunchKernel, "cuLaunchKernel");
    SYM(OccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor");
    SYM(GetErrorString, "cuGetErrorString");
#undef SYM
    g_drv.lib = lib;
    return HB_OK;
}

static int ensure_device() {
    if (g_device < 0) return fail(HB_ERR_CUDA, "hb_init has not been called");
    CUDA_TRY(cudaSetDevice(g_device));
    return HB_OK;
}

extern "C" int hb_init(int device) {
    std::lock_guard<std::mutex> lock(g_mutex);
    int rc = load_driver();
    if (rc) return rc;
    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(HB_ERR_ARG, "no such CUDA device");
    if (g_device >= 0 && g_device != device)
        return fail(HB_ERR_ARG, "the engine is bound to cuda:" + std::to_string(g_device) + " (one process per GPU: streams, staging and arenas live there)");
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaFree(0));                        // make the primary context current
    if (g_device != device) {
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        g_sm_count = prop.multiProcessorCount;
        g_max_smem_optin = (int)prop.sharedMemPerBlockOptin;
        if (g_copy_stream) { cudaStreamDestroy(g_copy_stream); g_copy_stream = 0; }
        CUDA_TRY(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
        g_device = device;
    }
    return HB_OK;
}

extern "C" int hb_device_info(int* sm_count, int* max_smem_optin, int64_t* free_bytes, int64_t* total_bytes) {
    int rc = ensure_device();
    if (rc) return rc;
    size_t f = 0, t = 0;
    CUDA_TRY(cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = g_sm_count;
    if (max_smem_optin) *max_smem_optin = g_max_smem_optin;
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    return HB_OK;
}

extern "C" int hb_set_stream(void* s) { g_stream = (cudaStream_t)s; return HB_OK; }

// Order two streams: everything queued on `signaler` so far completes before anything queued on `waiter` from now on.
// nullptr stands for the engine's compute stream.  Used around a fill whose device columns were produced on a stream
// other than the engine's (a non-blocking side stream is not ordered with stream 0 by CUDA itself).
extern "C" int hb_stream_wait(void* waiter, void* signaler) {
    int rc = ensure_device();
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_mutex);
    static cudaEvent_t ev = nullptr;
    if (!ev) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaStream_t w = waiter ? (cudaStream_t)waiter : g_stream, s = signaler ? (cudaStream_t)signaler : g_stream;
    if (w == s) return HB_OK;
    CUDA_TRY(cudaEventRecord(ev, s));
    CUDA_TRY(cudaStreamWaitEvent(w, ev, 0));
    return HB_OK;
}

extern "C" int hb_sync(void) {
    int rc = ensure_device();
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_copy_stream));
    return HB_OK;
}

extern "C" int hb_set_chunk_events(int64_t events) {
    if (events < 4096) return fail(HB_ERR_ARG, "chunk too small");
    g_chunk_events = events;
    return HB_OK;
}

// ------------------------------------------------------------------------------------------------
// programs

struct hb_program {
    CUmodule module = nullptr;
    CUfunction fn = nullptr;
    int threads = 256;
    int tile = 1024;            // events per block and round of the kernel's grid-stride loop (HB_TILE of the generated source)
    int max_blocks_per_sm = 0;  // > 0: the grid is sized for at most this many resident blocks per SM (parts of a split book share the SMs)
    int smem = 0;
    int blocks_per_sm = 1;
    int regs = 0;
    int static_smem = 0;
};

// NVRTC is resolved with dlopen by absolute path first: a process that imported torch already holds torch's
// bundled libnvrtc.so.12 (CUDA 12.8), and a plain -lnvrtc would silently bind to that older copy (same soname),
// whose ptxas rejects the 256-bit loads of hb_template.cuh.  Opening the toolkit's file by path gives a
// separate, private copy.
struct Nvrtc {
    void* lib = nullptr;
    nvrtcResult (*Version)(int*, int*) = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    const char* (*GetErrorString)(nvrtcResult) = nullptr;
    int major = 0, minor = 0;
};
static Nvrtc g_nvrtc;

static int load_nvrtc() {
    if (g_nvrtc.lib) return HB_OK;
    std::vector<std::string> names;
    if (const char* env = getenv("HB_NVRTC_PATH")) names.push_back(env);
    names.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
    names.push_back("/usr/local/cuda/targets/x86_64-linux/lib/libnvrtc.so.12");
    names.push_back("libnvrtc.so.12");
    names.push_back("libnvrtc.so");
    void* lib = nullptr;
    std::string tried;
    for (const std::string& n : names) {
        lib = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (lib) break;
        tried += n + " ";
    }
    if (!lib) return fail(HB_ERR_NVRTC, "libnvrtc not found (tried: " + tried + ")");
#define SYM(field, name)                                                                       \
    *(void**)(&g_nvrtc.field) = dlsym(lib, name);                                              \
    if (!g_nvrtc.field) return fail(HB_ERR_NVRTC, std::string("missing NVRTC symbol ") + name)
    SYM(Version, "nvrtcVersion");
    SYM(CreateProgram, "nvrtcCreateProgram");
    SYM(CompileProgram, "nvrtcCompileProgram");
    SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    SYM(GetProgramLog, "nvrtcGetProgramLog");
    SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    SYM(GetCUBIN, "nvrtcGetCUBIN");
    SYM(DestroyProgram, "nvrtcDestroyProgram");
    SYM(GetErrorString, "nvrtcGetErrorString");
#undef SYM
    g_nvrtc.Version(&g_nvrtc.major, &g_nvrtc.minor);
    g_nvrtc.lib = lib;
    return HB_OK;
}

extern "C" int hb_nvrtc_version(void) {
    if (load_nvrtc()) return -1;
    return g_nvrtc.major * 1000 + g_nvrtc.minor * 10;
}

extern "C" int hb_program_compile(const char* source, const char* header, int lineinfo,
                                  void** cubin, size_t* cubin_size) {
    if (!source || !header || !cubin || !cubin_size) return fail(HB_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lock(g_mutex);
    NvtxRange range("hb_program_compile (NVRTC)");
    int rc = load_nvrtc();
    if (rc) return rc;
    nvrtcProgram prog;
    const char* headers[1] = {header};
    const char* names[1] = {"hb_template.cuh"};
    nvrtcResult r = g_nvrtc.CreateProgram(&prog, source, "hb_fill_generated.cu", 1, headers, names);
    if (r != NVRTC_SUCCESS) return fail(HB_ERR_NVRTC, std::string("nvrtcCreateProgram: ") + g_nvrtc.GetErrorString(r));
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "--fmad=false", "--std=c++17",
                                     "--prec-div=true", "--prec-sqrt=true", "--ftz=false"};
    // LDG.256 (ld.global.v4.u64) needs the CUDA >= 12.9 ptxas
    if (g_nvrtc.major > 12 || (g_nvrtc.major == 12 && g_nvrtc.minor >= 9)) opts.push_back("-DHB_LDG256=1");
    if (lineinfo) opts.push_back("-lineinfo");
    r = g_nvrtc.CompileProgram(prog, (int)opts.size(), opts.data());
    if (r != NVRTC_SUCCESS) {
        size_t n = 0;
        g_nvrtc.GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) g_nvrtc.GetProgramLog(prog, &log[0]);
        g_nvrtc.DestroyProgram(&prog);
        return fail(HB_ERR_NVRTC, std::string("NVRTC: ") + g_nvrtc.GetErrorString(r) + "\n" + log);
    }
    size_t n = 0;
    r = g_nvrtc.GetCUBINSize(prog, &n);
    if (r != NVRTC_SUCCESS || n == 0) {
        g_nvrtc.DestroyProgram(&prog);
        return fail(HB_ERR_NVRTC, "nvrtcGetCUBINSize failed");
    }
    void* out = malloc(n);
    g_nvrtc.GetCUBIN(prog, (char*)out);
    g_nvrtc.DestroyProgram(&prog);
    *cubin = out;
    *cubin_size = n;
    return HB_OK;
}

extern "C" int hb_program_load(const void* cubin, size_t cubin_size, const char* kernel_name,
                               int threads, int smem_bytes, hb_program** out) {
    (void)cubin_size;
    int rc = ensure_device();
    if (rc) return rc;
    if (!cubin || !kernel_name || !out) return fail(HB_ERR_ARG, "null argument");
    if (smem_bytes > g_max_smem_optin)
        return fail(HB_ERR_ARG, "kernel needs more dynamic shared memory than the device allows");
    hb_program* p = new hb_program();
    p->threads = threads;
    p->tile = threads * 4;
    p->smem = smem_bytes;
    CUresult r = g_drv.ModuleLoadData(&p->module, cubin);
    if (r != CUDA_SUCCESS) { delete p; return fail(HB_ERR_CUDA, "cuModuleLoadData: " + drv_err(r)); }
    r = g_drv.ModuleGetFunction(&p->fn, p->module, kernel_name);
    if (r != CUDA_SUCCESS) { g_drv.ModuleUnload(p->module); delete p; return fail(HB_ERR_CUDA, "cuModuleGetFunction: " + drv_err(r)); }
    if (smem_bytes > 48 * 1024) {
        r = g_drv.FuncSetAttribute(p->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem_bytes);
        if (r != CUDA_SUCCESS) { g_drv.ModuleUnload(p->module); delete p; return fail(HB_ERR_CUDA, "cuFuncSetAttribute(smem): " + drv_err(r)); }
    }
    // the whole unified L1 as shared memory: every kernel here streams its inputs past L1 (ld.global.nc.L1::no_allocate) and
    // lives in shared memory; and the two half-book kernels of a split book only share an SM if the carve-out holds both
    g_drv.FuncSetAttribute(p->fn, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, 100);
    g_drv.FuncGetAttribute(&p->regs, CU_FUNC_ATTRIBUTE_NUM_REGS, p->fn);
    g_drv.FuncGetAttribute(&p->static_smem, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, p->fn);
    int occ = 0;
    r = g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&occ, p->fn, threads, (size_t)smem_bytes);
    if (r != CUDA_SUCCESS || occ < 1) { g_drv.ModuleUnload(p->module); delete p; return fail(HB_ERR_CUDA, "kernel cannot be resident on an SM (occupancy 0)"); }
    p->blocks_per_sm = occ;
    *out = p;
    return HB_OK;
}

extern "C" int hb_program_destroy(hb_program* p) {
    if (!p) return HB_OK;
    if (p->module && g_drv.ModuleUnload) g_drv.ModuleUnload(p->module);
    delete p;
    return HB_OK;
}

extern "C" int hb_program_set_tile(hb_program* p, int tile_events) {
    if (!p || tile_events < 1) return fail(HB_ERR_ARG, "bad tile size");
    p->tile = tile_events;
    return HB_OK;
}

extern "C" int hb_program_set_max_blocks_per_sm(hb_program* p, int blocks) {
    if (!p || blocks < 0) return fail(HB_ERR_ARG, "bad block limit");
    p->max_blocks_per_sm = blocks;
    return HB_OK;
}

extern "C" int hb_program_info(const hb_program* p, int* regs, int* static_smem, int* blocks_per_sm) {
    if (!p) return fail(HB_ERR_ARG, "null program");
    if (regs) *regs = p->regs;
    if (static_smem) *static_smem = p->static_smem;
    if (blocks_per_sm) *blocks_per_sm = p->blocks_per_sm;
    return HB_OK;
}

// ------------------------------------------------------------------------------------------------
// host copies

// a small pool of copy threads (created at the first pageable column)
struct CopyJob { char* dst; const char* src; size_t bytes; };
struct CopyPool {
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable wake, done;
    std::vector<CopyJob> jobs;
    size_t next = 0, finished = 0;
    bool stop = false;

    void run() {
        std::unique_lock<std::mutex> lock(m);
        while (true) {
            wake.wait(lock, [this] { return stop || next < jobs.size(); });
            if (stop) return;
            const CopyJob j = jobs[next++];
            lock.unlock();
            memcpy(j.dst, j.src, j.bytes);
            lock.lock();
            if (++finished == jobs.size()) done.notify_all();
        }
    }
    void start(int n) {
        for (int i = 0; i < n; ++i) workers.emplace_back([this] { run(); });
    }
    // copies all jobs; the calling thread works too
    void copy(std::vector<CopyJob>& batch) {
        if (batch.empty()) return;
        std::unique_lock<std::mutex> lock(m);
        jobs.swap(batch);
        next = 0;
        finished = 0;
        wake.notify_all();
        while (next < jobs.size()) {
            const CopyJob j = jobs[next++];
            lock.unlock();
            memcpy(j.dst, j.src, j.bytes);
            lock.lock();
            ++finished;
        }
        done.wait(lock, [this] { return finished == jobs.size(); });
        jobs.clear();
        next = finished = 0;
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lock(m);
            stop = true;
        }
        wake.notify_all();
        for (auto& t : workers) t.join();
    }
};
static CopyPool* g_pool = nullptr;
static int g_copy_threads = -1;

static CopyPool* copy_pool() {
    if (!g_pool) {
        if (g_copy_threads < 0) {
            const char* env = getenv("HB_COPY_THREADS");
            unsigned hw = std::thread::hardware_concurrency();
            // measured on a 16-core B200 host, c4 from pageable numpy columns: 4 threads 23.9 GB/s, 8: 32.1, 12: 36.3, 16: 35.8
            // (page-locked columns: 44.8 GB/s) -- tools/gpu/e2e_threads.sh
            g_copy_threads = env ? atoi(env) : (int)(hw >= 16 ? 12 : (hw >= 4 ? hw * 3 / 4 : 1));
            // one process per GPU: the ranks of a box share its cores (8 ranks on 32 cores: 3 threads each, not 12)
            const char* lws = getenv("LOCAL_WORLD_SIZE");
            const int ranks = lws ? atoi(lws) : 1;
            if (!env && ranks > 1) {
                const int share = (int)(hw * 3 / 4) / ranks;
                if (g_copy_threads > share) g_copy_threads = share < 2 ? 2 : share;
            }
            if (g_copy_threads < 1) g_copy_threads = 1;
        }
        g_pool = new CopyPool();
        g_pool->start(g_copy_threads - 1);
    }
    return g_pool;
}

extern "C" int hb_set_copy_threads(int n) {
    if (n < 1 || n > 64) return fail(HB_ERR_ARG, "copy threads must be 1..64");
    if (g_pool) { delete g_pool; g_pool = nullptr; }
    g_copy_threads = n;
    return HB_OK;
}

static bool is_pageable(const void* ptr) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return true; }
    return attr.type == cudaMemoryTypeUnregistered;
}

// ------------------------------------------------------------------------------------------------
// arenas

struct hb_arena {
    double* ptr = nullptr;
    int64_t n = 0;
};

__global__ void hb_axpy_kernel(double* __restrict__ dst, const double* __restrict__ src, double alpha, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = __dadd_rn(dst[i], __dmul_rn(alpha, src[i]));      // two roundings like numpy's a + alpha*b, never an FMA
}
__global__ void hb_scale_kernel(double* __restrict__ dst, double alpha, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = dst[i] * alpha;
}
__global__ void hb_flush_kernel(float4* __restrict__ buf, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
// second half of the flush: stream a different buffer through L2 so that what the cache holds afterwards is clean
// (unrelated) data -- after the write pass alone the timed kernel would also pay for the write-back of 126 MB of
// dirty lines, which no real caller's fill does
__global__ void hb_flush_read_kernel(const float4* __restrict__ buf, int64_t n, float* __restrict__ sink) {
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = buf[i];
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456f) *sink = acc;
}

static int grid_for(int64_t n) {
    int64_t g = (n + 255) / 256;
    int64_t cap = (int64_t)(g_sm_count > 0 ? g_sm_count : 148) * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" int hb_arena_create(int64_t ndoubles, hb_arena** out) {
    int rc = ensure_device();
    if (rc) return rc;
    if (ndoubles < 0 || !out) return fail(HB_ERR_ARG, "bad arena size");
    hb_arena* a = new hb_arena();
    a->n = ndoubles;
    size_t bytes = (size_t)(ndoubles > 0 ? ndoubles : 1) * sizeof(double);
    cudaError_t e = cudaMalloc((void**)&a->ptr, bytes);
    if (e != cudaSuccess) { delete a; return fail(HB_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    e = cudaMemsetAsync(a->ptr, 0, bytes, g_stream);
    if (e != cudaSuccess) { cudaFree(a->ptr); delete a; return fail(HB_ERR_CUDA, std::string("cudaMemsetAsync: ") + cudaGetErrorString(e)); }
    *out = a;
    return HB_OK;
}

extern "C" int hb_arena_destroy(hb_arena* a) {
    if (!a) return HB_OK;
    if (a->ptr) cudaFree(a->ptr);
    delete a;
    return HB_OK;
}

extern "C" int hb_arena_clear(hb_arena* a) {
    int rc = ensure_device();
    if (rc) return rc;
    if (!a) return fail(HB_ERR_ARG, "null arena");
    CUDA_TRY(cudaMemsetAsync(a->ptr, 0, (size_t)a->n * sizeof(double), g_stream));
    return HB_OK;
}

extern "C" int hb_arena_clone(const hb_arena* src, hb_arena** out) {
    if (!src) return fail(HB_ERR_ARG, "null arena");
    int rc = hb_arena_create(src->n, out);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync((*out)->ptr, src->ptr, (size_t)src->n * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
    return HB_OK;
}

extern "C" int hb_arena_resize(hb_arena* a, int64_t ndoubles) {
    int rc = ensure_device();
    if (rc) return rc;
    if (!a || ndoubles < 0) return fail(HB_ERR_ARG, "bad arena size");
    if (ndoubles == a->n) return HB_OK;
    double* fresh = nullptr;
    size_t bytes = (size_t)(ndoubles > 0 ? ndoubles : 1) * sizeof(double);
    CUDA_TRY(cudaMalloc((void**)&fresh, bytes));
    CUDA_TRY(cudaMemsetAsync(fresh, 0, bytes, g_stream));
    int64_t keep = ndoubles < a->n ? ndoubles : a->n;
    if (keep > 0) CUDA_TRY(cudaMemcpyAsync(fresh, a->ptr, (size_t)keep * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    cudaFree(a->ptr);
    a->ptr = fresh;
    a->n = ndoubles;
    return HB_OK;
}

// Device -> host copy of bin contents.  Large reads into pageable memory (numpy arrays: every `Hist._content`, pickle and
// JSON of c5's 134 MB) go through two page-locked staging buffers: the DMA of chunk k + 1 runs while the copy threads
// move chunk k into the caller's array (a plain cudaMemcpy into pageable memory is a serial staged copy in the driver).
static void* g_read_pin[2] = {nullptr, nullptr};
static cudaEvent_t g_read_done[2] = {nullptr, nullptr};
static const size_t kReadChunk = (size_t)8 << 20;

extern "C" int hb_arena_read(const hb_arena* a, int64_t offset, int64_t count, double* host_out) {
    int rc = ensure_device();
    if (rc) return rc;
    if (!a || offset < 0 || count < 0 || offset + count > a->n || !host_out) return fail(HB_ERR_ARG, "arena read out of range");
    NvtxRange range("hb_arena_read");
    const size_t bytes = (size_t)count * sizeof(double);
    if (bytes < 2 * kReadChunk || !is_pageable(host_out)) {
        CUDA_TRY(cudaMemcpyAsync(host_out, a->ptr + offset, bytes, cudaMemcpyDeviceToHost, g_stream));
        CUDA_TRY(cudaStreamSynchronize(g_stream));
        return HB_OK;
    }
    std::lock_guard<std::mutex> lock(g_mutex);
    for (int b = 0; b < 2; ++b) {
        if (!g_read_pin[b]) {
            CUDA_TRY(cudaHostAlloc(&g_read_pin[b], kReadChunk, cudaHostAllocDefault));
            CUDA_TRY(cudaEventCreateWithFlags(&g_read_done[b], cudaEventDisableTiming | cudaEventBlockingSync));
        }
    }
    const char* src = (const char*)(a->ptr + offset);
    char* dst = (char*)host_out;
    const size_t nchunks = (bytes + kReadChunk - 1) / kReadChunk;
    std::vector<CopyJob> jobs;
    copy_pool();
    const int np = g_copy_threads > 0 ? g_copy_threads : 1;
    for (size_t k = 0; k <= nchunks; ++k) {
        if (k < nchunks) {
            const size_t o = k * kReadChunk, m = std::min(kReadChunk, bytes - o);
            CUDA_TRY(cudaMemcpyAsync(g_read_pin[k & 1], src + o, m, cudaMemcpyDeviceToHost, g_stream));
            CUDA_TRY(cudaEventRecord(g_read_done[k & 1], g_stream));
        }
        if (k > 0) {
            const size_t o = (k - 1) * kReadChunk, m = std::min(kReadChunk, bytes - o);
            CUDA_TRY(cudaEventSynchronize(g_read_done[(k - 1) & 1]));
            jobs.clear();
            const size_t step = ((m + np - 1) / np + 4095) / 4096 * 4096;
            for (size_t q = 0; q < m; q += step)
                jobs.push_back({dst + o + q, (const char*)g_read_pin[(k - 1) & 1] + q, std::min(step, m - q)});
            copy_pool()->copy(jobs);
        }
    }
    return HB_OK;
}

extern "C" int hb_arena_write(hb_arena* a, int64_t offset, int64_t count, const double* host_in) {
    int rc = ensure_device();
    if (rc) return rc;
    if (!a || offset < 0 || count < 0 || offset + count > a->n || !host_in) return fail(HB_ERR_ARG, "arena write out of range");
    CUDA_TRY(cudaMemcpyAsync(a->ptr + offset, host_in, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return HB_OK;
}

extern "C" int hb_arena_axpy(hb_arena* dst, const hb_arena* src, double alpha) {
    int rc = ensure_device();
    if (rc) return rc;
    if (!dst || !src || dst->n != src->n) return fail(HB_ERR_ARG, "arena sizes differ");
    if (dst->n == 0) return HB_OK;
    hb_axpy_kernel<<<grid_for(dst->n), 256, 0, g_stream>>>(dst->ptr, src->ptr, alpha, dst->n);
    CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

// dst[dst_off .. dst_off + count) += alpha * src[src_off .. src_off + count): one slab slot of a group-axis histogram
// added to another's (Hist.__add__ on dict contents, histbook/hist.py:513-537)
extern "C" int hb_arena_axpy_at(hb_arena* dst, int64_t dst_off, const hb_arena* src, int64_t src_off, int64_t count, double alpha) {
    int rc = ensure_device();
    if (rc) return rc;
    if (!dst || !src || count < 0 || dst_off < 0 || src_off < 0 || dst_off + count > dst->n || src_off + count > src->n)
        return fail(HB_ERR_ARG, "hb_arena_axpy_at out of range");
    if (count == 0) return HB_OK;
    hb_axpy_kernel<<<grid_for(count), 256, 0, g_stream>>>(dst->ptr + dst_off, src->ptr + src_off, alpha, count);
    CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

// Snapshot of many arenas (or slab slots) into one flat buffer with ONE launch: the rank merge copies every histogram of a
// book into the buffer its collective reduces (dist.Merger).  64 small torch copies cost ~0.4 ms of launches per merge.
#define HB_SNAP_MAX 128
struct HbSnapshot {
    const double* src[HB_SNAP_MAX];
    long long dst[HB_SNAP_MAX];
    long long count[HB_SNAP_MAX];
};
__global__ void hb_snapshot_kernel(const HbSnapshot s, double* __restrict__ out) {
    const int seg = blockIdx.y;
    const double* __restrict__ src = s.src[seg];
    double* __restrict__ dst = out + s.dst[seg];
    const long long n = s.count[seg];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}

extern "C" int hb_arena_snapshot(int nseg, hb_arena* const* src, const int64_t* src_off, const int64_t* dst_off, const int64_t* count, void* dst) {
    int rc = ensure_device();
    if (rc) return rc;
    if (nseg < 0 || (nseg > 0 && (!src || !src_off || !dst_off || !count || !dst))) return fail(HB_ERR_ARG, "bad hb_arena_snapshot arguments");
    for (int base = 0; base < nseg; base += HB_SNAP_MAX) {
        HbSnapshot s;
        memset(&s, 0, sizeof(s));
        const int m = nseg - base < HB_SNAP_MAX ? nseg - base : HB_SNAP_MAX;
        long long most = 0;
        for (int i = 0; i < m; ++i) {
            const hb_arena* a = src[base + i];
            if (!a || src_off[base + i] < 0 || count[base + i] < 0 || src_off[base + i] + count[base + i] > a->n)
                return fail(HB_ERR_ARG, "hb_arena_snapshot: segment out of range");
            s.src[i] = a->ptr + src_off[base + i];
            s.dst[i] = dst_off[base + i];
            s.count[i] = count[base + i];
            if (count[base + i] > most) most = count[base + i];
        }
        if (most == 0) continue;
        long long gx = (most + 1023) / 1024;
        if (gx > 1024) gx = 1024;
        hb_snapshot_kernel<<<dim3((unsigned)gx, (unsigned)m, 1), 256, 0, g_stream>>>(s, (double*)dst);
        CUDA_TRY(cudaGetLastError());
    }
    return HB_OK;
}

extern "C" int hb_arena_scale(hb_arena* dst, double alpha) {
    int rc = ensure_device();
    if (rc) return rc;
    if (!dst) return fail(HB_ERR_ARG, "null arena");
    if (dst->n == 0) return HB_OK;
    hb_scale_kernel<<<grid_for(dst->n), 256, 0, g_stream>>>(dst->ptr, alpha, dst->n);
    CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

// Projection (histbook/proj.py:227-296 `project`: numpy.sum of the content over the fixed axes that are not kept):
// every thread walks input rows (bins), finds the output row by dropping the summed axes from the index, and adds
// the row's columns there -- into a block-private shared-memory copy of the output when it is small (then one RED
// per non-zero output element and block), directly with float64 REDs otherwise.  Zero input elements are skipped.
struct HbProject {
    int ndim;
    long long shape[8];       // bins per fixed axis
    long long ostride[8];     // stride of the axis in the output row index, 0 if the axis is summed over
};

__global__ void hb_project_kernel(const double* __restrict__ src, double* __restrict__ dst, long long nrows, int ncol,
                                  long long orows, int use_smem, const HbProject pr) {
    extern __shared__ double sm[];
    const long long osize = orows * ncol;
    if (use_smem) {
        for (long long i = threadIdx.x; i < osize; i += blockDim.x) sm[i] = 0.0;
        __syncthreads();
    }
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nrows; b += (long long)gridDim.x * blockDim.x) {
        long long rest = b, orow = 0;
        for (int k = pr.ndim - 1; k >= 0; --k) {
            const long long t = pr.shape[k];
            orow += (rest % t) * pr.ostride[k];
            rest /= t;
        }
        for (int c = 0; c < ncol; ++c) {
            const double v = src[b * ncol + c];
            if (v != 0.0) {
                if (use_smem) atomicAdd(&sm[orow * ncol + c], v);
                else hb_red_f64(dst + orow * ncol + c, v);
            }
        }
    }
    if (use_smem) {
        __syncthreads();
        for (long long i = threadIdx.x; i < osize; i += blockDim.x)
            if (sm[i] != 0.0) hb_red_f64(dst + i, sm[i]);
    }
}

extern "C" int hb_arena_project(const hb_arena* src, int ndim, const int64_t* shape, int ncol, const int32_t* keep, hb_arena** out) {
    int rc = ensure_device();
    if (rc) return rc;
    if (!src || !shape || !keep || !out || ndim < 1 || ndim > 8 || ncol < 1) return fail(HB_ERR_ARG, "bad hb_arena_project arguments");
    HbProject pr;
    memset(&pr, 0, sizeof(pr));
    pr.ndim = ndim;
    long long nrows = 1, orows = 1;
    for (int k = 0; k < ndim; ++k) { pr.shape[k] = shape[k]; nrows *= shape[k]; }
    for (int k = ndim - 1; k >= 0; --k) {
        if (keep[k]) { pr.ostride[k] = orows; orows *= shape[k]; } else pr.ostride[k] = 0;
    }
    if (nrows * ncol > src->n) return fail(HB_ERR_ARG, "hb_arena_project: shape larger than the arena");
    rc = hb_arena_create(orows * ncol, out);
    if (rc) return rc;
    if (nrows == 0) return HB_OK;
    const long long osize = orows * ncol;
    const int use_smem = osize * 8 <= 32 * 1024 ? 1 : 0;
    int grid = grid_for(nrows);
    if (use_smem && grid > g_sm_count * 4) grid = g_sm_count * 4;
    hb_project_kernel<<<grid, 256, use_smem ? (size_t)osize * 8 : 0, g_stream>>>(src->ptr, (*out)->ptr, nrows, ncol, orows, use_smem, pr);
    CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

// Hist.select (histbook/proj.py:470-496, `content[slc].copy()` with a slice on ONE axis) on the device: the arena is viewed
// as [outer][len][inner], element k of the cut axis comes from start + k * step.
__global__ void hb_slice_kernel(const double* __restrict__ src, double* __restrict__ dst, long long total, long long len,
                                long long inner, long long start, long long step, long long count) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long in = i % inner;
        const long long rest = i / inner;
        const long long k = rest % count;
        const long long o = rest / count;
        dst[i] = src[(o * len + start + k * step) * inner + in];
    }
}

extern "C" int hb_arena_slice(const hb_arena* src, int64_t outer, int64_t len, int64_t inner, int64_t start, int64_t step, int64_t count, hb_arena** out) {
    int rc = ensure_device();
    if (rc) return rc;
    if (!src || !out || outer < 0 || len < 0 || inner < 1 || count < 0 || step == 0) return fail(HB_ERR_ARG, "bad hb_arena_slice arguments");
    if (outer * len * inner > src->n) return fail(HB_ERR_ARG, "hb_arena_slice: shape larger than the arena");
    if (count > 0) {
        const long long last = start + (count - 1) * step;
        if (start < 0 || start >= len || last < 0 || last >= len) return fail(HB_ERR_ARG, "hb_arena_slice: slice outside the axis");
    }
    const long long total = outer * count * inner;
    rc = hb_arena_create(total, out);
    if (rc) return rc;
    if (total == 0) return HB_OK;
    hb_slice_kernel<<<grid_for(total), 256, 0, g_stream>>>(src->ptr, (*out)->ptr, total, len, inner, start, step, count);
    CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

// Hist.table (histbook/proj.py:577-627, `handlearray`) on the device: per bin the (weighted) count, its error, the effective
// count and mean / error of the mean of the profiles, operation for operation what numpy does there (IEEE multiply, divide,
// subtract and square root, no contraction into FMAs), so the columns are bit-identical to the reference's.
#define HB_TABLE_MAX_PROF 16
struct HbTable {
    int ncol, sumw, sumw2, flags, nprof, ncolumns;
    int sumwx[HB_TABLE_MAX_PROF], sumwx2[HB_TABLE_MAX_PROF];
};
__global__ void hb_table_kernel(const double* __restrict__ src, double* __restrict__ dst, long long nrows, const HbTable t) {
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nrows; b += (long long)gridDim.x * blockDim.x) {
        const double* __restrict__ row = src + b * t.ncol;
        double* __restrict__ o = dst + b * t.ncolumns;
        const double sumw = row[t.sumw];
        const double sumw2 = t.sumw2 >= 0 ? row[t.sumw2] : 0.0;
        int k = 0;
        if (t.flags & 1) {
            o[k++] = sumw;
            if (t.flags & 2) o[k++] = sqrt(t.sumw2 >= 0 ? sumw2 : sumw);
        }
        const bool good = sumw > 0.0;
        const double effcnt = t.sumw2 >= 0 ? __ddiv_rn(__dmul_rn(sumw, sumw), sumw2) : sumw;
        if (t.flags & 4) o[k++] = good ? effcnt : 0.0;
        for (int j = 0; j < t.nprof; ++j) {
            const double mean = __ddiv_rn(row[t.sumwx[j]], sumw);
            o[k++] = good ? mean : 0.0;
            if (t.flags & 2) {
                const double var = __dsub_rn(__ddiv_rn(row[t.sumwx2[j]], sumw), __dmul_rn(mean, mean));
                o[k++] = good ? sqrt(__ddiv_rn(var, effcnt)) : 0.0;
            }
        }
    }
}

extern "C" int hb_arena_table(const hb_arena* src, int64_t nrows, int ncol, int sumw, int sumw2, int flags, int nprof,
                              const int32_t* sumwx, const int32_t* sumwx2, hb_arena** out) {
    int rc = ensure_device();
    if (rc) return rc;
    if (!src || !out || nrows < 0 || ncol < 1 || sumw < 0 || sumw >= ncol || sumw2 >= ncol || nprof < 0 || nprof > HB_TABLE_MAX_PROF ||
        (nprof > 0 && (!sumwx || !sumwx2)))
        return fail(HB_ERR_ARG, "bad hb_arena_table arguments");
    if (nrows * ncol > src->n) return fail(HB_ERR_ARG, "hb_arena_table: shape larger than the arena");
    HbTable t;
    memset(&t, 0, sizeof(t));
    t.ncol = ncol; t.sumw = sumw; t.sumw2 = sumw2; t.flags = flags; t.nprof = nprof;
    for (int j = 0; j < nprof; ++j) {
        if (sumwx[j] < 0 || sumwx[j] >= ncol || sumwx2[j] < 0 || sumwx2[j] >= ncol) return fail(HB_ERR_ARG, "hb_arena_table: profile column outside the row");
        t.sumwx[j] = sumwx[j]; t.sumwx2[j] = sumwx2[j];
    }
    t.ncolumns = ((flags & 1) ? ((flags & 2) ? 2 : 1) : 0) + ((flags & 4) ? 1 : 0) + nprof * ((flags & 2) ? 2 : 1);
    if (t.ncolumns == 0) return fail(HB_ERR_ARG, "hb_arena_table: no columns asked for");
    rc = hb_arena_create(nrows * t.ncolumns, out);
    if (rc) return rc;
    if (nrows == 0) return HB_OK;
    hb_table_kernel<<<grid_for(nrows), 256, 0, g_stream>>>(src->ptr, (*out)->ptr, nrows, t);
    CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

extern "C" void* hb_arena_ptr(const hb_arena* a) { return a ? (void*)a->ptr : nullptr; }
extern "C" int64_t hb_arena_size(const hb_arena* a) { return a ? a->n : 0; }

// ------------------------------------------------------------------------------------------------
// fill

static const int kDtypeSize[11] = {8, 4, 8, 4, 2, 1, 8, 4, 2, 1, 1};

// Host columns travel in chunks through two device staging buffers per column (chunk k + 1 is copied while the kernels
// read chunk k).  Page-locked host memory (cudaHostAlloc / cudaHostRegister, torch pin_memory) is copied from where it
// lies.  Pageable memory -- what numpy.random or pandas hand out -- first goes through a ring of page-locked bounce
// buffers, filled by a few worker threads (one memcpy stream per thread): cudaMemcpyAsync from pageable memory is a
// synchronous, single-threaded staging copy inside the driver (measured 9.4 GB/s against 44.7 GB/s from page-locked
// memory, BENCH r2 c4), the ring restores the overlap of the host copy, the DMA and the kernels.
#define HB_RING 3
struct Staging {
    void* dev[2][HB_MAX_COLS] = {};
    size_t cap[2][HB_MAX_COLS] = {};
    cudaEvent_t copied[2] = {};
    cudaEvent_t consumed[2] = {};
    void* pin[HB_RING][HB_MAX_COLS] = {};
    size_t pin_cap[HB_RING][HB_MAX_COLS] = {};
    cudaEvent_t sent[HB_RING] = {};          // the DMA out of ring slot r has completed
    bool init = false;
};
static Staging g_stage;

static int staging_reserve(int buf, int col, size_t bytes) {
    if (g_stage.cap[buf][col] >= bytes) return HB_OK;
    if (g_stage.dev[buf][col]) cudaFree(g_stage.dev[buf][col]);
    g_stage.dev[buf][col] = nullptr;
    g_stage.cap[buf][col] = 0;
    CUDA_TRY(cudaMalloc(&g_stage.dev[buf][col], bytes));
    g_stage.cap[buf][col] = bytes;
    return HB_OK;
}

static int ring_reserve(int slot, int col, size_t bytes) {
    if (g_stage.pin_cap[slot][col] >= bytes) return HB_OK;
    if (g_stage.pin[slot][col]) cudaFreeHost(g_stage.pin[slot][col]);
    g_stage.pin[slot][col] = nullptr;
    g_stage.pin_cap[slot][col] = 0;
    CUDA_TRY(cudaHostAlloc(&g_stage.pin[slot][col], bytes, cudaHostAllocDefault));
    g_stage.pin_cap[slot][col] = bytes;
    return HB_OK;
}

// Counters of the dynamic tile hand-out (HbParams::sched): one pair per stream, because launches on one stream run one
// after the other (the last block of a launch zeroes the pair again) while launches on different streams may overlap.
static std::vector<std::pair<cudaStream_t, hb_u32*>> g_sched;
static int sched_for(cudaStream_t stream, hb_u32** out) {
    for (auto& e : g_sched)
        if (e.first == stream) { *out = e.second; return HB_OK; }
    hb_u32* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, 256));
    CUDA_TRY(cudaMemset(p, 0, 256));
    g_sched.emplace_back(stream, p);
    *out = p;
    return HB_OK;
}

// streams of the parts of a multi-program fill (hb_fill_multi): the kernels of a book that was split run side by side
#define HB_MAX_PARTS 8
static cudaStream_t g_part_stream[HB_MAX_PARTS] = {};
static cudaEvent_t g_part_done[HB_MAX_PARTS] = {};
static cudaEvent_t g_fork = nullptr;

static int launch(hb_program* prog, const HbParams& params, int grid, cudaStream_t stream) {
    void* args[1] = {(void*)&params};
    DRV_TRY(g_drv.LaunchKernel(prog->fn, (unsigned)grid, 1, 1, (unsigned)prog->threads, 1, 1,
                               (unsigned)prog->smem, (CUstream)stream, args, nullptr));
    return HB_OK;
}

extern "C" int hb_fill(hb_program* prog, int ncols, const hb_column* cols, int64_t n,
                       int nhists, hb_arena* const* hists, int naux, const void* const* aux,
                       int nwin, const int32_t* win, int flags, hb_stats* stats) {
    return hb_fill_exact(prog, ncols, cols, n, nhists, hists, nullptr, naux, aux, nwin, win, flags, stats);
}

extern "C" int hb_fill_exact(hb_program* prog, int ncols, const hb_column* cols, int64_t n,
                             int nhists, hb_arena* const* hists, hb_arena* const* exact, int naux, const void* const* aux,
                             int nwin, const int32_t* win, int flags, hb_stats* stats) {
    hb_launch one;
    memset(&one, 0, sizeof(one));
    one.prog = prog;
    one.ncols = ncols;
    one.colmap = nullptr;
    one.nhists = nhists;
    one.hists = hists;
    one.exact = exact;
    one.naux = naux;
    one.aux = aux;
    one.nwin = nwin;
    one.win = win;
    return hb_fill_multi(1, &one, ncols, cols, n, flags, stats);
}

extern "C" int hb_fill_multi(int nlaunch, const hb_launch* L, int ncols, const hb_column* cols, int64_t n, int flags, hb_stats* stats) {
    int rc = ensure_device();
    if (rc) return rc;
    if (nlaunch < 1 || nlaunch > HB_MAX_PARTS || !L || ncols < 0 || ncols > HB_MAX_COLS || n < 0)
        return fail(HB_ERR_ARG, "bad hb_fill arguments");
    std::lock_guard<std::mutex> lock(g_mutex);
    NvtxRange range("hb_fill");

    bool any_host = false, any_pageable = false;
    bool pageable[HB_MAX_COLS] = {};
    for (int c = 0; c < ncols; ++c) {
        if (cols[c].dtype < 0 || cols[c].dtype > 10) return fail(HB_ERR_ARG, "bad column dtype");
        if (cols[c].is_scalar) continue;
        if (!cols[c].ptr && n > 0) return fail(HB_ERR_ARG, "null column pointer");
        if (cols[c].location == 0) {
            any_host = true;
            if (n > 0 && is_pageable(cols[c].ptr)) { pageable[c] = true; any_pageable = true; }
        }
    }
    std::vector<HbParams> base((size_t)nlaunch);
    int tile = 1;
    for (int l = 0; l < nlaunch; ++l) {
        const hb_launch& q = L[l];
        if (!q.prog || q.ncols < 0 || q.ncols > HB_MAX_COLS || q.nhists < 0 || q.nhists > HB_MAX_HISTS || q.naux < 0 || q.naux > HB_MAX_AUX ||
            q.nwin < 0 || q.nwin > HB_MAX_WIN)
            return fail(HB_ERR_ARG, "bad hb_fill arguments");
        HbParams& b = base[l];
        memset(&b, 0, sizeof(b));
        for (int i = 0; i < q.ncols; ++i) {
            const int c = q.colmap ? q.colmap[i] : i;
            if (c < 0 || c >= ncols) return fail(HB_ERR_ARG, "column map out of range");
            if (cols[c].is_scalar) b.scal[i] = cols[c].scalar_bits;
        }
        for (int h = 0; h < q.nhists; ++h) {
            if (!q.hists[h]) return fail(HB_ERR_ARG, "null arena");
            b.hist[h] = q.hists[h]->ptr;
            b.fxg[h] = (q.exact && q.exact[h]) ? reinterpret_cast<hb_u64*>(q.exact[h]->ptr) : nullptr;
        }
        for (int a = 0; a < q.naux; ++a) b.aux[a] = q.aux[a];
        for (int i = 0; i < q.nwin; ++i) b.win[i] = q.win[i];
        // chunk boundaries must be tile boundaries of every program: least common multiple (tiles are threads x 1, 2 or 4)
        int t = q.prog->tile, a = tile, g = t;
        while (a) { int r = g % a; g = a; a = r; }
        tile = tile / g * t;
    }
    // Parts run one after the other on the compute stream unless HB_FILL_CONCURRENT asks for a stream each.  (Measured on
    // c4's two half-books, kernels only, 1e8 events: side by side with one block per SM each 10.0 ms; one after the other with
    // two blocks per SM each 8.3 ms; the whole book as one 1024-thread kernel 9.4 ms -- tools/parts_probe.py.)
    const bool multi = nlaunch > 1 && (flags & HB_FILL_CONCURRENT) != 0;
    if (multi) {
        if (!g_fork) CUDA_TRY(cudaEventCreateWithFlags(&g_fork, cudaEventDisableTiming));
        for (int l = 0; l < nlaunch; ++l) {
            if (!g_part_stream[l]) {
                CUDA_TRY(cudaStreamCreateWithFlags(&g_part_stream[l], cudaStreamNonBlocking));
                CUDA_TRY(cudaEventCreateWithFlags(&g_part_done[l], cudaEventDisableTiming));
            }
        }
    }
    for (int l = 0; l < nlaunch; ++l) {
        rc = sched_for(multi ? g_part_stream[l] : g_stream, &base[l].sched);
        if (rc) return rc;
    }

    // pageable host memory: smaller chunks, so that the three stages (host copy, DMA, kernels) overlap on modest fills too
    int64_t per_launch = any_host ? (any_pageable ? std::min<int64_t>(g_chunk_events, (int64_t)1 << 22) : g_chunk_events) : ((int64_t)1 << 30);
    per_launch = per_launch / tile * tile;
    if (per_launch < tile) per_launch = tile;

    hb_stats st;
    memset(&st, 0, sizeof(st));
    st.blocks_per_sm = L[0].prog->blocks_per_sm;

    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (flags & HB_FILL_TIMED) {
        CUDA_TRY(cudaEventCreate(&t0));
        CUDA_TRY(cudaEventCreate(&t1));
        CUDA_TRY(cudaEventRecord(t0, g_stream));
    }
    if (any_host && !g_stage.init) {
        for (int b = 0; b < 2; ++b) {
            CUDA_TRY(cudaEventCreateWithFlags(&g_stage.copied[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&g_stage.consumed[b], cudaEventDisableTiming));
        }
        for (int r = 0; r < HB_RING; ++r) CUDA_TRY(cudaEventCreateWithFlags(&g_stage.sent[r], cudaEventDisableTiming | cudaEventBlockingSync));
        g_stage.init = true;
    }
    if (multi) {
        // the parts start after everything queued on the compute stream so far (clears, uploads, the previous fill)
        CUDA_TRY(cudaEventRecord(g_fork, g_stream));
        for (int l = 0; l < nlaunch; ++l) CUDA_TRY(cudaStreamWaitEvent(g_part_stream[l], g_fork, 0));
    }

    std::vector<CopyJob> jobs;
    int chunk_index = 0;
    for (int64_t off = 0; off < n; off += per_launch, ++chunk_index) {
        const int64_t m = (n - off < per_launch) ? (n - off) : per_launch;
        const int buf = chunk_index & 1, slot = chunk_index % HB_RING;
        const void* colptr[HB_MAX_COLS] = {};
        if (any_pageable) {
            // ring slot reuse: its last DMA must be through before the workers overwrite it
            CUDA_TRY(cudaEventSynchronize(g_stage.sent[slot]));
            jobs.clear();
            copy_pool();
            const int np = g_copy_threads > 0 ? g_copy_threads : 1;
            for (int c = 0; c < ncols; ++c) {
                if (cols[c].is_scalar || !pageable[c]) continue;
                const size_t esize = (size_t)kDtypeSize[cols[c].dtype], bytes = (size_t)m * esize;
                rc = ring_reserve(slot, c, (size_t)per_launch * esize);
                if (rc) return rc;
                const char* src = (const char*)cols[c].ptr + (size_t)off * esize;
                const size_t step = ((bytes + np - 1) / np + 4095) / 4096 * 4096;
                for (size_t o = 0; o < bytes; o += step)
                    jobs.push_back({(char*)g_stage.pin[slot][c] + o, src + o, std::min(step, bytes - o)});
            }
            { NvtxRange r2("hb_fill: pageable -> page-locked ring"); copy_pool()->copy(jobs); }
        }
        // staging buffer reuse: the copy waits until the last kernel that read this buffer is done -- the one two
        // chunks back, or the tail of the PREVIOUS hb_fill call (waiting on a never-recorded event is a no-op)
        if (any_host)
            CUDA_TRY(cudaStreamWaitEvent(g_copy_stream, g_stage.consumed[buf], 0));
        for (int c = 0; c < ncols; ++c) {
            if (cols[c].is_scalar) continue;
            const size_t esize = (size_t)kDtypeSize[cols[c].dtype];
            const char* src = (const char*)cols[c].ptr + (size_t)off * esize;
            if (cols[c].location == 0) {
                rc = staging_reserve(buf, c, (size_t)per_launch * esize);
                if (rc) return rc;
                const void* from = pageable[c] ? g_stage.pin[slot][c] : (const void*)src;
                CUDA_TRY(cudaMemcpyAsync(g_stage.dev[buf][c], from, (size_t)m * esize, cudaMemcpyHostToDevice, g_copy_stream));
                st.h2d_bytes += (int64_t)((size_t)m * esize);
                colptr[c] = g_stage.dev[buf][c];
            } else {
                colptr[c] = src;
            }
        }
        if (any_pageable) CUDA_TRY(cudaEventRecord(g_stage.sent[slot], g_copy_stream));
        if (any_host) CUDA_TRY(cudaEventRecord(g_stage.copied[buf], g_copy_stream));
        for (int l = 0; l < nlaunch; ++l) {
            const hb_launch& q = L[l];
            cudaStream_t stream = multi ? g_part_stream[l] : g_stream;
            HbParams p = base[l];
            p.n = m;
            p.flags = flags;
            bool aligned = true;
            for (int i = 0; i < q.ncols; ++i) {
                const int c = q.colmap ? q.colmap[i] : i;
                if (cols[c].is_scalar) continue;
                p.cols[i] = colptr[c];
                if (((uintptr_t)p.cols[i]) % (4 * (size_t)kDtypeSize[cols[c].dtype]) != 0) aligned = false;
            }
            p.vec_ok = aligned ? 1 : 0;
            if (any_host) CUDA_TRY(cudaStreamWaitEvent(stream, g_stage.copied[buf], 0));
            const int ptile = q.prog->tile;
            int per_sm = q.prog->blocks_per_sm;
            if (q.prog->max_blocks_per_sm > 0 && per_sm > q.prog->max_blocks_per_sm) per_sm = q.prog->max_blocks_per_sm;
            const int resident = g_sm_count * per_sm;
            int64_t ntiles = (m + ptile - 1) / ptile;
            int grid = (int)(ntiles < resident ? ntiles : resident);
            if (grid < 1) grid = 1;
            // equal shares: with r = ceil(ntiles / grid) rounds of the grid-stride loop, ceil(ntiles / r) blocks do r tiles each
            // instead of a last round that keeps a fraction of the SMs busy (c1 at 1e7 events: 8.25 tiles per resident block)
            const int64_t rounds = (ntiles + grid - 1) / grid;
            if (rounds > 1) grid = (int)((ntiles + rounds - 1) / rounds);
            rc = launch(q.prog, p, grid, stream);
            if (rc) return rc;
            st.launches += 1;
            st.grid = grid;
            if (multi) {
                CUDA_TRY(cudaEventRecord(g_part_done[l], stream));
                CUDA_TRY(cudaStreamWaitEvent(g_stream, g_part_done[l], 0));       // join: the compute stream continues after every part
            }
        }
        if (any_host) CUDA_TRY(cudaEventRecord(g_stage.consumed[buf], g_stream));
        if (multi && off + per_launch < n) {
            CUDA_TRY(cudaEventRecord(g_fork, g_stream));
            for (int l = 0; l < nlaunch; ++l) CUDA_TRY(cudaStreamWaitEvent(g_part_stream[l], g_fork, 0));
        }
    }
    st.events = n;
    if (any_host) CUDA_TRY(cudaStreamSynchronize(g_copy_stream));   // caller may now reuse / free its host columns
    if (flags & HB_FILL_TIMED) {
        CUDA_TRY(cudaEventRecord(t1, g_stream));
        CUDA_TRY(cudaEventSynchronize(t1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, t0, t1));
        st.ms_kernel = ms;
        cudaEventDestroy(t0);
        cudaEventDestroy(t1);
    }
    if (stats) *stats = st;
    return HB_OK;
}

// ------------------------------------------------------------------------------------------------
// deterministic mode: exact accumulators -> float64 contents

#define HB_FX_MAXSLOTS 32
struct HbFinalize {
    int k[HB_FX_MAXSLOTS];
    unsigned long long mask[HB_FX_MAXSLOTS];
};

__global__ void hb_fx_finalize_kernel(double* __restrict__ content, hb_u64* __restrict__ g, int64_t nrows, int ncol, int nslots,
                                      const HbFinalize f) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrows * nslots; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / nslots;
        const int j = (int)(i - row * nslots);
        hb_u64* const acc = g + (int64_t)HB_FXG_SLOT * i;
        hb_u64 any = 0ull;
        for (int w = 0; w < HB_FXG_SLOT; ++w) any |= acc[w];
        if (any == 0ull) continue;
        const double v = hb_fxg_value(acc, f.k[j] - 1075);
        for (int w = 0; w < HB_FXG_SLOT; ++w) acc[w] = 0ull;
        unsigned long long m = f.mask[j];
        for (int c = 0; m != 0ull; ++c, m >>= 1)
            if (m & 1ull) content[row * ncol + c] += v;     // one writer per (row, column): no two slots feed the same column
    }
}

extern "C" int hb_fx_finalize(hb_arena* content, hb_arena* exact, int64_t nrows, int ncol, int nslots,
                              const int32_t* k, const uint64_t* dest_mask) {
    int rc = ensure_device();
    if (rc) return rc;
    if (!content || !exact || !k || !dest_mask || nrows < 0 || ncol < 1 || ncol > 64 || nslots < 1 || nslots > HB_FX_MAXSLOTS)
        return fail(HB_ERR_ARG, "bad hb_fx_finalize arguments");
    if (nrows * ncol > content->n || nrows * nslots * HB_FXG_SLOT > exact->n) return fail(HB_ERR_ARG, "hb_fx_finalize: arenas too small");
    HbFinalize f;
    memset(&f, 0, sizeof(f));
    unsigned long long seen = 0ull;
    for (int j = 0; j < nslots; ++j) {
        f.k[j] = k[j];
        f.mask[j] = dest_mask[j];
        if (seen & dest_mask[j]) return fail(HB_ERR_ARG, "hb_fx_finalize: two slots feed one column");
        seen |= dest_mask[j];
    }
    if (nrows == 0) return HB_OK;
    hb_fx_finalize_kernel<<<grid_for(nrows * nslots), 256, 0, g_stream>>>(content->ptr, reinterpret_cast<hb_u64*>(exact->ptr), nrows, ncol, nslots, f);
    CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

// ------------------------------------------------------------------------------------------------
// key sets

struct hb_keyset {
    hb_u64* table = nullptr;     // HbKeySet header (4 words) + slots
    int capacity = 0;
};

__global__ void hb_keyset_init_kernel(hb_u64* table, int capacity) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { table[0] = (hb_u64)capacity; table[1] = 0; table[2] = 0; table[3] = 0; }
    if (i < capacity) table[4 + i] = HB_KEY_EMPTY;
}

extern "C" int hb_keyset_clear(hb_keyset* ks) {
    int rc = ensure_device();
    if (rc) return rc;
    if (!ks) return fail(HB_ERR_ARG, "null keyset");
    hb_keyset_init_kernel<<<(ks->capacity + 255) / 256, 256, 0, g_stream>>>(ks->table, ks->capacity);
    CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

extern "C" int hb_keyset_create(int capacity, hb_keyset** out) {
    int rc = ensure_device();
    if (rc) return rc;
    if (capacity < 2 || (capacity & (capacity - 1)) != 0 || !out) return fail(HB_ERR_ARG, "keyset capacity must be a power of two");
    hb_keyset* ks = new hb_keyset();
    ks->capacity = capacity;
    cudaError_t e = cudaMalloc((void**)&ks->table, (size_t)(capacity + 4) * sizeof(hb_u64));
    if (e != cudaSuccess) { delete ks; return fail(HB_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    rc = hb_keyset_clear(ks);
    if (rc) { cudaFree(ks->table); delete ks; return rc; }
    *out = ks;
    return HB_OK;
}

extern "C" int hb_keyset_destroy(hb_keyset* ks) {
    if (!ks) return HB_OK;
    if (ks->table) cudaFree(ks->table);
    delete ks;
    return HB_OK;
}

extern "C" void* hb_keyset_ptr(const hb_keyset* ks) { return ks ? (void*)ks->table : nullptr; }

extern "C" int hb_keyset_read(const hb_keyset* ks, uint64_t* keys_out, int max_keys, int* nkeys, int* overflowed) {
    int rc = ensure_device();
    if (rc) return rc;
    if (!ks || !nkeys) return fail(HB_ERR_ARG, "null keyset");
    std::vector<hb_u64> host((size_t)ks->capacity + 4);
    CUDA_TRY(cudaMemcpyAsync(host.data(), ks->table, host.size() * sizeof(hb_u64), cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    int count = 0;
    for (int i = 0; i < ks->capacity; ++i) {
        if (host[4 + i] != HB_KEY_EMPTY) {
            if (keys_out && count < max_keys) keys_out[count] = host[4 + i];
            ++count;
        }
    }
    *nkeys = count;
    if (overflowed) *overflowed = (int)host[2];
    if (count > max_keys && keys_out) return fail(HB_ERR_OVERFLOW, "more keys than the output buffer holds");
    return HB_OK;
}

// ------------------------------------------------------------------------------------------------
// utilities

static void* g_flush_buf = nullptr;
static const size_t kFlushBytes = (size_t)256 << 20;     // 256 MiB > 126 MB L2

extern "C" int hb_flush_l2(void) {
    int rc = ensure_device();
    if (rc) return rc;
    if (!g_flush_buf) {
        CUDA_TRY(cudaMalloc(&g_flush_buf, 2 * kFlushBytes + 256));
        CUDA_TRY(cudaMemsetAsync(g_flush_buf, 0, 2 * kFlushBytes + 256, g_stream));
    }
    const int64_t n4 = (int64_t)(kFlushBytes / sizeof(float4));
    hb_flush_kernel<<<g_sm_count * 8, 256, 0, g_stream>>>((float4*)g_flush_buf, n4);
    hb_flush_read_kernel<<<g_sm_count * 8, 256, 0, g_stream>>>((const float4*)g_flush_buf + n4, n4, (float*)((char*)g_flush_buf + 2 * kFlushBytes));
    CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

extern "C" int hb_memcpy_h2d(void* dst, const void* src, size_t bytes) {
    int rc = ensure_device();
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return HB_OK;
}

extern "C" int hb_memcpy_d2h(void* dst, const void* src, size_t bytes) {
    int rc = ensure_device();
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return HB_OK;
}

extern "C" int hb_device_malloc(size_t bytes, void** out) {
    int rc = ensure_device();
    if (rc) return rc;
    CUDA_TRY(cudaMalloc(out, bytes ? bytes : 1));
    return HB_OK;
}

extern "C" int hb_device_free(void* p) {
    if (p) cudaFree(p);
    return HB_OK;
}
