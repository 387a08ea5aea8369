// cuda_emul.h -- TEST INFRASTRUCTURE ONLY (tests/emul_util.py).  A minimal host-side stand-in for the CUDA execution
// model, enough to run the *generated* fill kernels (histbook_b200/codegen.py + csrc/hb_template.cuh) of a book on CPU
// threads: one std::thread per CUDA thread of a block, blocks one after the other, block / named / warp barriers as
// mutex + condition-variable barriers, warp shuffles as an exchange through a per-warp buffer, shared memory as one global
// array, atomics as GCC __atomic builtins.  It exists so that the code generator's logic (tickets, scatter, chunked sums,
// tails, flags, group-slot lookups ...) can be checked against the oracle in the `-m "not gpu"` suite.  It is never
// imported, linked or executed by the product path, which has no CPU fallback (engine.py raises without the device).
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#define __device__
#define __global__
#define __forceinline__ inline
#define __shared__
#define __constant__ static const
#define __align__(n) alignas(n)
#define __launch_bounds__(...)
#define HB_EMU_NOASM(...) do { } while (0)

struct hb_emu_dim { unsigned x, y, z; };
static thread_local hb_emu_dim threadIdx = {0, 0, 0};
static thread_local hb_emu_dim blockIdx = {0, 0, 0};
static hb_emu_dim blockDim = {1, 1, 1};
static hb_emu_dim gridDim = {1, 1, 1};

alignas(128) static unsigned char hb_smem[256 * 1024];

struct double2 { double x, y; };
struct uint2 { unsigned x, y; };
struct ulonglong2 { unsigned long long x, y; };
struct int4 { int x, y, z, w; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }

template <typename T> static inline T min(T a, T b) { return b < a ? b : a; }
template <typename T> static inline T max(T a, T b) { return a < b ? b : a; }
static inline long long min(long long a, int b) { return a < b ? a : (long long)b; }
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }

static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline double __hiloint2double(int hi, int lo) { return __longlong_as_double((long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo)); }
static inline int __double2hiint(double d) { return (int)(__double_as_longlong(d) >> 32); }
static inline int __double2loint(double d) { return (int)(__double_as_longlong(d) & 0xffffffffLL); }
// cvt.rmi.s32.f64: round toward -inf, saturate, NaN -> 0
static inline int __double2int_rd(double t) {
    if (t != t) return 0;
    const double f = floor(t);
    if (f >= 2147483648.0) return 2147483647;
    if (f < -2147483648.0) return -2147483647 - 1;
    return (int)f;
}
static inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline void __nanosleep(unsigned) { }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) {
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)((v << (shift & 31u)) >> 32);
}
static inline double __ull2double_rn(unsigned long long v) { return (double)v; }

// ---- barriers -------------------------------------------------------------------------------------------------------
// A barrier nobody completes (a kernel whose threads do not all arrive: a bug of the generator, or a construct this stand-in does
// not model) must fail the test, not hang it: after HB_EMU_BARRIER_TIMEOUT seconds the launch is marked failed, every barrier
// opens, the threads run to their end and hb_emu_run reports the failure.
#ifndef HB_EMU_BARRIER_TIMEOUT
#define HB_EMU_BARRIER_TIMEOUT 120
#endif
static std::atomic<int> hb_emu_failed{0};
struct hb_emu_barrier {
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0;
    unsigned gen = 0;
    void sync(int count) {
        if (hb_emu_failed.load()) return;
        std::unique_lock<std::mutex> l(m);
        const unsigned g = gen;
        if (++waiting >= count) { waiting = 0; ++gen; cv.notify_all(); }
        else if (!cv.wait_for(l, std::chrono::seconds(HB_EMU_BARRIER_TIMEOUT), [&] { return gen != g || hb_emu_failed.load() != 0; })) {
            hb_emu_failed.store(1);
            waiting = 0;
            ++gen;
            cv.notify_all();
        }
    }
};
static hb_emu_barrier hb_emu_block_bar;
static hb_emu_barrier hb_emu_named[16];
struct hb_emu_warp { hb_emu_barrier bar; unsigned long long buf[32]; };
static hb_emu_warp hb_emu_warps[64];

static inline void __syncthreads() { hb_emu_block_bar.sync((int)blockDim.x); }
static inline void hb_emu_bar(int id, int count) { hb_emu_named[id & 15].sync(count); }

template <typename T> static inline T hb_emu_exchange(T v, int src) {
    hb_emu_warp& w = hb_emu_warps[threadIdx.x >> 5];
    const int lane = (int)(threadIdx.x & 31u);
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof(T));
    w.buf[lane] = bits;
    w.bar.sync(32);
    const unsigned long long got = w.buf[(src >= 0 && src < 32) ? src : lane];
    w.bar.sync(32);
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}
// (full-warp forms only: every caller in hb_template.cuh passes 0xffffffff)
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) { const int lane = (int)(threadIdx.x & 31u); return hb_emu_exchange(v, lane >= d ? lane - d : lane); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int d) { return hb_emu_exchange(v, (int)(threadIdx.x & 31u) ^ d); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return hb_emu_exchange(v, src & 31); }
static inline unsigned __ballot_sync(unsigned, bool pred) {
    unsigned out = 0;
    for (int l = 0; l < 32; ++l) out |= (hb_emu_exchange<unsigned>(pred ? 1u : 0u, l) & 1u) << l;
    return out;
}
static inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0u; }
static inline bool __all_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
static inline unsigned __activemask() { return 0xffffffffu; }
// (called from divergent code -- hb_keyset_insert behind `if (live)` and an early return: no rendezvous here; every lane is
//  its own peer group, i.e. inserts its key itself -- the warp-level deduplication is an optimisation, not semantics)
template <typename T> static inline unsigned __match_any_sync(unsigned, T) { return 1u << (threadIdx.x & 31u); }

static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    unsigned out = 0;
    for (int l = 0; l < 32; ++l) { const unsigned o = hb_emu_exchange<unsigned>(v, l); out = o > out ? o : out; }
    return out;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { hb_emu_warps[threadIdx.x >> 5].bar.sync(32); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- the inline PTX of hb_template.cuh's fixed-point adds, as host functions (tests/emul_util.py substitutes them) ----------
static inline unsigned hb_emu_bfe_20_11(unsigned hi) { return (hi >> 20) & 0x7ffu; }
static inline void hb_emu_neg96(unsigned& q0, unsigned& q1, unsigned& q2) {          // sub.cc / subc.cc / subc: 0 - (q2:q1:q0)
    const unsigned b0 = q0 != 0u ? 1u : 0u;
    const unsigned b1 = (q1 != 0u || b0) ? 1u : 0u;
    q0 = 0u - q0;
    q1 = 0u - q1 - b0;
    q2 = 0u - q2 - b1;
}
static inline unsigned hb_emu_carry(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a + b) >> 32); }
static inline void hb_emu_stage1a(unsigned& a1, unsigned& a2, unsigned& pre, unsigned x1, unsigned x2, unsigned carry) {
    const unsigned long long s1 = (unsigned long long)x1 + carry;                     // add.cc a1 = x1 + carry
    a1 = (unsigned)s1;
    const unsigned long long s2 = (unsigned long long)x2 + (unsigned)(s1 >> 32);      // addc.cc a2 = x2 + 0 + c
    a2 = (unsigned)s2;
    pre = (unsigned)(s2 >> 32);                                                       // addc pre = 0 + 0 + c
}
static inline void hb_emu_stage1b(unsigned& a2, unsigned& pre, unsigned o, unsigned a1) {
    const unsigned c = hb_emu_carry(o, a1);                                           // add.cc t = o + a1
    const unsigned long long s2 = (unsigned long long)a2 + c;                         // addc.cc a2 = a2 + 0 + c
    a2 = (unsigned)s2;
    pre = pre + (unsigned)(s2 >> 32);                                                 // addc pre = pre + 0 + c
}
static inline unsigned hb_emu_stage2(unsigned o, unsigned a2, unsigned hi, unsigned ext) {
    return hi + ext + hb_emu_carry(o, a2);                                            // add.cc u = o + a2; addc top = hi + ext + c
}

// ---- atomics --------------------------------------------------------------------------------------------------------
static inline unsigned atomicMax(unsigned* a, unsigned v) {
    unsigned old = __atomic_load_n(a, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) { }
    return old;
}
static inline unsigned atomicAdd(unsigned* a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicCAS(unsigned long long* a, unsigned long long cmp, unsigned long long v) {
    __atomic_compare_exchange_n(a, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
static inline unsigned long long atomicExch(unsigned long long* a, unsigned long long v) { return __atomic_exchange_n(a, v, __ATOMIC_SEQ_CST); }
static inline double atomicAdd(double* a, double v) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(a);
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    for (;;) {
        double d;
        memcpy(&d, &old, 8);
        d += v;
        unsigned long long want;
        memcpy(&want, &d, 8);
        if (__atomic_compare_exchange_n(p, &old, want, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) { memcpy(&d, &old, 8); return d; }
    }
}
// red.global.add.f64 (hb_template.cuh hb_red_f64): the emulated form the preprocessed template calls
static inline void hb_red_f64(double* addr, double v) { atomicAdd(addr, v); }
