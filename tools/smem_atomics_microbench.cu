// smem_atomics_microbench.cu -- what does one scattered update of a block-private (shared-memory) histogram
// cost on sm_100a, per primitive?  No global traffic: every thread issues ITERS updates to pseudo-random slots of
// a shared table, so the number is the throughput of the shared-memory / LSU data pipe for that primitive.
// Not part of the product path; it exists so that the accumulate strategies in histbook_b200/codegen.py are
// chosen from measurements (profiles/smem_atomics_microbench_r01.txt).
//
//   ./smem_atomics_microbench
//
// Output per (primitive, table size): updates per clock per SM (at the clock rate measured with clock64 vs
// the CUDA-event time) and the equivalent events/s for one update per event.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef long long i64;
typedef unsigned long long u64;
typedef unsigned int u32;

enum Mode {
    ADD_U32 = 0,        // atomicAdd(u32), result unused       -> ATOMS.ADD / ATOMS.POPC.INC
    ADD_U32_RET,        // atomicAdd(u32), result used         -> ATOMS.ADD with return
    ADD_U64,            // atomicAdd(u64)                      -> ?
    ADD_F32,            // atomicAdd(float)                    -> ?
    ADD_F64,            // atomicAdd(double)                   -> CAS.64 spin loop
    CAS128_PAIR,        // LDS.128 + ATOMS.CAS.128 loop on (double, double)
    FIXED_2LIMB,        // 64-bit fixed point as two u32 limbs: returning add (carry out) + plain add
    FIXED_3LIMB,        // 96-bit fixed point: two returning adds + one plain add
    LDS_STS_64,         // non-atomic read-modify-write of a double (racy: throughput reference only)
    LDS_STS_128,        // non-atomic read-modify-write of (double, double)
    EXCH_64,            // atomicExch(u64)
    FIXED_2LIMB_SOA,    // as FIXED_2LIMB, limbs in separate arrays (all 32 banks in play)
    FIXED_3LIMB_SOA,    // as FIXED_3LIMB, limbs in separate arrays
    FIXED_3LIMB_SOA_X2, // two independent 3-limb updates per iteration, stages interleaved (ILP across events)
    FIXED_4LIMB_SKIP,   // 128-bit field, value placed at a random bit offset: 2-3 limbs touched, carry limb only when needed
    LOCK32_ROW128,      // u32 lock per bin (ATOMS.EXCH), then LDS.128/STS.128 of the (double,double) row, STS unlock
    LOCK64_VALUE,       // the double itself is the lock: EXCH.64 takes it out (sentinel left behind), STS.64 puts the sum back
    LOCK32_ROW3,        // u32 lock, row of 3 doubles (LDS.128+LDS.64 / STS.128+STS.64) + count bump under the lock
    NMODES
};
static const char* kNames[NMODES] = {"atoms.add.u32", "atoms.add.u32 (returning)", "atomicAdd(u64)", "atomicAdd(float)", "atomicAdd(double)",
                                     "cas.128 (double,double)", "fixed 2x u32 limbs", "fixed 3x u32 limbs", "lds.64+sts.64 (racy)",
                                     "lds.128+sts.128 (racy)", "atoms.exch.64",
                                     "fixed 2 limbs SoA", "fixed 3 limbs SoA", "fixed 3 limbs SoA x2 ILP", "fixed 4 limbs skip SoA",
                                     "lock32 + row128", "lock = value (exch.64)", "lock32 + row of 3 f64 + count"};

#define THREADS 512
#define ITERS 2048

__device__ __forceinline__ void sadd_f64x2(double* addr, double a, double b) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(addr);
    u64 lo, hi;
    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(sa) : "memory");
    while (true) {
        const u64 nlo = (u64)__double_as_longlong(__longlong_as_double((i64)lo) + a);
        const u64 nhi = (u64)__double_as_longlong(__longlong_as_double((i64)hi) + b);
        u64 olo, ohi;
        asm volatile("{\n .reg .b128 c, n, o;\n mov.b128 c, {%2,%3};\n mov.b128 n, {%4,%5};\n"
                     " atom.shared.cas.b128 o, [%6], c, n;\n mov.b128 {%0,%1}, o;\n}"
                     : "=l"(olo), "=l"(ohi) : "l"(lo), "l"(hi), "l"(nlo), "l"(nhi), "r"(sa) : "memory");
        if (olo == lo && ohi == hi) break;
        lo = olo; hi = ohi;
    }
}

// `nslots` table entries (power of two); entry stride in bytes is the natural one of the primitive.
// `hot` = 0: uniform over the table; otherwise slots are drawn from a triangular-ish distribution (sum of two
// uniforms over a quarter of the table) to mimic a peaked histogram.
template <int MODE>
__global__ void __launch_bounds__(THREADS) k(int nslots, int hot, u64* out, i64* cycles) {
    extern __shared__ __align__(16) unsigned char smem[];
    u32* s32 = reinterpret_cast<u32*>(smem);
    u64* s64 = reinterpret_cast<u64*>(smem);
    double* sd = reinterpret_cast<double*>(smem);
    float* sf = reinterpret_cast<float*>(smem);
    const int words = nslots * 4;                       // 16 bytes per slot reserved
    for (int i = threadIdx.x; i < words; i += THREADS) s32[i] = 0u;
    __syncthreads();
    u32 rng = (blockIdx.x * THREADS + threadIdx.x) * 2654435761u + 12345u;
    const u32 mask = (u32)nslots - 1u;
    const i64 t0 = clock64();
    u32 sink = 0;
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        rng = rng * 1664525u + 1013904223u;
        u32 slot;
        if (hot) { const u32 a = (rng >> 8) & (mask >> 2), b = (rng >> 20) & (mask >> 2); slot = (a + b) >> 1; }
        else slot = (rng >> 8) & mask;
        const double v = 1.0 + (double)(rng & 255u) * (1.0 / 256.0);
        if (MODE == ADD_U32) atomicAdd(&s32[slot], 1u + (rng & 3u));
        if (MODE == ADD_U32_RET) sink += atomicAdd(&s32[slot], 1u + (rng & 3u));
        if (MODE == ADD_U64) atomicAdd(&s64[slot], (u64)(rng | 1u));
        if (MODE == ADD_F32) atomicAdd(&sf[slot], (float)v);
        if (MODE == ADD_F64) atomicAdd(&sd[slot], v);
        if (MODE == CAS128_PAIR) sadd_f64x2(&sd[2 * slot], v, v * v);
        if (MODE == FIXED_2LIMB) {
            const u64 q = (u64)__double2ll_rz(v * 4294967296.0);      // 32.32 fixed point
            const u32 lo = (u32)q, hi = (u32)(q >> 32);
            const u32 old = atomicAdd(&s32[2 * slot], lo);
            const u32 carry = (old + lo < old) ? 1u : 0u;
            atomicAdd(&s32[2 * slot + 1], hi + carry);
        }
        if (MODE == FIXED_3LIMB) {
            const u64 q = (u64)__double2ll_rz(v * 4294967296.0);
            const u32 l0 = (u32)(q << 8), l1 = (u32)(q >> 24), l2 = (u32)(q >> 56);
            const u32 o0 = atomicAdd(&s32[4 * slot], l0);
            const u32 c0 = (o0 + l0 < o0) ? 1u : 0u;
            const u32 a1 = l1 + c0;
            const u32 o1 = atomicAdd(&s32[4 * slot + 1], a1);
            const u32 c1 = ((a1 < l1) || (o1 + a1 < o1)) ? 1u : 0u;
            atomicAdd(&s32[4 * slot + 2], l2 + c1);
        }
        if (MODE == LDS_STS_64) { volatile double* p = &sd[slot]; *p = *p + v; }
        if (MODE == LDS_STS_128) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(&sd[2 * slot]);
            double a, b;
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(sa) : "memory");
            a += v; b += v * v;
            asm volatile("st.shared.v2.f64 [%0], {%1,%2};" :: "r"(sa), "d"(a), "d"(b) : "memory");
        }
        if (MODE == EXCH_64) sink += (u32)atomicExch(&s64[slot], (u64)rng);
        if (MODE == FIXED_2LIMB_SOA) {
            const u64 q = (u64)__double2ll_rz(v * 4294967296.0);
            const u32 lo = (u32)q, hi = (u32)(q >> 32);
            const u32 old = atomicAdd(&s32[slot], lo);
            const u32 carry = (old + lo < old) ? 1u : 0u;
            atomicAdd(&s32[nslots + slot], hi + carry);
        }
        if (MODE == FIXED_3LIMB_SOA) {
            const u64 q = (u64)__double2ll_rz(v * 4294967296.0);
            const u32 l0 = (u32)(q << 8), l1 = (u32)(q >> 24), l2 = (u32)(q >> 56);
            const u32 o0 = atomicAdd(&s32[slot], l0);
            const u32 c0 = (o0 + l0 < o0) ? 1u : 0u;
            const u32 a1 = l1 + c0;
            const u32 o1 = atomicAdd(&s32[nslots + slot], a1);
            const u32 c1 = ((a1 < l1) || (o1 + a1 < o1)) ? 1u : 0u;
            atomicAdd(&s32[2 * nslots + slot], l2 + c1);
        }
        if (MODE == FIXED_3LIMB_SOA_X2) {
            const u32 slotb = (rng >> 4) & mask;
            const u64 q = (u64)__double2ll_rz(v * 4294967296.0), qb = q + (rng & 0xffffu);
            const u32 l0 = (u32)(q << 8), l1 = (u32)(q >> 24), l2 = (u32)(q >> 56);
            const u32 m0 = (u32)(qb << 8), m1 = (u32)(qb >> 24), m2 = (u32)(qb >> 56);
            const u32 o0 = atomicAdd(&s32[slot], l0);
            const u32 p0 = atomicAdd(&s32[slotb], m0);
            const u32 c0 = (o0 + l0 < o0) ? 1u : 0u, d0 = (p0 + m0 < p0) ? 1u : 0u;
            const u32 a1 = l1 + c0, b1 = m1 + d0;
            const u32 o1 = atomicAdd(&s32[nslots + slot], a1);
            const u32 p1 = atomicAdd(&s32[nslots + slotb], b1);
            const u32 c1 = ((a1 < l1) || (o1 + a1 < o1)) ? 1u : 0u, d1 = ((b1 < m1) || (p1 + b1 < p1)) ? 1u : 0u;
            atomicAdd(&s32[2 * nslots + slot], l2 + c1);
            atomicAdd(&s32[2 * nslots + slotb], m2 + d1);
        }
        if (MODE == FIXED_4LIMB_SKIP) {
            // 53-bit mantissa at bit offset sh (0..50) of a 128-bit field
            const u64 m = (u64)__double_as_longlong(v) & 0xfffffffffffffull | 0x10000000000000ull;
            const u32 sh = (rng >> 3) % 51u;
            const u32 w = sh >> 5, b = sh & 31u;
            // three consecutive 32-bit pieces of m << b
            const u32 q0 = (u32)(m << b), q1 = (u32)((m << b) >> 32), q2 = b ? (u32)(m >> (64 - b)) : 0u;
            u32 carry = 0;
            u32* base = &s32[w * nslots + slot];
            if (q0) { const u32 o = atomicAdd(base, q0); carry = (o + q0 < o) ? 1u : 0u; }
            { const u32 a = q1 + carry; const u32 o = atomicAdd(base + nslots, a); carry = ((a < q1) || (o + a < o)) ? 1u : 0u; }
            if (q2 | carry) {
                const u32 a = q2 + carry;
                if (w < 1) { const u32 o = atomicAdd(base + 2 * nslots, a); carry = ((a < q2) || (o + a < o)) ? 1u : 0u;
                             if (carry) atomicAdd(base + 3 * nslots, 1u); }
                else atomicAdd(base + 2 * nslots, a);
            }
        }
        if (MODE == LOCK32_ROW128) {
            // 32-byte rows: [f64 | f64 | lock u32 | pad]; half the slots
            const u32 sl = slot >> 1;
            const unsigned sa = (unsigned)__cvta_generic_to_shared(&s32[8 * sl]);
            bool done = false;
            while (!done) {
                if (atomicExch(&s32[8 * sl + 4], 1u) == 0u) {
                    double a, b;
                    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(sa) : "memory");
                    a += v; b += v * v;
                    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" :: "r"(sa), "d"(a), "d"(b) : "memory");
                    asm volatile("st.shared.u32 [%0+16], %1;" :: "r"(sa), "r"(0u) : "memory");
                    done = true;
                }
            }
        }
        if (MODE == LOCK64_VALUE) {
            const u64 BUSY = 0xFFF0DEADBEEFCAFEull;
            bool done = false;
            while (!done) {
                const u64 old = atomicExch(&s64[slot], BUSY);
                if (old != BUSY) {
                    *(volatile double*)&sd[slot] = __longlong_as_double((i64)old) + v;
                    done = true;
                }
            }
        }
        if (MODE == LOCK32_ROW3) {
            // rows of 32 bytes: [lock u32 | count u32 | f64 | f64 | f64] -> needs 32 B per slot: use half the slots
            const u32 sl = slot >> 1;
            u32* row = &s32[8 * sl];
            const unsigned sa = (unsigned)__cvta_generic_to_shared(row);
            bool done = false;
            while (!done) {
                if (atomicExch(row, 1u) == 0u) {
                    u32 cnt; double a, b, c;
                    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(cnt) : "r"(sa) : "memory");
                    asm volatile("ld.shared.f64 %0, [%1+8];" : "=d"(a) : "r"(sa) : "memory");
                    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+16];" : "=d"(b), "=d"(c) : "r"(sa) : "memory");
                    a += v; b += v * v; c += v * v * v * v; cnt += 1u;
                    asm volatile("st.shared.v2.f64 [%0+16], {%1,%2};" :: "r"(sa), "d"(b), "d"(c) : "memory");
                    // lock word and count released/updated with the first double in one 16-byte store
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(sa), "r"(0u), "r"(cnt), "r"((u32)__double2loint(a)), "r"((u32)__double2hiint(a)) : "memory");
                    done = true;
                }
            }
        }
    }
    const i64 t1 = clock64();
    __syncthreads();
    u64 total = sink;
    for (int i = threadIdx.x; i < words; i += THREADS) total += s32[i];
    if (total == 0x123456789abcull) out[0] = total;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(int nslots, int hot, int sms, int blocks_per_sm, u64* out, i64* cycles) {
    size_t smem = (size_t)nslots * 16;
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE>, THREADS, smem));
    if (occ < blocks_per_sm) blocks_per_sm = occ;
    int grid = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        CK(cudaEventRecord(e0));
        k<MODE><<<grid, THREADS, smem>>>(nslots, hot, out, cycles);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it >= 1 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    i64 hc[148 * 4];
    CK(cudaMemcpy(hc, cycles, sizeof(i64) * grid, cudaMemcpyDeviceToHost));
    i64 mx = 0; for (int i = 0; i < grid; ++i) if (hc[i] > mx) mx = hc[i];
    const double updates_per_sm = (double)blocks_per_sm * THREADS * ITERS;
    const double per_clk = updates_per_sm / (double)mx;
    const double gups = (double)grid * THREADS * ITERS / (best * 1e-3) * 1e-9;
    printf("%-28s slots=%6d %-7s blocks/SM=%d  %7.3f updates/clk/SM  %8.1f Gupdates/s  (%.3f ms, %lld clk)\n",
           kNames[MODE], nslots, hot ? "peaked" : "uniform", blocks_per_sm, per_clk, gups, best, mx);
    fflush(stdout);
}

template <int MODE>
static void sweep(int sms, u64* out, i64* cycles) {
    run<MODE>(256, 0, sms, 2, out, cycles);
    run<MODE>(4096, 0, sms, 2, out, cycles);
    run<MODE>(4096, 1, sms, 2, out, cycles);
    run<MODE>(4096, 0, sms, 1, out, cycles);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs; %d threads/block, %d updates/thread\n", prop.name, prop.multiProcessorCount, THREADS, ITERS);
    u64* out; i64* cycles;
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&cycles, sizeof(i64) * 148 * 4));
    int sms = prop.multiProcessorCount;
    sweep<ADD_U32>(sms, out, cycles);
    sweep<ADD_U32_RET>(sms, out, cycles);
    sweep<ADD_U64>(sms, out, cycles);
    sweep<ADD_F32>(sms, out, cycles);
    sweep<ADD_F64>(sms, out, cycles);
    sweep<CAS128_PAIR>(sms, out, cycles);
    sweep<FIXED_2LIMB>(sms, out, cycles);
    sweep<FIXED_3LIMB>(sms, out, cycles);
    sweep<LDS_STS_64>(sms, out, cycles);
    sweep<LDS_STS_128>(sms, out, cycles);
    sweep<EXCH_64>(sms, out, cycles);
    sweep<FIXED_2LIMB_SOA>(sms, out, cycles);
    sweep<FIXED_3LIMB_SOA>(sms, out, cycles);
    sweep<FIXED_3LIMB_SOA_X2>(sms, out, cycles);
    sweep<FIXED_4LIMB_SKIP>(sms, out, cycles);
    sweep<LOCK32_ROW128>(sms, out, cycles);
    sweep<LOCK64_VALUE>(sms, out, cycles);
    sweep<LOCK32_ROW3>(sms, out, cycles);
    return 0;
}
