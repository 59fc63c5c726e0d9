// wgsassign_b200 - sm_100a kernels of the WGSassign genotype-likelihood hot path.
//
// Device layout ("G"): float2 G[M][ldg], one (g0,g1) pair per (site, individual); the third
// genotype likelihood is never stored - it is recomputed as 1-g0-g1 exactly like the
// reference does (emMAF_cy.pyx:21, glassy_cy.pyx:20).  Columns are population-sorted (stable
// in Beagle order inside a population) and every population slab starts on a 32-byte
// boundary (4 individuals), so one population is a contiguous, sector-aligned run of a row.
//
// None of these kernels is a contraction: there are no tensor-core instructions here.  The
// likelihood kernels are bound by FP32 issue (K >= ~10) or HBM (small K), the leave-one-out
// EM by the MUFU reciprocal rate, the per-population EM and Fisher kernels by HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wgs {

constexpr int kWarp = 32;
constexpr int kPopLikeThreads = 256;
constexpr int kPopLikeTS = 64;      // sites per shared-memory AF tile
constexpr float kLn2 = 0.6931471805599453f;

struct PopDesc {
    int col0;     // first sorted column of the slab
    int n;        // individuals in the population
};

__device__ __forceinline__ float2 ld_stream2(const float2* p) {
    // streaming 64-bit load: read once, keep it out of L1
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

// 1 - g0 - g1 with the rounding error of the first subtraction carried (the reference
// evaluates this sub-expression in double before it is multiplied, emMAF_cy.pyx:21).
__device__ __forceinline__ float third_gl(float g0, float g1) {
    float t = 1.0f - g0;
    float e = (1.0f - t) - g0;      // exact: Fast2Sum error term of t
    return (t - g1) + e;
}

// one MUFU.RCP, no denormal fix-up code (operands here are sums of likelihood terms, never denormal
// unless the likelihood itself is 0, where 0 * inf = NaN reproduces the reference's 0/0)
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Packed FP32x2 arithmetic (Blackwell FFMA2/FMUL2): two independent FP32 lanes per 64-bit
// register.  Measured on B200 (scripts/microbench/issue_rates.cu): FFMA2 issues at half the
// FFMA rate, i.e. the same FLOP/s for half the issue slots - which is what an issue-bound
// kernel needs.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(f32x2 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// Ampere-style asynchronous global->shared copies (LDGSTS): tiles are staged through a ring so
// that HBM latency is covered by bytes in flight, not by resident warps.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N)); }

// TMA bulk copies (cp.async.bulk, 1-D) completing on an mbarrier: one instruction moves a whole
// contiguous row segment global -> shared without touching registers or the LSU issue slots.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" :: "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem);
    unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(d), "l"(gmem), "r"(bytes), "r"(b) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------
// repack: staging rows in the reference layout [rows][2N] -> G[rows][ldg] (sorted columns)
// ---------------------------------------------------------------------------------------
__global__ void repack_kernel(const float2* __restrict__ stage, int N, float2* __restrict__ G, int ldg,
                              const int* __restrict__ ind_of_col, long rows)
{
    long total = rows * (long)ldg;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long r = e / ldg;
        int c = (int)(e - r * ldg);
        int src = ind_of_col[c];
        float2 v = make_float2(0.f, 0.f);
        if (src >= 0) v = stage[r * (long)N + src];
        G[e] = v;
    }
}

// G[s][pad_cols[j]] <- (0,0) for the padding columns of the population slabs: the asynchronous upload
// copies the individuals' columns straight into G (strided DMA), which leaves the pads to this kernel
__global__ void pad_fill_kernel(float2* __restrict__ G, int ldg, long M, const int* __restrict__ pad_cols, int npad)
{
    const long total = M * (long)npad;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long s = e / npad;
        const int j = (int)(e - s * npad);
        G[s * (long)ldg + pad_cols[j]] = make_float2(0.f, 0.f);
    }
}

// inverse of repack for wgs_download: G rows -> reference layout
__global__ void unpack_kernel(const float2* __restrict__ G, int ldg, const int* __restrict__ col_of_ind, int N,
                              float2* __restrict__ out, long rows)
{
    long total = rows * (long)N;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long r = e / N;
        int i = (int)(e - r * N);
        out[e] = G[r * (long)ldg + col_of_ind[i]];
    }
}

// allele depths: int32 [rows][2N] -> uchar2 [rows][ldg].  A count above 254 cannot be held: the pair is stored as the
// sentinel (255, 255) = "deeper than any class table" - such a site can never be kept by the reference either (its
// depth would need all depth+1 splits observed, zscore.py:36-39) - and counted in flags[1]; a NEGATIVE count is
// invalid input: flags[0] |= 1.
__global__ void repack_ad_kernel(const int2* __restrict__ stage, int N, uchar2* __restrict__ AD, int ldg,
                                 const int* __restrict__ ind_of_col, long rows, int* __restrict__ flags)
{
    long total = rows * (long)ldg;
    int bad = 0, sat = 0;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long r = e / ldg;
        int c = (int)(e - r * ldg);
        int src = ind_of_col[c];
        uchar2 v = make_uchar2(0, 0);
        if (src >= 0) {
            int2 a = stage[r * (long)N + src];
            if (a.x < 0 || a.y < 0) bad = 1;
            if (a.x > 254 || a.y > 254) { ++sat; v = make_uchar2(255, 255); }
            else v = make_uchar2((unsigned char)a.x, (unsigned char)a.y);
        }
        AD[e] = v;
    }
    if (bad) atomicOr(&flags[0], 1);
    if (sat) atomicAdd(&flags[1], sat);
}

// the same from saturating uint8 pairs (255 = "255 reads or more"): a count of 255 becomes the sentinel pair
__global__ void repack_ad_u8_kernel(const uchar2* __restrict__ stage, int N, uchar2* __restrict__ AD, int ldg,
                                    const int* __restrict__ ind_of_col, long rows, int* __restrict__ flags)
{
    long total = rows * (long)ldg;
    int sat = 0;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long r = e / ldg;
        int c = (int)(e - r * ldg);
        int src = ind_of_col[c];
        uchar2 v = make_uchar2(0, 0);
        if (src >= 0) {
            v = stage[r * (long)N + src];
            if (v.x == 255 || v.y == 255) { ++sat; v = make_uchar2(255, 255); }
        }
        AD[e] = v;
    }
    if (sat) atomicAdd(&flags[1], sat);
}

__global__ void unpack_ad_kernel(const uchar2* __restrict__ AD, int ldg, const int* __restrict__ col_of_ind, int N,
                                 int2* __restrict__ out, long rows)
{
    long total = rows * (long)N;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long r = e / N;
        int i = (int)(e - r * N);
        uchar2 v = AD[r * (long)ldg + col_of_ind[i]];
        out[e] = make_int2(v.x, v.y);
    }
}

// ---------------------------------------------------------------------------------------
// min over the AF matrix of min(a, 1-a): picks the renormalisation interval of pop_like
// ---------------------------------------------------------------------------------------
__global__ void af_margin_kernel(const float* __restrict__ A, long n, int* __restrict__ out_bits)
{
    float m = 0.5f;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        float a = A[e];
        float v = fminf(a, 1.0f - a);
        if (!(v > 0.f)) v = 0.f;          // NaN, <= 0 -> 0
        m = fminf(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMin(out_bits, __float_as_int(m));   // m >= 0: int order == float order
}

// Explicit shared-space accesses with 32-bit addresses.  (nvcc re-derives the shared window base - S2UR
// SR_CgaCtaId, UMOV, ULEA - around every generic-pointer store into dynamic shared memory it cannot hoist: 5 of the
// 17 instructions per element of loo_like2's cell staging.)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts128(unsigned addr, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float lds32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
// max / min that return NaN when the first operand is NaN (the second never is): `if (a < lo) a = lo` in one instruction
__device__ __forceinline__ float fmax_nan(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmin_nan(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

// ---------------------------------------------------------------------------------------
// Likelihood accumulation without a logarithm per evaluation.
//
// sum_s log(like_s) is accumulated as a running PRODUCT of the per-site likelihoods in
// FP32 whose exponent is moved into an integer every R sites (integer ALU, no MUFU): the
// sum of logs is (esum + log2(mantissa)) * ln2, evaluated once per thread in FP64.  The
// rounding error of the product is one FP32 rounding per site, unbiased - smaller than the
// error of an FP32 log per site - and the integer part is exact.  R is chosen on the host
// from the smallest allele-frequency margin so that R factors can never underflow.
// A factor that is zero / negative / NaN (the reference then yields -inf or NaN, glassy_cy.pyx:21)
// is caught at the next renormalisation and recorded in a per-population bit mask.
// ---------------------------------------------------------------------------------------
template <int KT>
struct LikeAcc {
    float prod[KT];
    int esum[KT];
    unsigned zero_mask, nan_mask;
    int nren;
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int k = 0; k < KT; ++k) { prod[k] = 1.0f; esum[k] = 0; }
        zero_mask = nan_mask = 0u; nren = 0;
    }
    __device__ __forceinline__ void renorm() {
        unsigned worst = 0u;
#pragma unroll
        for (int k = 0; k < KT; ++k) worst = max(worst, (unsigned)(__float_as_int(prod[k]) - 0x00800000));
        if (worst < 0x7F000000u) {                             // all positive normal
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                int b = __float_as_int(prod[k]);
                esum[k] += (b >> 23);
                prod[k] = __int_as_float((b & 0x007fffff) | 0x3f800000);
            }
        } else {
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                int b = __float_as_int(prod[k]);
                if ((unsigned)(b - 0x00800000) < 0x7F000000u) {
                    esum[k] += (b >> 23);
                    prod[k] = __int_as_float((b & 0x007fffff) | 0x3f800000);
                } else {
                    if (prod[k] == 0.0f || (b > 0 && b < 0x00800000)) zero_mask |= (1u << k);   // 0 or denormal
                    else nan_mask |= (1u << k);                                             // negative, inf, NaN
                    esum[k] += 127;
                    prod[k] = 1.0f;
                }
            }
        }
        ++nren;
    }
    // natural-log sum of everything accumulated so far for slot k
    __device__ __forceinline__ double value(int k) const {
        if (nan_mask & (1u << k)) return __longlong_as_double(0x7ff8000000000000LL);
        if (zero_mask & (1u << k)) return __longlong_as_double(0xfff0000000000000LL);
        double e = (double)(esum[k] - 127 * nren);
        return (e + log2((double)prod[k])) * 0.693147180559945309417232121458;
    }
};

// ---------------------------------------------------------------------------------------
// pop_like: sum over sites of log(GL . HWE(A[s,k])) for every (individual, population)
// (glassy_cy.pyx:12-21 + glassy.py:31-42).  One thread = one individual (column); a warp
// = 32 consecutive columns at one site, so a site's GL row is read with one fully
// coalesced 256-byte request per warp, 8 sites in flight per thread.  The HWE triples of
// the site tile are computed once per block into shared memory, packed by PAIRS of
// populations {(h0a,h0b),(h1a,h1b)} + (h2a,h2b), so one evaluation pair costs
// LDS.128 + LDS.64 + FMUL2 + 2 FFMA2 (likelihoods) + FMUL2 (running products): three issue
// slots per evaluation; no logarithm (see LikeAcc).  HBM-bound up to K ~ 10, FP32-pipe
// bound above.
// grid.x = column groups (fast, so that co-scheduled blocks share the AF tile in L2),
// grid.y = site splits.  partials[split][col][K] (float64) are reduced in a fixed order
// by reduce_partials_kernel - deterministic, no atomics.
// ---------------------------------------------------------------------------------------
template <int KT>
struct LikeAcc2 {                       // LikeAcc over packed pairs of populations
    static constexpr int KP = (KT + 1) / 2;
    f32x2 prod[KP];
    int esum[2 * KP];
    unsigned zero_mask, nan_mask;
    int nren;
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int q = 0; q < KP; ++q) { prod[q] = pack2(1.0f, 1.0f); esum[2 * q] = esum[2 * q + 1] = 0; }
        zero_mask = nan_mask = 0u; nren = 0;
    }
    __device__ __forceinline__ float renorm1(float p, int k) {
        int b = __float_as_int(p);
        if ((unsigned)(b - 0x00800000) < 0x7F000000u) {
            esum[k] += (b >> 23);
            return __int_as_float((b & 0x007fffff) | 0x3f800000);
        }
        if (p == 0.0f || (b > 0 && b < 0x00800000)) zero_mask |= (1u << k);
        else nan_mask |= (1u << k);
        esum[k] += 127;
        return 1.0f;
    }
    __device__ __forceinline__ void renorm() {
        // fast path: every product is a positive normal number (one unsigned max per slot decides)
        unsigned worst = 0u;
#pragma unroll
        for (int q = 0; q < KP; ++q) {
            float2 v = unpack2(prod[q]);
            worst = max(worst, (unsigned)(__float_as_int(v.x) - 0x00800000));
            worst = max(worst, (unsigned)(__float_as_int(v.y) - 0x00800000));
        }
        if (worst < 0x7F000000u) {
#pragma unroll
            for (int q = 0; q < KP; ++q) {
                float2 v = unpack2(prod[q]);
                int b0 = __float_as_int(v.x), b1 = __float_as_int(v.y);
                esum[2 * q] += (b0 >> 23);
                esum[2 * q + 1] += (b1 >> 23);
                prod[q] = pack2(__int_as_float((b0 & 0x007fffff) | 0x3f800000), __int_as_float((b1 & 0x007fffff) | 0x3f800000));
            }
        } else {                                               // rare: a zero / negative / NaN factor somewhere
#pragma unroll
            for (int q = 0; q < KP; ++q) {
                float2 v = unpack2(prod[q]);
                prod[q] = pack2(renorm1(v.x, 2 * q), renorm1(v.y, 2 * q + 1));
            }
        }
        ++nren;
    }
    __device__ __forceinline__ double value(int k) const {
        if (nan_mask & (1u << k)) return __longlong_as_double(0x7ff8000000000000LL);
        if (zero_mask & (1u << k)) return __longlong_as_double(0xfff0000000000000LL);
        float2 v = unpack2(prod[k >> 1]);
        double e = (double)(esum[k] - 127 * nren);
        return (e + log2((double)((k & 1) ? v.y : v.x))) * 0.693147180559945309417232121458;
    }
};

// One thread carries I individuals (columns col, col+32, ...): the shared-memory HWE pair is
// read once per I evaluation pairs.  A broadcast LDS still delivers 24 B to every lane, and at
// one individual per thread that delivered-byte rate (128 B/clk/SM), not issue, is the limit
// (measured: 2.1e12 evaluations/s at I = 1).
template <int KT, int R, int I, bool FULL>
__device__ __forceinline__ void pop_like_tile(const float2* const (&Gcol)[I], int ldg, long s0, long s_end,
                                              int wy, int wy_count, int per_warp,
                                              const ulonglong2* __restrict__ HA, const f32x2* __restrict__ HB,
                                              LikeAcc2<KT> (&acc)[I])
{
    constexpr int KP = (KT + 1) / 2;
    constexpr int PF = (I >= 4) ? 4 : 8;                  // sites prefetched per thread
    for (int u0 = 0; u0 < per_warp; u0 += PF) {
        float2 g[PF][I];
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            long s = s0 + wy + (long)wy_count * (u0 + u);
#pragma unroll
            for (int i = 0; i < I; ++i) {
                if (FULL || s < s_end) g[u][i] = ld_stream2(Gcol[i] + s * (long)ldg);
                else g[u][i] = make_float2(1.0f, 0.0f);
            }
        }
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int sl = wy + wy_count * (u0 + u);
            if (FULL || s0 + sl < s_end) {                    // warp-uniform
                float g0[I], g1[I], g2[I];
#pragma unroll
                for (int i = 0; i < I; ++i) { g0[i] = g[u][i].x; g1[i] = g[u][i].y; g2[i] = third_gl(g0[i], g1[i]); }
#pragma unroll
                for (int q = 0; q < KP; ++q) {
                    const ulonglong2 h = HA[sl * KP + q];
                    const f32x2 h2 = HB[sl * KP + q];
#pragma unroll
                    for (int i = 0; i < I; ++i) {
                        f32x2 like = ffma2(pack2(g0[i], g0[i]), h.x, ffma2(pack2(g1[i], g1[i]), h.y, fmul2(pack2(g2[i], g2[i]), h2)));
                        acc[i].prod[q] = fmul2(acc[i].prod[q], like);
                    }
                }
            }
            if (((u0 + u + 1) % R) == 0) {
#pragma unroll
                for (int i = 0; i < I; ++i) acc[i].renorm();
            }
        }
    }
}

template <int KT, int R, int I>
__global__ void __launch_bounds__(kPopLikeThreads)
pop_like_kernel(const float2* __restrict__ G, int ldg, long M,
                const float* __restrict__ A, int K, int k0,
                int wx,                                  // column groups (of 32*I) per block (power of two <= 8)
                long sites_per_block,
                double* __restrict__ partials)
{
    constexpr int KP = (KT + 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ulonglong2* HA = reinterpret_cast<ulonglong2*>(smem_raw);        // [TS][KP]
    f32x2* HB = reinterpret_cast<f32x2*>(HA + kPopLikeTS * KP);      // [TS][KP]
    double* red = reinterpret_cast<double*>(HB + kPopLikeTS * KP);   // [8][32][4]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy_count = (kPopLikeThreads / 32) / wx;
    const int cgx = warp % wx, wy = warp / wx;
    const int col_base = (blockIdx.x * wx + cgx) * 32 * I + lane;
    const float2* Gcol[I];
    bool col_ok[I];
#pragma unroll
    for (int i = 0; i < I; ++i) {
        int c = col_base + 32 * i;
        col_ok[i] = c < ldg;
        Gcol[i] = G + (col_ok[i] ? c : ldg - 1);         // out-of-range slots read a valid column and are never stored
    }
    const long s_begin = (long)blockIdx.y * sites_per_block;
    const long s_end = min(M, s_begin + sites_per_block);

    LikeAcc2<KT> acc[I];
#pragma unroll
    for (int i = 0; i < I; ++i) acc[i].init();

    const int per_warp = kPopLikeTS / wy_count;       // sites of a tile handled by one warp (multiple of 8)
    for (long s0 = s_begin; s0 < s_end; s0 += kPopLikeTS) {
        __syncthreads();
        for (int e = threadIdx.x; e < kPopLikeTS * KP; e += kPopLikeThreads) {
            int sl = e / KP, q = e - sl * KP;
            long s = s0 + sl;
            float hh[2][3];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int k = k0 + 2 * q + h;
                hh[h][0] = hh[h][1] = hh[h][2] = 1.0f;      // dummy slot: likelihood 1
                if (s < s_end && 2 * q + h < KT && k < K) {
                    float a = __ldg(&A[s * K + k]);
                    float om = 1.0f - a;
                    hh[h][0] = om * om; hh[h][1] = 2.0f * a * om; hh[h][2] = a * a;
                }
            }
            ulonglong2 v;
            v.x = pack2(hh[0][0], hh[1][0]); v.y = pack2(hh[0][1], hh[1][1]);
            HA[e] = v;
            HB[e] = pack2(hh[0][2], hh[1][2]);
        }
        __syncthreads();
        if (col_base - lane >= ldg) continue;             // the whole warp is past the last column (warp-uniform)
        if (s0 + kPopLikeTS <= s_end) pop_like_tile<KT, R, I, true>(Gcol, ldg, s0, s_end, wy, wy_count, per_warp, HA, HB, acc);
        else pop_like_tile<KT, R, I, false>(Gcol, ldg, s0, s_end, wy, wy_count, per_warp, HA, HB, acc);
    }
#pragma unroll
    for (int i = 0; i < I; ++i) acc[i].renorm();   // brings every slot to a known state (also folds a trailing partial group)

    // fixed-order reduction over the wy_count warps that share this column group
#pragma unroll
    for (int i = 0; i < I; ++i) {
#pragma unroll
        for (int kk0 = 0; kk0 < KT; kk0 += 4) {
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (kk0 + q < KT) red[(warp * 32 + lane) * 4 + q] = acc[i].value(kk0 + q);
            __syncthreads();
            if (wy == 0 && col_ok[i]) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int kk = kk0 + q;
                    if (kk < KT && k0 + kk < K) {
                        double v = 0.0;
                        for (int w = 0; w < wy_count; ++w) v += red[((w * wx + cgx) * 32 + lane) * 4 + q];
                        partials[((long)blockIdx.y * ldg + col_base + 32 * i) * K + k0 + kk] = v;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// pop_like2: the same sums as pop_like in RATIO form - the kernel the drivers use whenever
// every allele frequency lies inside [2^-30, 1-2^-30] (always, after the reference's clipping).
//
//   log(g0 (1-a)^2 + g1 2a(1-a) + g2 a^2) = log(2a(1-a)) + log(g0 x + g1 + g2 y),
//   x = (1-a)/(2a),  y = a/(2(1-a))
//
// The first term does not depend on the individual: af_logsum_kernel sums it once per
// population (exactly, same mantissa-product scheme).  What is left per evaluation is two
// multiply-adds and the running-product multiply (was three + one), and the per-(site,
// population pair) coefficients shrink from 24 to 16 bytes: ONE broadcast LDS.128 per two
// evaluations, shared by the I individuals a thread carries.  Every per-site factor lies in
// [m/2, 1/(2m)] (m = smallest min(a, 1-a)), a range symmetric in log scale and half as wide
// as the [m^2, 1] of the direct form, so twice as many sites fit between renormalisations
// (R = 4, 8 or 16, picked on the host from m).
//
// Exponent sums of a population PAIR share one register (two 16-bit fields, at most 255 per
// renormalisation): a thread may renormalise at most 256 times, i.e. walks at most 256 R sites,
// which the launch geometry guarantees (sites_per_block <= 256 R).
//
// Block = wx warps, each a different group of 32 I columns, all walking the same sites, so the
// coefficient tile is built once per wx 32 I individuals.  partials[split][col][K] as pop_like.
// ---------------------------------------------------------------------------------------
template <int KT>
struct LikeAccP {
    static constexpr int KP = (KT + 1) / 2;
    f32x2 prod[KP];
    unsigned esum[KP];                       // lo 16 bits: slot 2q, hi 16 bits: slot 2q+1
    unsigned zero_mask, nan_mask;
    int nren;
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int q = 0; q < KP; ++q) { prod[q] = pack2(1.0f, 1.0f); esum[q] = 0u; }
        zero_mask = nan_mask = 0u; nren = 0;
    }
    __device__ __forceinline__ float slow1(float p, int k, unsigned& field) {
        const int b = __float_as_int(p);
        if ((unsigned)(b - 0x00800000) < 0x7F000000u) { field = (unsigned)(b >> 23); return __int_as_float((b & 0x007fffff) | 0x3f800000); }
        if (p == 0.0f || (b > 0 && b < 0x00800000)) zero_mask |= (1u << k);      // 0 or denormal
        else nan_mask |= (1u << k);                                          // negative, inf, NaN
        field = 127u;
        return 1.0f;
    }
    __device__ __forceinline__ void renorm() {
        unsigned worst = 0u;
#pragma unroll
        for (int q = 0; q < KP; ++q) {
            const float2 v = unpack2(prod[q]);
            worst = max(worst, max((unsigned)(__float_as_int(v.x) - 0x00800000), (unsigned)(__float_as_int(v.y) - 0x00800000)));
        }
        if (worst < 0x7F000000u) {                             // every product is a positive normal number
#pragma unroll
            for (int q = 0; q < KP; ++q) {
                const float2 v = unpack2(prod[q]);
                const unsigned b0 = __float_as_uint(v.x), b1 = __float_as_uint(v.y);
                esum[q] += (b0 >> 23) + ((b1 >> 7) & 0xFFFF0000u);
                prod[q] = pack2(__uint_as_float((b0 & 0x007fffffu) | 0x3f800000u), __uint_as_float((b1 & 0x007fffffu) | 0x3f800000u));
            }
        } else {                                               // rare: a zero / negative / NaN factor somewhere
#pragma unroll
            for (int q = 0; q < KP; ++q) {
                const float2 v = unpack2(prod[q]);
                unsigned f0, f1;
                const float p0 = slow1(v.x, 2 * q, f0), p1 = slow1(v.y, 2 * q + 1, f1);
                esum[q] += f0 + (f1 << 16);
                prod[q] = pack2(p0, p1);
            }
        }
        ++nren;
    }
    __device__ __forceinline__ double value(int k) const {     // natural-log sum of slot k
        if (nan_mask & (1u << k)) return __longlong_as_double(0x7ff8000000000000LL);
        if (zero_mask & (1u << k)) return __longlong_as_double(0xfff0000000000000LL);
        const float2 v = unpack2(prod[k >> 1]);
        const unsigned field = (k & 1) ? (esum[k >> 1] >> 16) : (esum[k >> 1] & 0xFFFFu);
        const double e = (double)((int)field - 127 * nren);
        return (e + log2((double)((k & 1) ? v.y : v.x))) * 0.693147180559945309417232121458;
    }
};

constexpr int kPL2TS = 32;       // sites per coefficient tile (one warp-private tile)
constexpr int kPL2PF = 8;        // sites prefetched per thread

// XY[s][q] = {(x_a, x_b), (y_a, y_b)} for the population pair (k0+2q, k0+2q+1) of site s, computed
// once per pass (16 bytes per two (site, population) cells); slots past K are (1/2, 1/2).
__global__ void xy_precompute_kernel(const float* __restrict__ A, long M, int K, int k0, int KP, ulonglong2* __restrict__ XY)
{
    const long total = M * KP;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long s = e / KP;
        const int q = (int)(e - s * KP);
        float x[2], y[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = k0 + 2 * q + h;
            x[h] = 0.5f; y[h] = 0.5f;                     // dummy slot: factor g0/2 + g1 + g2/2, never stored
            if (k < K) {
                const float a = __ldg(&A[s * K + k]);
                const float om = 1.0f - a;
                x[h] = om * fast_rcp(a + a);
                y[h] = a * fast_rcp(om + om);
            }
        }
        ulonglong2 v;
        v.x = pack2(x[0], x[1]); v.y = pack2(y[0], y[1]);
        XY[e] = v;
    }
}

template <int KT, int I, int R, bool FULL>
__device__ __forceinline__ void pop_like2_tile(const float2* const (&Gcol)[I], int ldg, long s0, long s_end,
                                               const ulonglong2* __restrict__ XY, LikeAccP<KT> (&acc)[I])
{
    constexpr int KP = (KT + 1) / 2;
    constexpr int CH = (R > kPL2PF) ? R : kPL2PF;              // sites per unrolled chunk
    static_assert(kPL2TS % CH == 0, "tile must hold whole chunks");
#pragma unroll 1
    for (int c0 = 0; c0 < kPL2TS; c0 += CH) {
#pragma unroll
        for (int u0 = 0; u0 < CH; u0 += kPL2PF) {
            float2 g[kPL2PF][I];
            const float2* base[I];
#pragma unroll
            for (int i = 0; i < I; ++i) base[i] = Gcol[i] + (s0 + c0 + u0) * (long)ldg;
#pragma unroll
            for (int u = 0; u < kPL2PF; ++u) {
                const long s = s0 + c0 + u0 + u;
#pragma unroll
                for (int i = 0; i < I; ++i) {
                    if (FULL || s < s_end) g[u][i] = ld_stream2(base[i] + u * ldg);
                    else g[u][i] = make_float2(1.0f, 0.0f);
                }
            }
#pragma unroll
            for (int u = 0; u < kPL2PF; ++u) {
                const int sl = c0 + u0 + u;
                if (FULL || s0 + sl < s_end) {                 // warp-uniform
                    f32x2 g0[I], g1[I], g2[I];
#pragma unroll
                    for (int i = 0; i < I; ++i) {
                        const float t = third_gl(g[u][i].x, g[u][i].y);
                        g0[i] = pack2(g[u][i].x, g[u][i].x); g1[i] = pack2(g[u][i].y, g[u][i].y); g2[i] = pack2(t, t);
                    }
#pragma unroll
                    for (int q = 0; q < KP; ++q) {
                        const ulonglong2 xy = XY[sl * KP + q];
#pragma unroll
                        for (int i = 0; i < I; ++i)
                            acc[i].prod[q] = fmul2(acc[i].prod[q], ffma2(g2[i], xy.y, ffma2(g0[i], xy.x, g1[i])));
                    }
                }
                if (((u0 + u + 1) % R) == 0) {
#pragma unroll
                    for (int i = 0; i < I; ++i) acc[i].renorm();
                }
            }
        }
    }
}

// One WARP = one work unit (32 I columns x one site split): its coefficient tiles are private
// (double-buffered LDGSTS copies of the precomputed XY rows), so there is no block-wide barrier
// anywhere and warps drift freely - the block-shared tile with two barriers per 64 sites cost
// 19 % at 9 warps per block.  A block is just up to 4 neighbouring column groups.
template <int KT, int I, int R>
__global__ void __launch_bounds__(128, 5)
pop_like2_kernel(const float2* __restrict__ G, int ldg, long M,
                 const ulonglong2* __restrict__ XYg,      // [M rounded up to kPL2TS][KP]
                 int K, int k0,
                 long sites_per_block,                    // multiple of kPL2TS, <= 256 R
                 double* __restrict__ partials)
{
    constexpr int KP = (KT + 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ulonglong2* XY = reinterpret_cast<ulonglong2*>(smem_raw) + (size_t)warp * 2 * kPL2TS * KP;   // [2][TS][KP], this warp's

    const int wx = blockDim.x >> 5;
    const int col_base = (blockIdx.x * wx + warp) * 32 * I + lane;
    if (col_base - lane >= ldg) return;                   // warp-uniform: nothing to do (no barriers below)
    const float2* Gcol[I];
    bool col_ok[I];
#pragma unroll
    for (int i = 0; i < I; ++i) {
        const int c = col_base + 32 * i;
        col_ok[i] = c < ldg;
        Gcol[i] = G + (col_ok[i] ? c : ldg - 1);         // out-of-range slots read a valid column and are never stored
    }
    const long s_begin = (long)blockIdx.y * sites_per_block;
    const long s_end = min(M, s_begin + sites_per_block);

    LikeAccP<KT> acc[I];
#pragma unroll
    for (int i = 0; i < I; ++i) acc[i].init();

    auto stage = [&](long s0, int buf) {                  // 32 rows of KP 16-byte cells: contiguous in XYg
        const ulonglong2* src = XYg + s0 * KP;
        ulonglong2* dst = XY + buf * (kPL2TS * KP);
#pragma unroll
        for (int j = 0; j < KP; ++j) cp_async16(dst + lane + 32 * j, src + lane + 32 * j);
        cp_async_commit();
    };
    if (s_begin < s_end) stage(s_begin, 0);
    int buf = 0;
    for (long s0 = s_begin; s0 < s_end; s0 += kPL2TS, buf ^= 1) {
        if (s0 + kPL2TS < s_end) { stage(s0 + kPL2TS, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncwarp();
        const ulonglong2* tile = XY + buf * (kPL2TS * KP);
        if (s0 + kPL2TS <= s_end) pop_like2_tile<KT, I, R, true>(Gcol, ldg, s0, s_end, tile, acc);
        else pop_like2_tile<KT, I, R, false>(Gcol, ldg, s0, s_end, tile, acc);
        __syncwarp();                                     // everyone is done with `buf` before it is refilled
    }
#pragma unroll
    for (int i = 0; i < I; ++i) {
        acc[i].renorm();                                 // brings every slot to a known state
        if (col_ok[i]) {
#pragma unroll
            for (int kk = 0; kk < KT; ++kk)
                if (k0 + kk < K) partials[((long)blockIdx.y * ldg + col_base + 32 * i) * K + k0 + kk] = acc[i].value(kk);
        }
    }
}

// per_block[b][k] = sum over the elements block b visits of log(2 a_e (1 - a_e)) for population k.
// T = total threads, a multiple of K (blockDim is): every element a thread visits (e = t, t+T, ...) belongs to
// population t % K.  Block and grid sums are taken in a fixed order (af_logsum_reduce_kernel): deterministic.
// (The population-only term of the ratio form above.)
__global__ void af_logsum_kernel(const float* __restrict__ A, long n, int K, double* __restrict__ per_block,
                                 long part_mod = 1, long part_rem = 0, long site_offset = 0)   // only sites with (site_offset + s) % part_mod == part_rem
{
    extern __shared__ double af_red[];                   // [blockDim.x]
    const long T = (long)gridDim.x * blockDim.x, t = blockIdx.x * (long)blockDim.x + threadIdx.x;
    LikeAcc<1> acc;
    acc.init();
    int cnt = 0;
    for (long e = t; e < n; e += T) {
        if (part_mod > 1 && (site_offset + e / K) % part_mod != part_rem) continue;
        const float a = __ldg(&A[e]);
        acc.prod[0] *= (a + a) * (1.0f - a);
        if (++cnt == 2) { acc.renorm(); cnt = 0; }        // two factors >= 2^-29 each stay far above underflow
    }
    acc.renorm();
    af_red[threadIdx.x] = acc.value(0);
    __syncthreads();
    if (threadIdx.x < K) {
        double v = 0.0;
        for (int j = threadIdx.x; j < (int)blockDim.x; j += K) v += af_red[j];
        per_block[(long)blockIdx.x * K + threadIdx.x] = v;
    }
}
__global__ void af_logsum_reduce_kernel(const double* __restrict__ per_block, int nblocks, int K, double* __restrict__ C)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double v = 0.0;
    for (int b = 0; b < nblocks; ++b) v += per_block[(long)b * K + k];
    C[k] = v;
}
// sums[col][k] += C[k]
__global__ void add_pop_const_kernel(double* __restrict__ sums, long n, int K, const double* __restrict__ C)
{
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) sums[e] += C[e % K];
}

// out[col][k] = sum over splits (fixed order) of partials[split][col][k]
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int nsplit, long n, double* __restrict__ out)
{
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        double v = 0.0;
        for (int sp = 0; sp < nsplit; ++sp) v += partials[(long)sp * n + e];
        out[e] = v;
    }
}

// ---------------------------------------------------------------------------------------
// loo_like: like pop_like, but the allele frequency of (individual c, population j) is a
// per-pair COLUMN of the leave-one-out state matrix Fx[M][ldf]: the reference overwrites
// af[:, pop(i)] with individual i's LOO estimate and never restores it (glassy.py:89), so
// population j's column seen by individual i is the LOO estimate of the latest earlier
// member of j (or the full-data AF, stored in the last K columns of Fx).  rc[col][K] holds
// that column index.  Same accumulation/reduction scheme as pop_like.
// ---------------------------------------------------------------------------------------
template <int KT, int R>
__global__ void __launch_bounds__(kPopLikeThreads)
loo_like_kernel(const float2* __restrict__ G, int ldg, long M,
                const float* __restrict__ Fx, int ldf,
                const int* __restrict__ rc, int K, int k0,
                int wx, long sites_per_block,
                long part_mod, long part_rem, long site_offset,
                double* __restrict__ partials)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* red = reinterpret_cast<double*>(smem_raw);    // [8][32][4]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy_count = (kPopLikeThreads / 32) / wx;
    const int cgx = warp % wx, wy = warp / wx;
    const int col = (blockIdx.x * wx + cgx) * 32 + lane;
    const bool col_ok = col < ldg;
    const long s_begin = (long)blockIdx.y * sites_per_block;
    const long s_end = min(M, s_begin + sites_per_block);

    int rcol[KT];
#pragma unroll
    for (int kk = 0; kk < KT; ++kk)
        rcol[kk] = (col_ok && k0 + kk < K) ? rc[(long)col * K + k0 + kk] : (ldf - 1);

    LikeAcc<KT> acc;
    acc.init();
    int cnt = 0;
    constexpr int PF = (KT > 10) ? 2 : 4;                  // sites in flight per thread (GL + KT state gathers each)
    const float2* Gc = G + (col_ok ? col : ldg - 1);
    const long s_stop = (col - lane >= ldg) ? s_begin : s_end;      // a warp entirely past the last column does nothing
    for (long sb = s_begin + wy; sb < s_stop; sb += (long)wy_count * PF) {
        float2 g[PF];
        float a[PF][KT];
        bool use[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            long s = sb + (long)wy_count * u;
            use[u] = s < s_end;
            if (part_mod > 1) use[u] = use[u] && ((site_offset + s) % part_mod == part_rem);
            g[u] = make_float2(1.0f, 0.0f);
            if (use[u]) {
                g[u] = ld_stream2(Gc + s * (long)ldg);
                const float* frow = Fx + s * (long)ldf;
#pragma unroll
                for (int kk = 0; kk < KT; ++kk) a[u][kk] = __ldg(&frow[rcol[kk]]);
            }
        }
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            if (use[u]) {                                     // warp-uniform
                float g0 = g[u].x, g1 = g[u].y, g2 = third_gl(g0, g1);
#pragma unroll
                for (int kk = 0; kk < KT; ++kk) {
                    float av = a[u][kk];
                    float om = 1.0f - av;
                    float like = fmaf(g0, om * om, fmaf(g1, 2.0f * av * om, g2 * (av * av)));
                    acc.prod[kk] *= like;
                }
                if (++cnt == R) { acc.renorm(); cnt = 0; }
            }
        }
    }
    acc.renorm();

#pragma unroll
    for (int kk0 = 0; kk0 < KT; kk0 += 4) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (kk0 + q < KT) red[(warp * 32 + lane) * 4 + q] = acc.value(kk0 + q);
        __syncthreads();
        if (wy == 0 && col_ok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int kk = kk0 + q;
                if (kk < KT && k0 + kk < K) {
                    double v = 0.0;
                    for (int w = 0; w < wy_count; ++w) v += red[((w * wx + cgx) * 32 + lane) * 4 + q];
                    partials[((long)blockIdx.y * ldg + col) * K + k0 + kk] = v;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// loo_like2: loo_like with the leave-one-out state staged in shared memory.  The gather of
// Fx[s][rc[c][j]] through L1 kept loo_like latency-bound (11 dependent global loads per site and
// thread).  Here a block walks a range of sites with ALL its warps on the same site tile: the TS
// state rows of the tile arrive by one TMA bulk copy (rows of consecutive sites are contiguous),
// are turned once into 16-byte HWE cells ((1-a)^2, 2a(1-a), a^2, -), and every (individual,
// population) evaluation is then one address add + LDS.128 + 3 multiply-adds + the running-product multiply.
// The block's GL columns of the tile arrive the same way (one bulk copy per site row, double-buffered,
// issued one tile ahead) so HBM latency hides behind the previous tile's arithmetic.
// Warps own 32 columns each and all sites of the block's split: no cross-warp reduction.
// ---------------------------------------------------------------------------------------
// Up to 10 warps per block.  KT <= 10: two blocks per SM (<= 96 registers), tiles of up to 8 sites; wider
// population tiles: one block per SM with up to 4 sites per tile (the 2 x KT accumulators need the registers).
// A third shape, <KT, 8, 1, kLL2BigW>: when the individuals fill 11..18 warps ONE block of up to 18 warps per SM takes
// all columns of its site split, so the cells of a state row are built once per tile instead of once per block.
constexpr int kLL2MaxW = 10;
constexpr int kLL2BigW = 18;
constexpr int kLL2WideW = 16;   // wide population tiles (K > 10): 16 warps at 128 registers
template <int KT, int TSMAX, int MINB, int MAXW = kLL2MaxW>
__global__ void __launch_bounds__(MAXW * 32, MINB)
loo_like2_kernel(const float2* __restrict__ G, int ldg, long M,
                 const float* __restrict__ Fx, int ldf,
                 const float* __restrict__ clip_lo, const float* __restrict__ clip_hi,   // [ldf] or null: clamp applied while staging
                 const int* __restrict__ rc, int K, int k0,
                 int TS, long sites_per_block,                  // TS <= TSMAX; sites_per_block a multiple of TS
                 long part_mod, long part_rem, long site_offset, int R,
                 double* __restrict__ partials)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long mbar, mbar_g[2];
    const int t = threadIdx.x, lane = t & 31;
    const int W = blockDim.x >> 5;
    const int wcols = W * 32;                                               // GL columns of this block
    float4* Hq = reinterpret_cast<float4*>(smem_raw);                       // [TS][ldf] planes ((1-a)^2, 2a(1-a), a^2, -)
    float2* Gs = reinterpret_cast<float2*>(Hq + (size_t)TS * ldf);          // [2][TS][wcols] GL pairs, double-buffered
    float* raw = reinterpret_cast<float*>(Gs + 2 * (size_t)TS * wcols);     // [TS][ldf] landing rows of the next tile's state
    float* clip_s = raw + (size_t)TS * ldf;                                 // [2][ldf] clamp bounds per state column
    const int col0 = blockIdx.x * wcols;
    const int col = col0 + t;
    const bool col_ok = col < ldg;
    const bool warp_live = col - lane < ldg;
    const int lcol = min(col, ldg - 1) - col0;                              // a valid column for the padding lanes of a live warp
    const long s_begin = (long)blockIdx.y * sites_per_block;
    const long s_end = min(M, s_begin + sites_per_block);
    const int ntiles = (int)((s_end - s_begin + TS - 1) / TS);

    unsigned roff[KT];                                    // byte offset of this thread's state cell inside a plane row
#pragma unroll
    for (int kk = 0; kk < KT; ++kk)
        roff[kk] = 16u * (unsigned)((col_ok && k0 + kk < K) ? rc[(long)col * K + k0 + kk] : (ldf - 1));

    for (int c = t; c < ldf; c += blockDim.x) {           // without bounds: (-inf, +inf) leaves every value, NaN included, as it is
        clip_s[c] = clip_lo ? clip_lo[c] : -INFINITY;
        clip_s[ldf + c] = clip_lo ? clip_hi[c] : INFINITY;
    }
    if (t == 0) { mbar_init(&mbar, 1); mbar_init(&mbar_g[0], 1); mbar_init(&mbar_g[1], 1); mbar_fence_init(); }
    __syncthreads();
    auto issue_state = [&](int j) {                       // thread 0: one bulk copy for the tile's state rows
        if (j < ntiles && t == 0) {
            const long s0 = s_begin + (long)j * TS;
            const unsigned bytes = (unsigned)min((long)TS, s_end - s0) * (unsigned)ldf * 4u;
            mbar_expect_tx(&mbar, bytes);
            bulk_g2s(raw, Fx + s0 * (long)ldf, bytes, &mbar);
        }
    };
    auto issue_gl = [&](int j) {                          // thread 0: one bulk copy per site row of the block's GL columns
        if (j < ntiles && t == 0) {
            const int buf = j & 1;
            const long s0 = s_begin + (long)j * TS;
            const int rows = (int)min((long)TS, s_end - s0);
            const unsigned wbytes = (unsigned)min(wcols, ldg - col0) * 8u;  // a multiple of 32: slabs are padded to 4 individuals
            mbar_expect_tx(&mbar_g[buf], wbytes * (unsigned)rows);
            for (int u = 0; u < rows; ++u)
                bulk_g2s(Gs + ((size_t)buf * TS + u) * wcols, G + (s0 + u) * (long)ldg + col0, wbytes, &mbar_g[buf]);
        }
    };
    issue_state(0);
    issue_gl(0);

    LikeAcc<KT> acc;
    acc.init();
    int cnt = 0;
    const unsigned row_bytes = 16u * (unsigned)ldf, row_bytes_raw = 4u * (unsigned)ldf;
    const unsigned hq_a = smem_u32(Hq), raw_a = smem_u32(raw);
    for (int j = 0; j < ntiles; ++j) {
        const long s0 = s_begin + (long)j * TS;
        const int rows = (int)min((long)TS, s_end - s0);
        issue_gl(j + 1);                                  // its buffer was released by the barrier that ended tile j-1
        mbar_wait(&mbar, (unsigned)(j & 1));
        // the clipping of glassy.py:80-85 happens here, on the way into the cells: the state matrix is never
        // rewritten (a separate clamp pass over it was 1.4 ms at 1M x 500)
        for (int c = t; c < ldf; c += blockDim.x) {       // a thread takes whole columns: its bounds are loaded once per tile
            const float lo = clip_s[c], hi = clip_s[ldf + c];
            const unsigned ra = raw_a + 4u * (unsigned)c, ha = hq_a + 16u * (unsigned)c;
            float av[TSMAX];
#pragma unroll
            for (int u = 0; u < TSMAX; ++u) av[u] = u < rows ? lds32(ra + (unsigned)u * row_bytes_raw) : 0.5f;
#pragma unroll
            for (int u = 0; u < TSMAX; ++u) {
                if (u < rows) {
                    const float a = fmin_nan(fmax_nan(av[u], lo), hi);      // NaN survives, like the reference's compare-and-assign
                    const float om = 1.0f - a;
                    sts128(ha + (unsigned)u * row_bytes, om * om, (a + a) * om, a * a, 0.0f);
                }
            }
        }
        __syncthreads();                                  // planes complete, landing rows free: the next tile's copy overlaps the compute
        issue_state(j + 1);
        mbar_wait(&mbar_g[j & 1], (unsigned)((j >> 1) & 1));
        // partition of the tile's first site (site indices fit 32 bits: the reference's sizes are C ints)
        const unsigned tile_rem = part_mod > 1 ? (unsigned)(site_offset + s0) % (unsigned)part_mod : 0u;
        if (warp_live) {
            const unsigned char* hrow = reinterpret_cast<const unsigned char*>(Hq);
            const float2* grow = Gs + (size_t)(j & 1) * TS * wcols + lcol;
#pragma unroll
            for (int u = 0; u < TSMAX; ++u) {
                if (u < rows) {                           // block-uniform
                    const bool use = part_mod <= 1 || (tile_rem + (unsigned)u) % (unsigned)part_mod == (unsigned)part_rem;
                    if (use) {
                        const float2 gq = grow[(size_t)u * wcols];
                        const float g0 = gq.x, g1 = gq.y, g2 = third_gl(g0, g1);
#pragma unroll
                        for (int kk = 0; kk < KT; ++kk) {
                            // 8 + 4 bytes: 3 shared-memory wavefronts per warp instead of the 4 of one LDS.128
                            const float2 h = *reinterpret_cast<const float2*>(hrow + roff[kk]);
                            const float hz = *reinterpret_cast<const float*>(hrow + roff[kk] + 8);
                            const float like = fmaf(g0, h.x, fmaf(g1, h.y, g2 * hz));
                            acc.prod[kk] *= like;
                        }
                        if (++cnt == R) { acc.renorm(); cnt = 0; }
                    }
                }
                hrow += row_bytes;
            }
        }
        __syncthreads();                                  // everyone is done with the planes and this tile's GL buffer
    }
    acc.renorm();
    if (col_ok) {
#pragma unroll
        for (int kk = 0; kk < KT; ++kk)
            if (k0 + kk < K) partials[((long)blockIdx.y * ldg + col) * K + k0 + kk] = acc.value(kk);
    }
}

// ---------------------------------------------------------------------------------------
// loo_like3: the leave-one-out likelihoods for reference panels whose populations are CONTIGUOUS runs of the ID
// file (the usual layout; anything else takes loo_like2).  The reference overwrites af[:, pop(i)] with individual
// i's leave-one-out estimate and never restores it (glassy.py:89), so for a run-ordered panel the column that
// individual i of population p reads for population j is
//     j == p            : i's own leave-one-out estimate,
//     run(j) before p   : the estimate of the LAST member of j (the file has passed all of j),
//     run(j) after p    : the caller's full-data frequency of j (nobody of j has been seen yet).
// Per site that is 2K shared values and one private value per individual - not an ldf-wide gather table.  The 2K
// shared values are turned once per call into ratio-form coefficient pairs XY2[s][{last, full}][pair] (the
// pop_like2 scheme, slots in RUN order, their log(2a(1-a)) sums added afterwards per individual), staged per warp
// with LDGSTS into a private double buffer; a lane picks `last` or `full` per pair with a per-thread offset (two
// distinct shared-memory addresses per warp at most).  The own term is the direct form on the individual's own
// clipped state value, read coalesced next to the GL pair, with its own running product.
// No block-wide barrier; 12 bytes per (site, individual) from HBM + 32 KP bytes per site.
// partials[split][col][K] in POPULATION order (slot -> pop_of_slot); the own slot carries the own-term sum.
// ---------------------------------------------------------------------------------------
template <int KT, int R, bool FULL>
__device__ __forceinline__ void loo_like3_tile(const float4* Gp, const float2* Fp, int ldg2, int ldf2,
                                               float lo0, float hi0, float lo1, float hi1, long s0, long s_end,
                                               long part_mod, long part_rem, long site_offset, int r_own,
                                               const ulonglong2* __restrict__ XY, const int (&qoff)[(KT + 1) / 2],
                                               LikeAccP<KT> (&acc)[2], LikeAcc<1> (&own)[2], int& cnt, int& cnt_own)
{
    constexpr int KP = (KT + 1) / 2;
    constexpr int PF = KT > 10 ? 4 : 8;                   // sites prefetched per thread (registers)
#pragma unroll 1
    for (int c0 = 0; c0 < kPL2TS; c0 += PF) {
        float4 g[PF];
        float2 fo[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const long s = s0 + c0 + u;
            if (FULL || s < s_end) { g[u] = ld_stream4(Gp + s * (long)ldg2); fo[u] = __ldg(Fp + s * (long)ldf2); }
            else { g[u] = make_float4(1.0f, 0.0f, 1.0f, 0.0f); fo[u] = make_float2(0.5f, 0.5f); }
        }
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int sl = c0 + u;
            bool use = FULL || s0 + sl < s_end;           // warp-uniform
            if (part_mod > 1) use = use && ((site_offset + s0 + sl) % part_mod == part_rem);
            if (use) {
                const float ta = third_gl(g[u].x, g[u].y), tb = third_gl(g[u].z, g[u].w);
                const f32x2 g0a = pack2(g[u].x, g[u].x), g1a = pack2(g[u].y, g[u].y), g2a = pack2(ta, ta);
                const f32x2 g0b = pack2(g[u].z, g[u].z), g1b = pack2(g[u].w, g[u].w), g2b = pack2(tb, tb);
                {   // own terms (direct form): the clipping of glassy.py:80-85 happens here, NaN survives it
                    const float a = fmin_nan(fmax_nan(fo[u].x, lo0), hi0), oa = 1.0f - a;
                    const float b = fmin_nan(fmax_nan(fo[u].y, lo1), hi1), ob = 1.0f - b;
                    own[0].prod[0] *= fmaf(g[u].x, oa * oa, fmaf(g[u].y, (a + a) * oa, ta * (a * a)));
                    own[1].prod[0] *= fmaf(g[u].z, ob * ob, fmaf(g[u].w, (b + b) * ob, tb * (b * b)));
                }
#pragma unroll
                for (int q = 0; q < KP; ++q) {
                    const ulonglong2 xy = XY[sl * (2 * KP) + qoff[q]];
                    acc[0].prod[q] = fmul2(acc[0].prod[q], ffma2(g2a, xy.y, ffma2(g0a, xy.x, g1a)));
                    acc[1].prod[q] = fmul2(acc[1].prod[q], ffma2(g2b, xy.y, ffma2(g0b, xy.x, g1b)));
                }
                if (++cnt == R) { cnt = 0; acc[0].renorm(); acc[1].renorm(); }
                if (++cnt_own == r_own) { cnt_own = 0; own[0].renorm(); own[1].renorm(); }
            }
        }
    }
}

// Thread = TWO ADJACENT columns (2 lane, 2 lane + 1 of the warp's 64): slabs are padded to 4 columns, so the two always
// belong to the same population and share every coefficient load; their GL pairs are one 16-byte load, their state
// values one 8-byte load.
template <int KT, int R>
__global__ void __launch_bounds__(128, 4)
loo_like3_kernel(const float2* __restrict__ G, int ldg, long M,
                 const float* __restrict__ F, int ldf,          // leave-one-out state: column c = the estimate without individual c
                 const float* __restrict__ clip_lo, const float* __restrict__ clip_hi,   // [ldf]
                 const ulonglong2* __restrict__ XYg,            // [M rounded up to kPL2TS][2 KP]: {last pairs | full pairs}, slots in run order
                 const int* __restrict__ slot_of_col,           // [ldg] run-order slot of the column's population
                 const int* __restrict__ pop_of_slot,           // [K]
                 int K, long sites_per_block,                   // multiple of kPL2TS, <= 256 R
                 long part_mod, long part_rem, long site_offset, int r_own,
                 double* __restrict__ partials)
{
    constexpr int KP = (KT + 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ulonglong2* XY = reinterpret_cast<ulonglong2*>(smem_raw) + (size_t)warp * 2 * kPL2TS * 2 * KP;   // [2][TS][2 KP], this warp's

    const int wx = blockDim.x >> 5;
    const int wcol = (blockIdx.x * wx + warp) * 64;
    if (wcol >= ldg) return;                              // warp-uniform: nothing to do (no barriers below)
    const int col = wcol + 2 * lane;
    const bool col_ok = col < ldg;                        // ldg is a multiple of 4: col + 1 < ldg as well
    const int cc = col_ok ? col : ldg - 2;                // out-of-range lanes read a valid pair and are never stored
    const float4* Gp = reinterpret_cast<const float4*>(G + cc);
    const float2* Fp = reinterpret_cast<const float2*>(F + cc);
    const float lo0 = clip_lo[cc], hi0 = clip_hi[cc], lo1 = clip_lo[cc + 1], hi1 = clip_hi[cc + 1];
    const int slot = slot_of_col[cc];
    // pair q (slots 2q, 2q+1) comes from the `last` set when its slots lie at or before the own slot (the own slot's
    // value is never used), else from the `full` set
    int qoff[KP];
#pragma unroll
    for (int q = 0; q < KP; ++q) qoff[q] = (2 * q + 1 <= slot) ? q : KP + q;
    const long s_begin = (long)blockIdx.y * sites_per_block;
    const long s_end = min(M, s_begin + sites_per_block);

    LikeAccP<KT> acc[2];
    LikeAcc<1> own[2];
    acc[0].init(); acc[1].init(); own[0].init(); own[1].init();
    int cnt = 0, cnt_own = 0;

    auto stage = [&](long s0, int buf) {                  // 32 rows of 2 KP 16-byte cells: contiguous in XYg
        const ulonglong2* src = XYg + s0 * (2 * KP);
        ulonglong2* dst = XY + buf * (kPL2TS * 2 * KP);
#pragma unroll
        for (int j = 0; j < 2 * KP; ++j) cp_async16(dst + lane + 32 * j, src + lane + 32 * j);
        cp_async_commit();
    };
    if (s_begin < s_end) stage(s_begin, 0);
    int buf = 0;
    for (long s0 = s_begin; s0 < s_end; s0 += kPL2TS, buf ^= 1) {
        if (s0 + kPL2TS < s_end) { stage(s0 + kPL2TS, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncwarp();
        const ulonglong2* tile = XY + buf * (kPL2TS * 2 * KP);
        if (s0 + kPL2TS <= s_end)
            loo_like3_tile<KT, R, true>(Gp, Fp, ldg >> 1, ldf >> 1, lo0, hi0, lo1, hi1, s0, s_end, part_mod, part_rem, site_offset, r_own, tile, qoff, acc, own, cnt, cnt_own);
        else
            loo_like3_tile<KT, R, false>(Gp, Fp, ldg >> 1, ldf >> 1, lo0, hi0, lo1, hi1, s0, s_end, part_mod, part_rem, site_offset, r_own, tile, qoff, acc, own, cnt, cnt_own);
        __syncwarp();                                     // everyone is done with `buf` before it is refilled
    }
    acc[0].renorm(); acc[1].renorm(); own[0].renorm(); own[1].renorm();   // brings every slot to a known state
    if (col_ok) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            double* out = partials + ((long)blockIdx.y * ldg + col + i) * K;
#pragma unroll
            for (int kk = 0; kk < KT; ++kk)
                if (kk < K) out[pop_of_slot[kk]] = (kk == slot) ? own[i].value(0) : acc[i].value(kk);
        }
    }
}

// sums[col][pop_of_slot[kk]] += C[(kk before the column's own slot ? last : full)][kk] for kk != own slot: the
// population-only log(2a(1-a)) sums of the ratio form; C = [2 KP | 2 KP] doubles (last set, full set), run-order slots
__global__ void add_loo_const_kernel(double* __restrict__ sums, int ldg, int K, int KP, const int* __restrict__ slot_of_col,
                                     const int* __restrict__ pop_of_slot, const double* __restrict__ C)
{
    const long total = (long)ldg * K;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int col = (int)(e / K), kk = (int)(e - (long)col * K);
        const int own = slot_of_col[col];
        if (kk == own) continue;
        const int q = kk >> 1;
        sums[(long)col * K + pop_of_slot[kk]] += (2 * q + 1 <= own) ? C[kk] : C[2 * KP + kk];
    }
}

// A2[s][{last set: 2 KP | full set: 2 KP}] in run-order slots: the clipped leave-one-out estimate of each population's
// LAST member and the caller's full-data frequency; slots past K are 0.5 (a harmless dummy)
__global__ void loo_shared_af_kernel(const float* __restrict__ F, int ldf, int ldg, long M, int K, int KP,
                                     const int* __restrict__ lastcol_of_slot, const int* __restrict__ pop_of_slot,
                                     const float* __restrict__ clip_lo, const float* __restrict__ clip_hi, float* __restrict__ A2)
{
    const int W = 4 * KP;
    const long total = M * W;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long s = e / W;
        const int j = (int)(e - s * W);
        const int set = j >= 2 * KP, kk = set ? j - 2 * KP : j;
        float a = 0.5f;
        if (kk < K) {
            if (!set) {
                const int c = lastcol_of_slot[kk];
                a = fmin_nan(fmax_nan(F[s * (long)ldf + c], clip_lo[c]), clip_hi[c]);
            } else a = F[s * (long)ldf + ldg + pop_of_slot[kk]];
        }
        A2[e] = a;
    }
}

// ---------------------------------------------------------------------------------------
// EM posterior term of one individual at allele frequency f (emMAF_cy.pyx:19-22):
//   (p1 + 2 p2) / (2 (p0 + p1 + p2)),  p = GL * HWE(f)
// with H0 = 2(1-f)^2, H1 = 2f(1-f), H2 = 2f^2 precomputed per (site, problem):
//   num = g1 H1 + g2 H2 ; den = g0 H0 + g1 H1 + num.
// ---------------------------------------------------------------------------------------
struct EmCoef { float H0, H1, H2; };
__device__ __forceinline__ EmCoef em_coef(float f) {
    float om = 1.0f - f;
    EmCoef c;
    c.H0 = 2.0f * om * om; c.H1 = 2.0f * f * om; c.H2 = 2.0f * f * f;
    return c;
}
struct EmFrac { float num, rden; };
__device__ __forceinline__ EmFrac em_frac(float g0, float g1, float g2, const EmCoef& c) {
    EmFrac r;
    r.num = fmaf(g1, c.H1, g2 * c.H2);
    float den = fmaf(g0, c.H0, fmaf(g1, c.H1, r.num));
    r.rden = fast_rcp(den);
    return r;
}
__device__ __forceinline__ float em_term(float g0, float g1, float g2, const EmCoef& c) {
    float num = fmaf(g1, c.H1, g2 * c.H2);
    float den = fmaf(g0, c.H0, fmaf(g1, c.H1, num));
    return num * fast_rcp(den);
}

// ---------------------------------------------------------------------------------------
// em_pop_step: ONE EM iteration of every still-active population (emMAF_cy.pyx:10-23 for
// all K groups at once).  HBM-bound: 8 bytes per (site, individual, iteration).
// grid.y = population; a block owns tiles of R = blockDim.x consecutive sites of that
// population's slab.  The slab tile is streamed into shared memory with 16-byte asynchronous
// copies (a warp copies one row: contiguous, sector-aligned), double buffered, so the next
// tile is in flight while this one is consumed.  One thread = one site: it walks its row with
// conflict-free 128-bit loads (odd row stride) - no cross-lane reduction, no idle lanes.
// The state is population-major, FT[k][s], so that a warp's f values are contiguous.
// Each thread accumulates the squared change of its sites; partials[block.x][K] feed the
// global stop rule (emMAF_cy.pyx:26-33).
// ---------------------------------------------------------------------------------------
constexpr int kMaxKq = 4;   // K <= 128 (em_decide uses one lane per population in places)
constexpr int kEmT = 4;     // threads per site row: 4x the resident warps for the same shared-memory tile
__global__ void __launch_bounds__(512)
em_pop_step_kernel(const float2* __restrict__ G, int ldg, long M,
                   const PopDesc* __restrict__ pops, int K,
                   float* __restrict__ FT,               // [K][M], in place
                   const int* __restrict__ active,       // [K]
                   int row16,                            // shared-memory row stride in 16-byte units (== kEmT mod 8: conflict-free)
                   double* __restrict__ partials)        // [gridDim.x][K]
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red[512];
    __shared__ __align__(8) unsigned long long mbar[2];
    const int k = blockIdx.y;
    const int R = blockDim.x / kEmT;                      // rows (sites) per tile
    const int t = threadIdx.x;
    const int r = t / kEmT, h = t % kEmT;                 // row of this thread, its slice of the row
    if (!active[k]) {                                     // converged population: nothing to do
        if (t == 0) partials[(long)blockIdx.x * K + k] = 0.0;
        return;
    }
    const PopDesc pd = pops[k];
    const int cpr = (pd.n + 1) >> 1;                      // 16-byte chunks (pairs of individuals) per row that hold data
    float4* ring = reinterpret_cast<float4*>(smem_raw);   // [2][R][row16]
    const size_t stage = (size_t)R * row16;
    const long ntiles = (M + R - 1) / R;
    float* F = FT + (size_t)k * M;

    if (t == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_fence_init(); }
    __syncthreads();

    // warp 0 stages a tile: one TMA bulk copy per row (its slab segment is contiguous and 32-byte aligned)
    auto issue = [&](long tile, int buf) {
        if (tile < ntiles && t < 32) {
            const long s0 = tile * R;
            const int rows = (int)min((long)R, M - s0);
            if (t == 0) mbar_expect_tx(&mbar[buf], (unsigned)(rows * cpr * 16));
            __syncwarp();
            float4* dst = ring + buf * stage;
            for (int rr = t; rr < rows; rr += 32)
                bulk_g2s(dst + (size_t)rr * row16, G + (s0 + rr) * (long)ldg + pd.col0, (unsigned)(cpr * 16), &mbar[buf]);
        }
    };

    float ssq = 0.f;
    const float fn = (float)pd.n;
    const int full = pd.n >> 1;                           // complete pairs
    issue(blockIdx.x, 0);
    issue(blockIdx.x + (long)gridDim.x, 1);
    int it = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const long s = tile * R + r;
        float f = 0.25f;
        if (s < M) f = F[s];
        mbar_wait(&mbar[buf], (unsigned)((it >> 1) & 1)); // this tile has landed
        {
            const EmCoef c = em_coef(f);
            const float4* row = ring + buf * stage + (size_t)r * row16;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            if (s < M) {
                int q = h;
                for (; q + kEmT < full; q += 2 * kEmT) {
                    float4 v = row[q], w = row[q + kEmT];
                    a0 += em_term(v.x, v.y, 1.0f - v.x - v.y, c);
                    a1 += em_term(v.z, v.w, 1.0f - v.z - v.w, c);
                    a2 += em_term(w.x, w.y, 1.0f - w.x - w.y, c);
                    a3 += em_term(w.z, w.w, 1.0f - w.z - w.w, c);
                }
                if (q < full) {
                    float4 v = row[q];
                    a0 += em_term(v.x, v.y, 1.0f - v.x - v.y, c);
                    a1 += em_term(v.z, v.w, 1.0f - v.z - v.w, c);
                }
                if ((pd.n & 1) && h == (full % kEmT)) { float4 v = row[full]; a2 += em_term(v.x, v.y, 1.0f - v.x - v.y, c); }
            }
            float sum = (a0 + a1) + (a2 + a3);
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);    // the kEmT slices of a row sit in adjacent lanes
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            if (s < M && h == 0) {
                float fnew = __fdiv_rn(sum, fn);
                float d = fnew - f;
                ssq += d * d;
                F[s] = fnew;
            }
        }
        __syncthreads();                                  // everyone is done with this buffer before it is refilled
        issue(tile + 2 * (long)gridDim.x, buf);
    }
    red[t] = ssq;
    __syncthreads();
    if (t == 0) {
        double v = 0.0;
        for (int q = 0; q < (int)blockDim.x; q += kEmT) v += (double)red[q];
        partials[(long)blockIdx.x * K + k] = v;
    }
}

// ---------------------------------------------------------------------------------------
// em_pop_multi: up to kEmChunk EM iterations of every still-running population per read of G.
// A site's trajectory depends on nothing but its own row; only the STOP decision is global (the
// RMSE over all sites, emMAF.py:21-25).  So a tile that is already in shared memory is iterated
// T times in place, each thread keeping one squared-change accumulator per iteration; the host
// then finds the first iteration whose RMSE is below the tolerance.  If that iteration t* lies
// inside the chunk, the population is replayed for exactly t* iterations from the chunk's start
// state (FT[cur]) - still 3 reads of G for a 14-iteration EM instead of 14.  The arithmetic per
// site and iteration, and the order of every sum, are those of em_pop_step: same bits.
// cur[k] selects the buffer holding population k's current state; the result goes to the other.
// ---------------------------------------------------------------------------------------
constexpr int kEmChunk = 8;
__global__ void __launch_bounds__(512)
em_pop_multi_kernel(const float2* __restrict__ G, int ldg, long M,
                    const PopDesc* __restrict__ pops, int K,
                    float* __restrict__ FT0, float* __restrict__ FT1,   // [K][M] each
                    const int* __restrict__ cur,          // [K] buffer that holds the start state
                    const int* __restrict__ iters_k,      // [K] iterations to run now (0 = skip)
                    int row16,                            // shared-memory row stride in 16-byte units (== kEmT mod 8: conflict-free)
                    double* __restrict__ partials)        // [gridDim.x][K][kEmChunk]
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red[512];
    __shared__ __align__(8) unsigned long long mbar[2];
    const int k = blockIdx.y;
    const int R = blockDim.x / kEmT;                      // rows (sites) per tile
    const int t = threadIdx.x;
    const int r = t / kEmT, h = t % kEmT;                 // row of this thread, its slice of the row
    const int T = iters_k[k];
    double* pout = partials + ((long)blockIdx.x * K + k) * kEmChunk;
    if (T <= 0) {                                         // finished population: nothing to do
        if (t < kEmChunk) pout[t] = 0.0;
        return;
    }
    const PopDesc pd = pops[k];
    const int cpr = (pd.n + 1) >> 1;                      // 16-byte chunks (pairs of individuals) per row that hold data
    float4* ring = reinterpret_cast<float4*>(smem_raw);   // [2][R][row16]
    const size_t stage = (size_t)R * row16;
    const long ntiles = (M + R - 1) / R;
    const float* Fsrc = (cur[k] ? FT1 : FT0) + (size_t)k * M;
    float* Fdst = (cur[k] ? FT0 : FT1) + (size_t)k * M;

    if (t == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_fence_init(); }
    __syncthreads();
    auto issue = [&](long tile, int buf) {
        if (tile < ntiles && t < 32) {
            const long s0 = tile * R;
            const int rows = (int)min((long)R, M - s0);
            if (t == 0) mbar_expect_tx(&mbar[buf], (unsigned)(rows * cpr * 16));
            __syncwarp();
            float4* dst = ring + buf * stage;
            for (int rr = t; rr < rows; rr += 32)
                bulk_g2s(dst + (size_t)rr * row16, G + (s0 + rr) * (long)ldg + pd.col0, (unsigned)(cpr * 16), &mbar[buf]);
        }
    };

    float ssq[kEmChunk];
#pragma unroll
    for (int u = 0; u < kEmChunk; ++u) ssq[u] = 0.f;
    const float fn = (float)pd.n;
    const int full = pd.n >> 1;                           // complete pairs
    issue(blockIdx.x, 0);
    issue(blockIdx.x + (long)gridDim.x, 1);
    int it = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const long s = tile * R + r;
        float f = 0.25f;
        if (s < M) f = Fsrc[s];
        mbar_wait(&mbar[buf], (unsigned)((it >> 1) & 1)); // this tile has landed
        const float4* row = ring + buf * stage + (size_t)r * row16;
#pragma unroll
        for (int u = 0; u < kEmChunk; ++u) {
            if (u < T) {                                  // block-uniform
                const EmCoef c = em_coef(f);
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                if (s < M) {
                    int q = h;
                    for (; q + kEmT < full; q += 2 * kEmT) {
                        float4 v = row[q], w = row[q + kEmT];
                        a0 += em_term(v.x, v.y, 1.0f - v.x - v.y, c);
                        a1 += em_term(v.z, v.w, 1.0f - v.z - v.w, c);
                        a2 += em_term(w.x, w.y, 1.0f - w.x - w.y, c);
                        a3 += em_term(w.z, w.w, 1.0f - w.z - w.w, c);
                    }
                    if (q < full) {
                        float4 v = row[q];
                        a0 += em_term(v.x, v.y, 1.0f - v.x - v.y, c);
                        a1 += em_term(v.z, v.w, 1.0f - v.z - v.w, c);
                    }
                    if ((pd.n & 1) && h == (full % kEmT)) { float4 v = row[full]; a2 += em_term(v.x, v.y, 1.0f - v.x - v.y, c); }
                }
                float sum = (a0 + a1) + (a2 + a3);
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);    // the kEmT slices of a row sit in adjacent lanes
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                const float fnew = __fdiv_rn(sum, fn);          // identical in the kEmT lanes of a row
                if (s < M && h == 0) { const float d = fnew - f; ssq[u] += d * d; }
                f = fnew;
            }
        }
        if (s < M && h == 0) Fdst[s] = f;
        __syncthreads();                                  // everyone is done with this buffer before it is refilled
        issue(tile + 2 * (long)gridDim.x, buf);
    }
#pragma unroll
    for (int u = 0; u < kEmChunk; ++u) {
        __syncthreads();
        red[t] = ssq[u];
        __syncthreads();
        if (t == 0) {
            double v = 0.0;
            for (int q = 0; q < (int)blockDim.x; q += kEmT) v += (double)red[q];
            pout[u] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------
// em_pop_multi2: em_pop_multi with the tile in REGISTERS and the arithmetic of the leave-one-out
// quad kernel.  One thread per (site row, slice) reads its QPT quads of individuals from the landed
// raw rows ONCE, packs them ({g0_ab,g1_ab,g2_ab,g0_cd,g1_cd,g2_cd} per quad) and iterates up to 8
// times without touching shared memory again: with one problem per thread a shared-memory tile
// costs 12 bytes per posterior term and iteration, which bound the first packed version at
// 0.3 ms per iteration.  Per four terms and iteration: 10 packed FP32 instructions + 2 MUFU.RCP
// (ratio form, one reciprocal per two terms).  TPR threads share a row (TPR * QPT * 4 >= n).
// f is kept inside [1e-12, 1-2^-24] (the ratio form divides by f and 1-f; a monomorphic site of
// certain genotypes would otherwise reach exactly 0) - 7 orders below the parity tolerance and
// far outside the clipping applied afterwards.
// ---------------------------------------------------------------------------------------
template <int TPR, int QPT>
__global__ void __launch_bounds__(256)
em_pop_multi2_kernel(const float2* __restrict__ G, int ldg, long M,
                     const PopDesc* __restrict__ pops, int K,
                     float* __restrict__ FT0, float* __restrict__ FT1,   // [K][M] each
                     const int* __restrict__ cur,          // [K] buffer that holds the start state
                     const int* __restrict__ iters_k,      // [K] iterations to run now (0 = skip)
                     int raw16,                            // raw row stride, 16-byte units (an odd number of pairs of units)
                     double* __restrict__ partials,        // [gridDim.x][K][kEmChunk]
                     float* __restrict__ Fhist)            // [kEmChunk][K][M]: the state after every iteration of this pass
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float red[256];
    const int k = blockIdx.y;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int R = 256 / TPR;                          // rows (sites) per tile
    constexpr int RW = 32 / TPR;                          // rows per warp: the warp stages and consumes exactly these
    const int r = t / TPR, h = t % TPR;                   // row of this thread, its slice of the row
    const int T = iters_k[k];
    double* pout = partials + ((long)blockIdx.x * K + k) * kEmChunk;
    if (T <= 0) {                                         // finished population: nothing to do
        if (t < kEmChunk) pout[t] = 0.0;
        return;
    }
    const PopDesc pd = pops[k];
    const int nq = (pd.n + 3) >> 2;                       // quads of individuals
    const int nq2 = 2 * nq;                               // 16-byte units per slab row
    float4* raw = reinterpret_cast<float4*>(smem_raw);    // [2][R][raw16]
    const long ntiles = (M + R - 1) / R;
    const float* Fsrc = (cur[k] ? FT1 : FT0) + (size_t)k * M;
    float* Fdst = (cur[k] ? FT0 : FT1) + (size_t)k * M;

    // Warp-private staging: every warp copies the RW slab rows it will consume with 16-byte LDGSTS into its own
    // slice of a double buffer, one tile ahead - no block-wide barrier in the loop (the TMA + __syncthreads version
    // stalled 3 cycles per issue at that barrier), warps drift freely.
    // The (row, unit) each lane copies is the same for every tile: worked out once, as a shared-memory offset and a
    // global offset (in 16-byte units) per copy - the copy loop itself is then 8 predicated LDGSTS.  (Walking the
    // row/unit counters in the loop was a quarter of this kernel's instructions.)
    constexpr int kCopies = (RW * 2 * TPR * QPT + 31) / 32;        // units of the warp's RW rows / 32 lanes, at most 8
    int so[kCopies], go[kCopies];
    {
        int rr = 0, q = lane;
#pragma unroll
        for (int i = 0; i < kCopies; ++i) {
            while (q >= nq2 && rr < RW) { q -= nq2; ++rr; }
            so[i] = rr < RW ? rr * raw16 + q : -1;
            go[i] = rr < RW ? rr * (ldg >> 1) + q : 0;              // a row of G is ldg pairs = ldg / 2 units
            q += 32;
        }
    }
    const float4* Gslab = reinterpret_cast<const float4*>(G + pd.col0);
    auto stage = [&](long tile, int buf) {
        if (tile < ntiles) {
            const long s0 = tile * R + warp * RW;
            float4* dst = raw + ((size_t)buf * R + warp * RW) * raw16;
            const float4* src = Gslab + s0 * (long)(ldg >> 1);
            const int rows_ok = (int)min((long)RW, M - s0);       // rows of this warp's slice that exist (<= 0: none)
#pragma unroll
            for (int i = 0; i < kCopies; ++i)
                if (so[i] >= 0 && (rows_ok >= RW || so[i] < rows_ok * raw16)) cp_async16(dst + so[i], src + go[i]);
        }
        cp_async_commit();
    };

    float ssq[kEmChunk];
#pragma unroll
    for (int u = 0; u < kEmChunk; ++u) ssq[u] = 0.f;
    const float fn = (float)pd.n;
    stage(blockIdx.x, 0);
    int it = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const long s = tile * R + r;
        float f = 0.25f;
        if (s < M) f = Fsrc[s];
        stage(tile + gridDim.x, buf ^ 1);                 // always commits a (possibly empty) group: wait<1> = this tile has landed
        cp_async_wait<1>();
        __syncwarp();
        // this thread's quads -> packed registers (pads and rows past M are (1,0,0): exactly zero terms)
        f32x2 g0ab[QPT], g1ab[QPT], g2ab[QPT], g0cd[QPT], g1cd[QPT], g2cd[QPT];
        {
            const float4* src = raw + ((size_t)buf * R + r) * raw16;
#pragma unroll
            for (int j = 0; j < QPT; ++j) {
                const int q = h + j * TPR;
                float4 ga = make_float4(1.f, 0.f, 1.f, 0.f), gc = ga;
                if (s < M && q < nq) { ga = src[2 * q]; gc = src[2 * q + 1]; }
                if (4 * q + 1 >= pd.n) { ga.z = 1.f; ga.w = 0.f; }
                if (4 * q + 2 >= pd.n) { gc.x = 1.f; gc.y = 0.f; }
                if (4 * q + 3 >= pd.n) { gc.z = 1.f; gc.w = 0.f; }
                g0ab[j] = pack2(ga.x, ga.z); g1ab[j] = pack2(ga.y, ga.w);
                g2ab[j] = pack2(1.0f - ga.x - ga.y, 1.0f - ga.z - ga.w);
                g0cd[j] = pack2(gc.x, gc.z); g1cd[j] = pack2(gc.y, gc.w);
                g2cd[j] = pack2(1.0f - gc.x - gc.y, 1.0f - gc.z - gc.w);
            }
        }
        __syncwarp();                                     // every lane has its registers: the slice may be refilled next time round
#pragma unroll
        for (int u = 0; u < kEmChunk; ++u) {
            if (u < T) {                                  // block-uniform
                const float om = 1.0f - f;
                const float ca = om * fast_rcp(f), cb = f * fast_rcp(om);
                const f32x2 A = pack2(ca, ca), B = pack2(cb, cb);
                float accx = 0.f, accy = 0.f;                 // scalar accumulation: no register-pair packing of the reciprocals
#pragma unroll
                for (int j = 0; j < QPT; ++j) {
                    const f32x2 nu = ffma2(g2ab[j], B, g1ab[j]);
                    const f32x2 nv = ffma2(g2cd[j], B, g1cd[j]);
                    const f32x2 du = ffma2(g0ab[j], A, fadd2(g1ab[j], nu));
                    const f32x2 dv = ffma2(g0cd[j], A, fadd2(g1cd[j], nv));
                    const f32x2 m = fmul2(du, dv);
                    const f32x2 x = ffma2(nv, du, fmul2(nu, dv));
                    const float2 mm = unpack2(m), xx = unpack2(x);
                    accx = fmaf(xx.x, fast_rcp(mm.x), accx);
                    accy = fmaf(xx.y, fast_rcp(mm.y), accy);
                }
                float sum = accx + accy;
#pragma unroll
                for (int o = 1; o < TPR; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);   // the TPR slices of a row sit in adjacent lanes
                float fnew = __fdiv_rn(sum, fn);                // identical in the TPR lanes of a row
                if (fnew < 1e-12f) fnew = 1e-12f;               // comparisons are false for NaN: NaN survives
                if (fnew > 0.99999994f) fnew = 0.99999994f;
                if (s < M && h == 0) {
                    const float d = fnew - f; ssq[u] += d * d;
                    Fhist[((size_t)u * K + k) * M + s] = fnew;      // the host picks the stop iteration's state: no replay pass
                }
                f = fnew;
            }
        }
        if (s < M && h == 0) Fdst[s] = f;
    }
#pragma unroll
    for (int u = 0; u < kEmChunk; ++u) {
        __syncthreads();
        red[t] = ssq[u];
        __syncthreads();
        if (t == 0) {
            double v = 0.0;
            for (int q = 0; q < (int)blockDim.x; q += TPR) v += (double)red[q];
            pout[u] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Leave-one-out EM (glassy.py:65-78: emMAF on the population minus individual i, for every i).
// ONE launch = ONE EM iteration of every still-active leave-one-out problem of ONE population.
// Problem i at site s follows its own trajectory f_i, so each (site, i) evaluates all n
// posterior terms at f_i: n^2 evaluations per site and iteration from ONE read of the
// population's GL tile, staged in shared memory.
//
// History of the inner loop (each step measured on B200, 1 M sites x 50 individuals):
//   pair kernel  - two problems per thread, FMUL2 + 3 FFMA2 + 2 MUFU.RCP + 2 FFMA per two evaluations:
//                  0.936 ms, 57 % of the MUFU reciprocal rate, which bound it (removed);
//   quad kernel  - loo_em_step4 below: ratio form, one reciprocal per two evaluations, four problems
//                  per thread, TMA-staged rows: 0.776 ms;
//   packed kernel - loo_em_step5: pre-packed pair polynomials, 8 packed instructions per four evaluations.
//
// The left-out individual's own term is removed after the loop.  (sum - own) can cancel to
// (nearly) 0 when nobody else carries the allele, and an exact 0 (or 1) would turn the next
// iteration's 0 * inf into NaN where the reference stays finite, so f is kept inside
// [1e-12, 1-2^-24]; both are far outside the clipping range applied afterwards
// (glassy.py:80-85) and 7 orders below the 1e-5 parity tolerance.
//
// A thread keeps its problems for the whole launch, so their squared changes are
// registers; partials[block][col] are reduced in fixed order by em_ssq_reduce_kernel.
// mask (optional): uchar keep[M][ldg] - sites whose squared change counts (reference
// z-score runs the EM on kept sites only, WGSassign.py:358-359; a site outside the mask
// cannot influence another site, so it is simply skipped).
// ---------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------
// loo_em_step4: the EM iteration re-derived so that the MUFU reciprocal is no longer the
// binding unit (it was: 57 % of 16 RCP/clk/SM at 1 M x 50).
//
//  * ratio form.  Dividing numerator and denominator of the posterior mean
//        (p1 + 2 p2) / (2 (p0 + p1 + p2)),  p = (g0 (1-f)^2, 2 g1 f (1-f), g2 f^2)
//    by 2 f (1-f) leaves TWO per-problem coefficients a = (1-f)/f, b = f/(1-f):
//        num = g1 + b g2,   den = a g0 + g1 + num          (FFMA2, FADD2, FFMA2 per packed pair)
//    With f inside [1e-12, 1-2^-24] both coefficients are finite and den >= 1e-12.
//  * one reciprocal per TWO posterior terms:  n1/d1 + n2/d2 = (n1 d2 + n2 d1) / (d1 d2).
//    The tile holds QUADS of individuals as two packed pairs (a,b) and (c,d); pair lanes are
//    combined across the two pairs, so the cross products stay packed (FMUL2, FMUL2, FFMA2),
//    then 2 MUFU.RCP and one packed accumulate.  d1 d2 is within [1e-24, 1e25]: no scaling.
//    Per four evaluations: 10 packed FP32 instructions + 2 MUFU (was 8 packed + 4 FFMA + 4 MUFU).
//  * one thread owns FOUR problems (the members of quad ti) and feeds them from the same
//    three LDS.128, halving the delivered shared-memory bytes per evaluation once more.
//
// Tile layout: tile[site][3 q + {0,1,2}] = {g0_ab, g1_ab}, {g2_ab, g0_cd}, {g1_cd, g2_cd}
// (16-byte units, odd row stride).  Pad individuals are (1,0,0): num = 0 exactly.
// The left-out individual's own term is subtracted after the loop as a plain n/d; unlike in
// the earlier pair kernel the cancellation is not bit-exact (the sum holds it inside a combined fraction),
// so "nobody else carries the allele" gives |f| ~ 1e-8 instead of the 1e-12 floor - both are
// orders of magnitude below the clipping bound applied afterwards (glassy.py:80-85).
// ---------------------------------------------------------------------------------------
constexpr int kLoo4MaxPasses = 4;
struct Loo4Coef { f32x2 A, B; float a, b; };
__device__ __forceinline__ Loo4Coef loo4_coef(float f) {
    Loo4Coef c;
    const float om = 1.0f - f;
    c.a = om * fast_rcp(f);
    c.b = f * fast_rcp(om);
    c.A = pack2(c.a, c.a); c.B = pack2(c.b, c.b);
    return c;
}
__device__ __forceinline__ float loo4_own(float g0, float g1, float g2, const Loo4Coef& c) {
    const float num = fmaf(g2, c.b, g1);
    const float den = fmaf(g0, c.a, g1 + num);
    return num * fast_rcp(den);
}
__device__ __forceinline__ void loo4_quad(const ulonglong2* __restrict__ q3, const Loo4Coef (&c)[4], f32x2 (&acc)[4]) {
    const ulonglong2 v0 = q3[0], v1 = q3[1], v2 = q3[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const f32x2 nu = ffma2(v1.x, c[k].B, v0.y);
        const f32x2 nv = ffma2(v2.y, c[k].B, v2.x);
        const f32x2 du = ffma2(v0.x, c[k].A, fadd2(v0.y, nu));
        const f32x2 dv = ffma2(v1.y, c[k].A, fadd2(v2.x, nv));
        const f32x2 m = fmul2(du, dv);
        const f32x2 x = ffma2(nv, du, fmul2(nu, dv));
        const float2 mm = unpack2(m);
        acc[k] = ffma2(x, pack2(fast_rcp(mm.x), fast_rcp(mm.y)), acc[k]);
    }
}

// Staging: warp 0 streams the RAW slab rows of the next tile into shared memory with TMA bulk
// copies (one per row, mbarrier-tracked) while the block computes the current tile; after its
// own compute every thread repacks its share of the landed rows into the other packed buffer,
// so the only block-wide barrier per tile sits where all warps have done the same work and no
// warp ever waits on HBM latency.
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
loo_em_step4_kernel(const float2* __restrict__ G, int ldg, long M,
                    int col0, int n, int rows_per_pass, int passes,
                    float* __restrict__ F, int ldf,              // [M][ldf], in place
                    const int* __restrict__ active,              // [ldg]
                    const unsigned char* __restrict__ mask,      // [M][ldg] or null
                    double* __restrict__ partials,               // [gridDim.x][ldg]
                    long ntiles,
                    float* __restrict__ D2)                      // [ldg / 4][M][4] (quad-major) or null: this iteration's squared change per (problem, site)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long mbar;
    const int nq = (n + 3) >> 2;                                // quads of individuals = threads per site row
    const int stride = (3 * nq) | 1;                            // 16-byte units, odd: rows never share a bank group
    const int TS = rows_per_pass * passes;
    ulonglong2* tile0 = reinterpret_cast<ulonglong2*>(smem_raw);                // [2][TS][stride]  packed quads
    float4* raw = reinterpret_cast<float4*>(tile0 + 2 * (size_t)TS * stride);   // [TS][2 nq]       raw slab rows
    float4* red = raw + (size_t)TS * 2 * nq;                                    // [blockDim.x]

    const int t = threadIdx.x;
    const int ti = t % nq, r = t / nq;                          // this thread's quad (problems 4ti..4ti+3) and row
    const bool worker = t < rows_per_pass * nq;
    const int c0 = col0 + 4 * ti;
    unsigned act = 0;                                           // bit k: problem 4ti+k exists and is still iterating
    if (worker) {
        const int4 a4 = *reinterpret_cast<const int4*>(&active[c0]);
        act = (a4.x != 0 ? 1u : 0u) | (a4.y != 0 ? 2u : 0u) | (a4.z != 0 ? 4u : 0u) | (a4.w != 0 ? 8u : 0u);
        if (4 * ti + 1 >= n) act &= 1u;
        if (4 * ti + 2 >= n) act &= 3u;
        if (4 * ti + 3 >= n) act &= 7u;
    }
    const float inv_div = 1.0f / (float)(n - 1);
    float ssq[4] = {0.f, 0.f, 0.f, 0.f};

    if (t == 0) { mbar_init(&mbar, 1); mbar_fence_init(); }
    __syncthreads();
    // warp 0: one TMA bulk copy per slab row of tile `tl` (rows are contiguous and 32-byte aligned)
    auto issue = [&](long tl) {
        if (tl < ntiles && t < 32) {
            const long s0 = tl * TS;
            const int rows = (int)min((long)TS, M - s0);
            if (t == 0) mbar_expect_tx(&mbar, (unsigned)(rows * nq * 32));
            __syncwarp();
            for (int rr = t; rr < rows; rr += 32)
                bulk_g2s(raw + (size_t)rr * 2 * nq, G + (s0 + rr) * (long)ldg + col0, (unsigned)(nq * 32), &mbar);
        }
    };
    // raw rows -> packed quads of buffer `dst` (rows past M are never read by the compute phase)
    auto repack = [&](ulonglong2* dst, long tl) {
        const int rows = (int)min((long)TS, M - tl * TS);
        for (int e = t; e < rows * nq; e += blockDim.x) {
            const int sl = e / nq, q = e - sl * nq;
            float4 ga = raw[(size_t)sl * 2 * nq + 2 * q], gc = raw[(size_t)sl * 2 * nq + 2 * q + 1];
            if (4 * q + 1 >= n) { ga.z = 1.f; ga.w = 0.f; }     // slab padding: (1,0,0) contributes exactly 0
            if (4 * q + 2 >= n) { gc.x = 1.f; gc.y = 0.f; }
            if (4 * q + 3 >= n) { gc.z = 1.f; gc.w = 0.f; }
            ulonglong2 v0, v1, v2;
            v0.x = pack2(ga.x, ga.z);                           // g0 of (a,b)
            v0.y = pack2(ga.y, ga.w);                           // g1 of (a,b)
            v1.x = pack2(third_gl(ga.x, ga.y), third_gl(ga.z, ga.w));
            v1.y = pack2(gc.x, gc.z);                           // g0 of (c,d)
            v2.x = pack2(gc.y, gc.w);                           // g1 of (c,d)
            v2.y = pack2(third_gl(gc.x, gc.y), third_gl(gc.z, gc.w));
            ulonglong2* d = dst + sl * stride + 3 * q;
            d[0] = v0; d[1] = v1; d[2] = v2;
        }
    };

    unsigned phase = 0;
    if (blockIdx.x < ntiles) {                                  // prologue: first tile packed, second in flight
        issue(blockIdx.x);
        mbar_wait(&mbar, phase); phase ^= 1u;
        repack(tile0, blockIdx.x);
        __syncthreads();
        issue(blockIdx.x + (long)gridDim.x);
    }
    int cur = 0;
    for (long tl = blockIdx.x; tl < ntiles; tl += gridDim.x, cur ^= 1) {
        const long s0 = tl * TS;
        const ulonglong2* tile = tile0 + (size_t)cur * TS * stride;
        // this thread's f quad of pass p+1 is fetched while pass p computes
        unsigned ok_next;
        float4 f_next;
        auto fetch = [&](int p) {
            const long s = s0 + p * rows_per_pass + r;
            ok_next = (p < passes && s < M) ? act : 0u;
            if (ok_next && mask) {
                const uchar4 mk = *reinterpret_cast<const uchar4*>(&mask[s * (long)ldg + c0]);
                ok_next &= (mk.x ? 1u : 0u) | (mk.y ? 2u : 0u) | (mk.z ? 4u : 0u) | (mk.w ? 8u : 0u);
            }
            f_next = make_float4(0.25f, 0.25f, 0.25f, 0.25f);
            if (ok_next) f_next = *reinterpret_cast<const float4*>(&F[s * (long)ldf + c0]);
        };
        fetch(0);
#pragma unroll 1
        for (int p = 0; p < passes; ++p) {
            const unsigned okp = ok_next;
            const float4 fq = f_next;
            fetch(p + 1);
            if (!okp) continue;
            const int sl = p * rows_per_pass + r;
            const float fin[4] = {fq.x, fq.y, fq.z, fq.w};
            Loo4Coef c[4];
            f32x2 acc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { c[k] = loo4_coef(fin[k]); acc[k] = 0ull; }
            const ulonglong2* row = tile + sl * stride;
            int q = 0;
#pragma unroll 1
            for (; q + 1 < nq; q += 2) {
                loo4_quad(row + 3 * q, c, acc);
                loo4_quad(row + 3 * q + 3, c, acc);
            }
            if (q < nq) loo4_quad(row + 3 * q, c, acc);
            // own terms: the four members of quad ti
            const ulonglong2 v0 = row[3 * ti], v1 = row[3 * ti + 1], v2 = row[3 * ti + 2];
            const float2 g0ab = unpack2(v0.x), g1ab = unpack2(v0.y), g2ab = unpack2(v1.x);
            const float2 g0cd = unpack2(v1.y), g1cd = unpack2(v2.x), g2cd = unpack2(v2.y);
            const float own[4] = {loo4_own(g0ab.x, g1ab.x, g2ab.x, c[0]), loo4_own(g0ab.y, g1ab.y, g2ab.y, c[1]),
                                  loo4_own(g0cd.x, g1cd.x, g2cd.x, c[2]), loo4_own(g0cd.y, g1cd.y, g2cd.y, c[3])};
            float fo[4], dq[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 a = unpack2(acc[k]);
                float fn = ((a.x + a.y) - own[k]) * inv_div;
                if (fn < 1e-12f) fn = 1e-12f;                   // comparisons are false for NaN: NaN survives
                if (fn > 0.99999994f) fn = 0.99999994f;
                fo[k] = fin[k];                                 // frozen / masked problems keep their value
                dq[k] = 0.f;
                if (okp & (1u << k)) {
                    const float d = __fsub_rn(fn, fin[k]);
                    dq[k] = __fmul_rn(d, d);
                    ssq[k] = __fadd_rn(ssq[k], dq[k]);
                    fo[k] = fn;
                }
            }
            *reinterpret_cast<float4*>(&F[(s0 + sl) * (long)ldf + c0]) = make_float4(fo[0], fo[1], fo[2], fo[3]);
            if (D2) *reinterpret_cast<float4*>(&D2[((long)(c0 >> 2) * M + (s0 + sl)) * 4]) = make_float4(dq[0], dq[1], dq[2], dq[3]);
        }
        const long nxt = tl + gridDim.x;
        if (nxt < ntiles) {                                     // the next tile's raw rows landed while this one computed
            mbar_wait(&mbar, phase); phase ^= 1u;
            repack(tile0 + (size_t)(cur ^ 1) * TS * stride, nxt);
        }
        __syncthreads();                                        // packed[cur] consumed, packed[cur^1] complete, raw free
        issue(nxt + gridDim.x);
    }
    red[t] = make_float4(ssq[0], ssq[1], ssq[2], ssq[3]);
    __syncthreads();
    for (int p = t; p < n; p += blockDim.x) {                   // problem p = member (p & 3) of quad p / 4 (n may exceed the block)
        double v = 0.0;
        for (int q = 0; q < rows_per_pass; ++q) {
            const float4 x = red[q * nq + (p >> 2)];
            const int k = p & 3;
            v += (double)(k == 0 ? x.x : k == 1 ? x.y : k == 2 ? x.z : x.w);
        }
        partials[(long)blockIdx.x * ldg + col0 + p] = v;
    }
}

// ---------------------------------------------------------------------------------------
// loo_em_step5: the quad iteration on PRE-PACKED pair polynomials.
//
// With d_i(b) = g0_i a + 2 g1_i + g2_i b and n_i(b) = g1_i + g2_i b (a = (1-f)/f = 1/b), the two posterior
// terms of a PAIR of individuals are one fraction N/D whose numerator and denominator are Laurent
// polynomials in b with coefficients that depend on the two individuals only:
//     D = d_i d_j         = c0 a^2 + 2 c1 a + c2 + 2 c3 b + c4 b^2
//     N = n_i d_j + n_j d_i =          c1 a + c2 + 3 c3 b + 2 c4 b^2
//     c0 = g0 g0', c1 = g0 g1' + g1 g0', c2 = 4 g1 g1' + g0 g2' + g2 g0', c3 = g1 g2' + g2 g1', c4 = g2 g2'
// (all terms non-negative: no cancellation anywhere).  The five coefficients are the same for all
// ~15 iterations of all n problems, so loo_prepack_kernel writes them ONCE per call as planes
// P = (c0, 2 c1, c2, 2 c3, c4), packed so that the two pairs (a,b), (c,d) of a quad sit in the two lanes
// of an f32x2.  Per four evaluations the step kernel then issues 7 FFMA2 + 2 MUFU.RCP + 1 FFMA2
// (the quad kernel: 10 packed + 2), reads 40 instead of 48 bytes of shared memory and has no packing
// code at all: warp 0 streams packed rows and raw rows (for the left-out individual's own term)
// straight into a double buffer with TMA bulk copies.
// Cell = two quads (8 individuals) = 5 x 16 bytes: {P0A,P1A} {P2A,P3A} {P4A,P0B} {P1B,P2B} {P3B,P4B}.
// A pad individual is (1,0,0): a pair of pads gives D = a^2, N = 0; one pad gives exactly the
// other member's n/d.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void loo5_pair_coefs(float g0, float g1, float g2, float h0, float h1, float h2, float (&P)[5])
{
    P[0] = g0 * h0;
    P[1] = 2.0f * fmaf(g0, h1, g1 * h0);
    P[2] = fmaf(4.0f * g1, h1, fmaf(g0, h2, g2 * h0));
    P[3] = 2.0f * fmaf(g1, h2, g2 * h1);
    P[4] = g2 * h2;
}
// Row layout (16-byte units): [5 nc packed cells | 2 nq raw float4 (the slab's (g0,g1) pairs, for the own terms) | pad to an odd count]
// - rows of consecutive sites are contiguous, so a whole group of rows is ONE bulk copy, and the odd row
// length keeps the rows of a quarter-warp on different bank groups.
__host__ __device__ __forceinline__ int loo5_row_units(int n) { const int nq = (n + 3) >> 2, nc = (nq + 1) >> 1; return (5 * nc + 2 * nq) | 1; }
__global__ void __launch_bounds__(256)
loo_prepack_kernel(const float2* __restrict__ G, int ldg, long M, int col0, int n, int nc,
                   ulonglong2* __restrict__ PK)          // [M][loo5_row_units(n)]
{
    const int nq = (n + 3) >> 2;
    const int ru = loo5_row_units(n);
    const long total = M * (long)nc;
    const long T = (long)gridDim.x * blockDim.x;
    // two cells per thread and trip: all eight 16-byte loads are issued before the first use (the one-cell version
    // sat on its loads: 54 stall cycles per issue at 38 % of DRAM bandwidth)
    for (long e0 = blockIdx.x * (long)blockDim.x + threadIdx.x; e0 < total; e0 += 2 * T) {
        float4 ga[2][2], gc[2][2];
        long sv[2]; int cv[2]; bool on[2];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const long e = e0 + w * T;
            on[w] = e < total;
            sv[w] = on[w] ? e / nc : 0;
            cv[w] = on[w] ? (int)(e - sv[w] * nc) : 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int q = 2 * cv[w] + h;
                ga[w][h] = make_float4(1.f, 0.f, 1.f, 0.f); gc[w][h] = ga[w][h];     // (1,0,0) pads
                if (on[w] && q < nq) {
                    const float4* src = reinterpret_cast<const float4*>(&G[sv[w] * (long)ldg + col0 + 4 * q]);
                    ga[w][h] = ld_stream4(src);
                    gc[w][h] = ld_stream4(src + 1);
                }
            }
        }
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            if (!on[w]) continue;
            const int c = cv[w];
            ulonglong2* rowp = PK + sv[w] * (long)ru;
            f32x2 P[2][5];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int q = 2 * c + h;
                float4 a4 = ga[w][h], c4 = gc[w][h];
                if (q < nq) {
                    float4* rawp = reinterpret_cast<float4*>(rowp + 5 * nc) + 2 * q;
                    rawp[0] = a4; rawp[1] = c4;
                }
                if (4 * q + 1 >= n) { a4.z = 1.f; a4.w = 0.f; }
                if (4 * q + 2 >= n) { c4.x = 1.f; c4.y = 0.f; }
                if (4 * q + 3 >= n) { c4.z = 1.f; c4.w = 0.f; }
                float Pab[5], Pcd[5];
                loo5_pair_coefs(a4.x, a4.y, third_gl(a4.x, a4.y), a4.z, a4.w, third_gl(a4.z, a4.w), Pab);
                loo5_pair_coefs(c4.x, c4.y, third_gl(c4.x, c4.y), c4.z, c4.w, third_gl(c4.z, c4.w), Pcd);
#pragma unroll
                for (int j = 0; j < 5; ++j) P[h][j] = pack2(Pab[j], Pcd[j]);
            }
            ulonglong2* dst = rowp + 5 * c;
            ulonglong2 v;
            v.x = P[0][0]; v.y = P[0][1]; dst[0] = v;
            v.x = P[0][2]; v.y = P[0][3]; dst[1] = v;
            v.x = P[0][4]; v.y = P[1][0]; dst[2] = v;
            v.x = P[1][1]; v.y = P[1][2]; dst[3] = v;
            v.x = P[1][3]; v.y = P[1][4]; dst[4] = v;
            if (c == 0 && (5 * nc + 2 * nq) != ru) rowp[ru - 1] = make_ulonglong2(0ull, 0ull);   // the pad unit: defined bytes for the copy
        }
    }
}

// loo_prepack2: the same rows, written through shared memory.  A block packs whole site rows (256 / nc rows per
// tile, one thread per cell) into a tile with the global row layout and then copies the tile out as one contiguous
// run of 16-byte units - full-line coalesced stores.  (One thread storing its own 80 + 64 bytes made every warp
// store touch 32 partial sectors: 2.5 TB/s.)
__global__ void __launch_bounds__(256)
loo_prepack2_kernel(const float2* __restrict__ G, int ldg, long M, int col0, int n, int nc,
                    ulonglong2* __restrict__ PK)         // [M][loo5_row_units(n)]
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ulonglong2* tile = reinterpret_cast<ulonglong2*>(smem_raw);      // [RB][ru]
    const int nq = (n + 3) >> 2;
    const int ru = loo5_row_units(n);
    const int RB = blockDim.x / nc;                       // rows per tile
    const int t = threadIdx.x, r = t / nc, c = t - r * nc;
    const bool worker = r < RB;
    const long ntiles = (M + RB - 1) / RB;
    for (long tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const long s0 = tl * RB;
        const int rows = (int)min((long)RB, M - s0);
        const long s = s0 + r;
        const bool on = worker && r < rows;
        float4 ga[2], gc[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int q = 2 * c + h;
            ga[h] = make_float4(1.f, 0.f, 1.f, 0.f); gc[h] = ga[h];          // (1,0,0) pads
            if (on && q < nq) {
                const float4* src = reinterpret_cast<const float4*>(&G[s * (long)ldg + col0 + 4 * q]);
                ga[h] = ld_stream4(src);
                gc[h] = ld_stream4(src + 1);
            }
        }
        if (on) {
            ulonglong2* rowp = tile + (size_t)r * ru;
            f32x2 P[2][5];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int q = 2 * c + h;
                float4 a4 = ga[h], c4 = gc[h];
                if (q < nq) {
                    float4* rawp = reinterpret_cast<float4*>(rowp + 5 * nc) + 2 * q;
                    rawp[0] = a4; rawp[1] = c4;
                }
                if (4 * q + 1 >= n) { a4.z = 1.f; a4.w = 0.f; }
                if (4 * q + 2 >= n) { c4.x = 1.f; c4.y = 0.f; }
                if (4 * q + 3 >= n) { c4.z = 1.f; c4.w = 0.f; }
                float Pab[5], Pcd[5];
                loo5_pair_coefs(a4.x, a4.y, third_gl(a4.x, a4.y), a4.z, a4.w, third_gl(a4.z, a4.w), Pab);
                loo5_pair_coefs(c4.x, c4.y, third_gl(c4.x, c4.y), c4.z, c4.w, third_gl(c4.z, c4.w), Pcd);
#pragma unroll
                for (int j = 0; j < 5; ++j) P[h][j] = pack2(Pab[j], Pcd[j]);
            }
            ulonglong2* dst = rowp + 5 * c;
            ulonglong2 v;
            v.x = P[0][0]; v.y = P[0][1]; dst[0] = v;
            v.x = P[0][2]; v.y = P[0][3]; dst[1] = v;
            v.x = P[0][4]; v.y = P[1][0]; dst[2] = v;
            v.x = P[1][1]; v.y = P[1][2]; dst[3] = v;
            v.x = P[1][3]; v.y = P[1][4]; dst[4] = v;
            if (c == 0 && (5 * nc + 2 * nq) != ru) rowp[ru - 1] = make_ulonglong2(0ull, 0ull);   // the pad unit: defined bytes for the copy
        }
        __syncthreads();
        {
            const int total = rows * ru;
            ulonglong2* out = PK + s0 * (long)ru;
            for (int u = t; u < total; u += blockDim.x) out[u] = tile[u];
        }
        __syncthreads();
    }
}

struct Loo5Coef { f32x2 A, B, AH, B43, BH; float a, b; };
__device__ __forceinline__ Loo5Coef loo5_coef(float f) {
    Loo5Coef c;
    const float om = 1.0f - f;
    const float r = fast_rcp(f * om);                   // one reciprocal for both ratios (the XU pipe is co-critical here)
    c.a = om * om * r;
    c.b = f * f * r;
    const float ah = 0.5f * c.a, b43 = 1.33333337f * c.b, bh = 1.5f * c.b;
    c.A = pack2(c.a, c.a); c.B = pack2(c.b, c.b); c.AH = pack2(ah, ah); c.B43 = pack2(b43, b43); c.BH = pack2(bh, bh);
    return c;
}
__device__ __forceinline__ float loo5_own(float g0, float g1, float g2, const Loo5Coef& c) {
    const float num = fmaf(g2, c.b, g1);
    const float den = fmaf(g0, c.a, g1 + num);
    return num * fast_rcp(den);
}
__device__ __forceinline__ void loo5_quad(f32x2 P0, f32x2 P1, f32x2 P2, f32x2 P3, f32x2 P4, const Loo5Coef (&c)[4], f32x2 (&acc)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const f32x2 u = ffma2(P0, c[k].A, P1);
        const f32x2 v = ffma2(P4, c[k].B, P3);
        const f32x2 D = ffma2(v, c[k].B, ffma2(u, c[k].A, P2));
        const f32x2 t = ffma2(P4, c[k].B43, P3);
        const f32x2 N = ffma2(t, c[k].BH, ffma2(P1, c[k].AH, P2));
        // the two quotients are accumulated with scalar FFMAs: packing the two reciprocals into a register pair for
        // one FFMA2 cost a move per reciprocal (12 of the 97 instructions of the cell loop) for the same FMA-pipe time
        const float2 dd = unpack2(D), nn = unpack2(N);
        float2 a = unpack2(acc[k]);
        a.x = fmaf(nn.x, fast_rcp(dd.x), a.x);
        a.y = fmaf(nn.y, fast_rcp(dd.y), a.y);
        acc[k] = pack2(a.x, a.y);
    }
}

// The first pair (low lanes) of a quad only, in scalar arithmetic: the same fused operations on half the lanes.
// For a quad whose second pair is padding (n mod 4 = 1 or 2) this is half the FMA-pipe time and one reciprocal
// instead of two per problem.
__device__ __forceinline__ void loo5_pair_lo(f32x2 P0, f32x2 P1, f32x2 P2, f32x2 P3, f32x2 P4, const Loo5Coef (&c)[4], f32x2 (&acc)[4]) {
    const float p0 = unpack2(P0).x, p1 = unpack2(P1).x, p2 = unpack2(P2).x, p3 = unpack2(P3).x, p4 = unpack2(P4).x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float a = c[k].a, b = c[k].b;
        const float ah = unpack2(c[k].AH).x, b43 = unpack2(c[k].B43).x, bh = unpack2(c[k].BH).x;
        const float u = fmaf(p0, a, p1);
        const float v = fmaf(p4, b, p3);
        const float D = fmaf(v, b, fmaf(u, a, p2));
        const float t = fmaf(p4, b43, p3);
        const float N = fmaf(t, bh, fmaf(p1, ah, p2));
        float2 ac = unpack2(acc[k]);
        ac.x = fmaf(N, fast_rcp(D), ac.x);
        acc[k] = pack2(ac.x, ac.y);
    }
}

// Pipeline: a ring of kLoo5Stages row groups (rows_per_pass site rows each) guarded by full/empty
// mbarrier pairs - no block-wide barrier in the loop.  Every warp waits for "full", computes its rows
// and arrives on "empty"; one lane of warp 0 additionally refills a slot with ONE TMA bulk copy (the rows
// of a group are contiguous in the packed array) as soon as all warps have released it, kLoo5Stages-1
// groups ahead of use.  (One copy per row from warp 0 made that warp 50 % slower than the others, and
// every other warp then waited for it at "full": 23 % of all stall samples.)
constexpr int kLoo5MaxStages = 6;
// Geometry: <512, 1> - one block of 16 warps per SM at 116 registers, ring of up to 190 KB - is the default; <256, 2>
// (two blocks of 8 warps, the same 116 registers) runs 3..7 % slower and is kept for option loo_block.  Measured and dropped in
// round 2, each on the same box as its control (0.554 ms per launch at 1M x 50 for <256, 2>): three blocks of 256 at 80
// registers with a ring small enough for three 0.670 ms; two blocks of 320 at 96 registers 0.620 ms; two problems
// per pass over the shared-memory row (half the coefficient registers) 1.1 ms.  The 97-instruction cell loop is the same
// in all of them: registers for its eight independent chains matter more than resident warps.
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
loo_em_step5_kernel(const ulonglong2* __restrict__ PK, int ldg, long M,
                    int col0, int n, int rows_per_pass,
                    float* __restrict__ F, int ldf,              // [M][ldf], in place
                    const int* __restrict__ active,              // [ldg]
                    const unsigned char* __restrict__ mask,      // [M][ldg] or null
                    double* __restrict__ partials,               // [gridDim.x][ldg]
                    long ntiles,                                 // groups of rows_per_pass rows
                    int nstages,                                 // ring depth, 2..kLoo5MaxStages
                    int dbg,                                     // experiments: 1 = no restaging (stale tiles), 2 = staging only
                    float* __restrict__ D2)                      // [ldg / 4][M][4] (quad-major) or null: this iteration's squared change per (problem, site)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full[kLoo5MaxStages], empty[kLoo5MaxStages];
    const int nq = (n + 3) >> 2;                                // quads of individuals = threads per site row
    const int nc = (nq + 1) >> 1;                               // cells (two quads) per row
    const int ru = loo5_row_units(n);                           // row length in 16-byte units (odd)
    const int TS = rows_per_pass;
    ulonglong2* pk0 = reinterpret_cast<ulonglong2*>(smem_raw);                            // [S][TS][ru]  packed cells + raw pairs
    float4* red = reinterpret_cast<float4*>(pk0 + nstages * (size_t)TS * ru);         // [blockDim.x]

    const int t = threadIdx.x, lane = t & 31;
    const int ti = t % nq, r = t / nq;                          // this thread's quad (problems 4ti..4ti+3) and row
    const bool worker = t < rows_per_pass * nq;
    const int c0 = col0 + 4 * ti;
    unsigned act = 0;                                           // bit k: problem 4ti+k exists and is still iterating
    if (worker) {
        const int4 a4 = *reinterpret_cast<const int4*>(&active[c0]);
        act = (a4.x != 0 ? 1u : 0u) | (a4.y != 0 ? 2u : 0u) | (a4.z != 0 ? 4u : 0u) | (a4.w != 0 ? 8u : 0u);
        if (4 * ti + 1 >= n) act &= 1u;
        if (4 * ti + 2 >= n) act &= 3u;
        if (4 * ti + 3 >= n) act &= 7u;
    }
    // a launch that finds every problem of the population frozen (the host queues one iteration ahead of the
    // decisions it reads back) has nothing to stage or compute
    if (__syncthreads_or(act != 0u) == 0) {
        for (int p = t; p < n; p += blockDim.x) partials[(long)blockIdx.x * ldg + col0 + p] = 0.0;
        return;
    }
    const float inv_div = 1.0f / (float)(n - 1);
    const bool last_pair_only = ((n - 1) & 3) < 2;              // n mod 4 is 1 or 2: the last quad holds one real pair
    float ssq[4] = {0.f, 0.f, 0.f, 0.f};

    if (t == 0) {
        for (int s = 0; s < nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], blockDim.x >> 5); }
        mbar_fence_init();
    }
    __syncthreads();
    // thread 0: one TMA bulk copy for the whole group `tl` into ring slot `slot`
    auto issue = [&](long tl, int slot) {
        if (tl < ntiles && t == 0) {
            const long s0 = tl * TS;
            const int rows = (int)min((long)TS, M - s0);
            const unsigned bytes = (unsigned)rows * (unsigned)ru * 16u;
            mbar_expect_tx(&full[slot], bytes);
            bulk_g2s(pk0 + (size_t)slot * TS * ru, PK + s0 * (long)ru, bytes, &full[slot]);
        }
    };
    if (t < 32) {
        for (int s = 0; s < nstages; ++s) issue(blockIdx.x + (long)s * gridDim.x, s);
    }

    // this thread's f quad of the next group is fetched while the current one computes
    unsigned ok_next;
    float4 f_next;
    auto fetch = [&](long tl) {
        const long s = tl * TS + r;
        ok_next = (tl < ntiles && s < M) ? act : 0u;
        if (ok_next && mask) {
            const uchar4 mk = *reinterpret_cast<const uchar4*>(&mask[s * (long)ldg + c0]);
            ok_next &= (mk.x ? 1u : 0u) | (mk.y ? 2u : 0u) | (mk.z ? 4u : 0u) | (mk.w ? 8u : 0u);
        }
        f_next = make_float4(0.25f, 0.25f, 0.25f, 0.25f);
        if (ok_next) f_next = *reinterpret_cast<const float4*>(&F[s * (long)ldf + c0]);
    };
    fetch(blockIdx.x);
    int slot = 0;
    unsigned phase = 0;
#pragma unroll 1
    for (long tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const unsigned okp = ok_next;
        const float4 fq = f_next;
        fetch(tl + gridDim.x);
        const bool staged = dbg != 1 || tl < blockIdx.x + (long)nstages * gridDim.x;
        if (staged) mbar_wait(&full[slot], phase);              // this group has landed
        if (okp && dbg != 2) {
            const float fin[4] = {fq.x, fq.y, fq.z, fq.w};
            const ulonglong2* row = pk0 + ((size_t)slot * TS + r) * ru;
            const int nfull = nq >> 1;                          // cells with both quads
            float fo[4], dq[4];
            Loo5Coef c[4];
            f32x2 acc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { c[k] = loo5_coef(fin[k]); acc[k] = 0ull; }
#pragma unroll 1
            for (int cc = 0; cc < nfull; ++cc) {
                const ulonglong2 v0 = row[5 * cc], v1 = row[5 * cc + 1], v2 = row[5 * cc + 2], v3 = row[5 * cc + 3], v4 = row[5 * cc + 4];
                loo5_quad(v0.x, v0.y, v1.x, v1.y, v2.x, c, acc);
                loo5_quad(v2.y, v3.x, v3.y, v4.x, v4.y, c, acc);
            }
            if (nq & 1) {
                const ulonglong2 v0 = row[5 * nfull], v1 = row[5 * nfull + 1], v2 = row[5 * nfull + 2];
                if (last_pair_only) loo5_pair_lo(v0.x, v0.y, v1.x, v1.y, v2.x, c, acc);   // the quad's second pair is padding
                else loo5_quad(v0.x, v0.y, v1.x, v1.y, v2.x, c, acc);
            }
            // own terms: the four members of quad ti, from the raw pairs at the end of the row (loaded after the cell
            // loop: eight fewer registers live across it)
            const float4* rawrow = reinterpret_cast<const float4*>(row + 5 * nc);
            const float4 ga = rawrow[2 * ti], gc = rawrow[2 * ti + 1];
            const float og0[4] = {ga.x, ga.z, gc.x, gc.z}, og1[4] = {ga.y, ga.w, gc.y, gc.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float own = loo5_own(og0[k], og1[k], third_gl(og0[k], og1[k]), c[k]);
                const float2 a = unpack2(acc[k]);
                float fn = ((a.x + a.y) - own) * inv_div;
                if (fn < 1e-12f) fn = 1e-12f;                   // comparisons are false for NaN: NaN survives
                if (fn > 0.99999994f) fn = 0.99999994f;
                fo[k] = fin[k]; dq[k] = 0.f;                    // frozen / masked problems keep their value
                if (okp & (1u << k)) {                          // (v1 - v2) * (v1 - v2) as rmse1d forms it (emMAF_cy.pyx:31): no FMA
                    const float d = __fsub_rn(fn, fin[k]);
                    dq[k] = __fmul_rn(d, d);
                    fo[k] = fn;
                }
                ssq[k] = __fadd_rn(ssq[k], dq[k]);
            }
            *reinterpret_cast<float4*>(&F[(tl * TS + r) * (long)ldf + c0]) = make_float4(fo[0], fo[1], fo[2], fo[3]);
            // quad-major [ldg / 4][M][4]: one 16-byte store per thread, the sites of a quad contiguous (a purely problem-major
            // layout cost four scattered 4-byte stores per thread: +14 % on this kernel at 2.5 M sites)
            if (D2) *reinterpret_cast<float4*>(&D2[((long)(c0 >> 2) * M + (tl * TS + r)) * 4]) = make_float4(dq[0], dq[1], dq[2], dq[3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);               // this warp is done with the slot
        if (t < 32) {                                           // warp 0 refills it once every warp has released it
            // (refilling the PREVIOUS group's slot instead, so that warp 0 never waits here, measured 0.6 .. 1.3 % slower:
            // one group less in flight costs more than the wait)
            const long nxt = tl + (long)nstages * gridDim.x;
            if (nxt < ntiles && dbg != 1) {
                mbar_wait(&empty[slot], phase);
                issue(nxt, slot);
            }
        }
        if (++slot == nstages) { slot = 0; phase ^= 1u; }
    }
    __syncthreads();
    red[t] = make_float4(ssq[0], ssq[1], ssq[2], ssq[3]);
    __syncthreads();
    for (int p = t; p < n; p += blockDim.x) {                   // problem p = member (p & 3) of quad p / 4 (n may exceed the block)
        double v = 0.0;
        for (int q = 0; q < rows_per_pass; ++q) {
            const float4 x = red[q * nq + (p >> 2)];
            const int k = p & 3;
            v += (double)(k == 0 ? x.x : k == 1 ? x.y : k == 2 ? x.z : x.w);
        }
        partials[(long)blockIdx.x * ldg + col0 + p] = v;
    }
}

constexpr int kFisherQ = 8;   // 16-byte loads in flight per thread in the register-tile row kernels (loo_first, fisher2)
// ---------------------------------------------------------------------------------------
// loo_first: the FIRST leave-one-out EM iteration of one population.  Every problem starts from the same
// f = 0.25 (emMAF.py:17), so the n posterior terms of a site are shared by all n problems: one pass of n
// evaluations per site, f_j = (S - own_j) / (n - 1), instead of the n^2 of a general iteration - a whole
// loo_em_step launch per population saved.  Same thread mapping as fisher2 (TPR adjacent lanes own a site row,
// straight 16-byte loads from G, shuffle row sum, register accumulators for the per-problem squared changes);
// same clamp, mask and `active` semantics and the same partials layout as the step kernels, whose grid it uses.
// ---------------------------------------------------------------------------------------
// (two blocks of 256 per SM at 91 registers; forcing three or four - 80 / 64 registers, ~100 / ~170 bytes of spills -
// measured 0.325 / 0.338 ms against 0.287 ms per launch at 1M x 50)
template <int TPR>
__global__ void __launch_bounds__(256)
loo_first_kernel(const float2* __restrict__ G, int ldg, long M, int col0, int n,
                 float* __restrict__ F, int ldf,
                 const int* __restrict__ active,            // [ldg]
                 const unsigned char* __restrict__ mask,    // [M][ldg] or null
                 double* __restrict__ partials,             // [gridDim.x][ldg]
                 float* __restrict__ D2)                    // [ldg / 4][M][4] (quad-major) or null: squared change per (problem, site)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* red = reinterpret_cast<float*>(smem_raw);      // [R][2 * TPR * kFisherQ]
    constexpr int R = 256 / TPR;
    constexpr int RW = 2 * TPR * kFisherQ;
    const int t = threadIdx.x, r = t / TPR, h = t % TPR;
    const int cpr = (n + 1) >> 1;
    const long ntiles = (M + R - 1) / R;
    const float f0 = 0.25f;
    const Loo5Coef c = loo5_coef(f0);
    const float inv_div = 1.0f / (float)(n - 1);
    unsigned act = 0;                                     // bit 2j / 2j+1: this thread's individuals (2q, 2q+1), q = h + j TPR, iterate
#pragma unroll
    for (int j = 0; j < kFisherQ; ++j) {
        const int q = h + j * TPR;
        if (q < cpr) {
            if (active[col0 + 2 * q]) act |= 1u << (2 * j);
            if (2 * q + 1 < n && active[col0 + 2 * q + 1]) act |= 2u << (2 * j);
        }
    }
    float sa[kFisherQ], sb[kFisherQ];
#pragma unroll
    for (int j = 0; j < kFisherQ; ++j) { sa[j] = 0.f; sb[j] = 0.f; }
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long s = tile * R + r;
        const bool live = s < M;
        float4 v[kFisherQ];
        const float4* row = reinterpret_cast<const float4*>(G + (live ? s : 0) * (long)ldg + col0);
#pragma unroll
        for (int j = 0; j < kFisherQ; ++j) {
            const int q = h + j * TPR;
            v[j] = make_float4(1.f, 0.f, 1.f, 0.f);         // (1,0,0): an exactly zero term
            if (live && q < cpr) v[j] = ld_stream4(row + q);
        }
        float pa[kFisherQ], pb[kFisherQ];
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < kFisherQ; ++j) {
            const int q = h + j * TPR;
            pa[j] = 0.f; pb[j] = 0.f;
            if (live && q < cpr) {
                pa[j] = loo5_own(v[j].x, v[j].y, third_gl(v[j].x, v[j].y), c);
                if (2 * q + 1 < n) pb[j] = loo5_own(v[j].z, v[j].w, third_gl(v[j].z, v[j].w), c);
            }
            sum += pa[j] + pb[j];
        }
#pragma unroll
        for (int o = 1; o < TPR; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (live) {
#pragma unroll
            for (int j = 0; j < kFisherQ; ++j) {
                const int q = h + j * TPR;
                if (q < cpr) {
                    unsigned ok = (act >> (2 * j)) & 3u;
                    if (ok && mask) {
                        const uchar2 mk = *reinterpret_cast<const uchar2*>(&mask[s * (long)ldg + col0 + 2 * q]);
                        ok &= (mk.x ? 1u : 0u) | (mk.y ? 2u : 0u);
                    }
                    float fa = (sum - pa[j]) * inv_div, fb = (sum - pb[j]) * inv_div;
                    if (fa < 1e-12f) fa = 1e-12f;           // comparisons are false for NaN: NaN survives
                    if (fa > 0.99999994f) fa = 0.99999994f;
                    if (fb < 1e-12f) fb = 1e-12f;
                    if (fb > 0.99999994f) fb = 0.99999994f;
                    float2 out = make_float2(f0, f0);       // frozen / masked problems and the pad keep the start value
                    float2 dq = make_float2(0.f, 0.f);
                    if (ok & 1u) { const float d = __fsub_rn(fa, f0); dq.x = __fmul_rn(d, d); sa[j] = __fadd_rn(sa[j], dq.x); out.x = fa; }
                    if (ok & 2u) { const float d = __fsub_rn(fb, f0); dq.y = __fmul_rn(d, d); sb[j] = __fadd_rn(sb[j], dq.y); out.y = fb; }
                    if (ok) {
                        *reinterpret_cast<float2*>(&F[s * (long)ldf + col0 + 2 * q]) = out;
                        if (D2) { const int p = col0 + 2 * q; *reinterpret_cast<float2*>(&D2[((long)(p >> 2) * M + s) * 4 + (p & 3)]) = dq; }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kFisherQ; ++j) {
        const int q = h + j * TPR;
        red[r * RW + 2 * q] = sa[j];
        red[r * RW + 2 * q + 1] = sb[j];
    }
    __syncthreads();
    for (int j = t; j < n; j += blockDim.x) {
        double vsum = 0.0;
        for (int rr = 0; rr < R; ++rr) vsum += (double)red[rr * RW + j];
        partials[(long)blockIdx.x * ldg + col0 + j] = vsum;
    }
}

// ssq[p] = sum over blocks of partials[block][p], in a fixed order: one warp per problem, lane l adds blocks
// l, l+32, ... in sequence, then a shuffle tree - the same bits for a given number of blocks, and a few
// microseconds instead of one thread walking every block (this kernel sits between two EM iterations).
__global__ void __launch_bounds__(256)
em_ssq_reduce_kernel(const double* __restrict__ partials, int nblocks, int np, int ld, double* __restrict__ ssq)
{
    const int p = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= np) return;
    double v = 0.0;
    int b = lane;
    for (; b + 96 < nblocks; b += 128) {                        // 4 independent loads in flight, fixed add order
        const double x0 = partials[(long)b * ld + p], x1 = partials[(long)(b + 32) * ld + p];
        const double x2 = partials[(long)(b + 64) * ld + p], x3 = partials[(long)(b + 96) * ld + p];
        v += x0; v += x1; v += x2; v += x3;
    }
    for (; b < nblocks; b += 32) v += partials[(long)b * ld + p];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) ssq[p] = v;
}

// ssq[p] = sum over ranks, IN RANK ORDER, of gathered[rank][p] (the output of an NCCL all-gather of every rank's
// local sums): every rank computes the same bits, and they do not depend on how NCCL would have reduced.
__global__ void em_rank_sum_kernel(const double* __restrict__ gathered, int world, int np, double* __restrict__ ssq)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
        double v = 0.0;
        for (int r = 0; r < world; ++r) v += gathered[(long)r * np + p];
        ssq[p] = v;
    }
}

// ---------------------------------------------------------------------------------------
// Sequential float32 summation, exactly, at warp speed.
//
// The reference's stop rule adds the squared changes of ALL sites into ONE float32 accumulator, left to right
// (rmse1d, emMAF_cy.pyx:29-32).  At millions of sites that sum differs from the exact one by up to
// n 2^-24 relative (-0.4 % measured at 5 M sites, -2.3 % at 20 M: addends below half an ulp of the accumulator
// vanish), enough to move the stop iteration.  Its bits depend on the order, so it cannot be a tree reduction - but
// it does not have to be one addition at a time either: while the accumulator res = S u stays inside one binade
// (u = its ulp, S a 24-bit integer), adding x >= 0 gives RN(S u + x) = (S + RN(x / u)) u unless x / u lies exactly
// half-way between two integers (then the tie goes to the even S).  So over a block of 256 addends with no tie and
// no binade crossing the accumulator simply advances by the INTEGER sum of the RN(x_i / u) - order-free, one REDUX
// per block.  Ties (about one per 2^20 addends here) and binade crossings (~25 per sum) are detected, and that block
// is then added one element at a time in order with real float32 additions.  The result is bit-identical to the
// serial loop for every input (negative, NaN and infinite addends take the in-order path).
// get(i): the i-th addend, 0 <= i < n; element order inside a chunk of 256 is i = base + 32 j + lane.
// ---------------------------------------------------------------------------------------
// One BLOCK per sum (kSeqWarps warps): a single warp would be bound by load latency (256 addends per ~1 us round trip,
// 4 ms per million).  Per round every warp takes one chunk of 256 addends - the next two rounds' are already in flight -
// leaves them in shared memory and works out the chunk's integer advance Q for the binade the accumulator is in at the
// start of the round.  Warp 0 then folds the round: all 32 chunks at once when none has a tie and the round stays in the
// binade, else chunk by chunk IN ORDER (when every thread folded every chunk redundantly, and later when warp 0 walked the
// chunks one by one every round, the fold itself was the cost: 2.5 and 0.8 ms per million addends).  A chunk with a tie, or one
// that would leave the binade, is added one element at a time, in order, from shared memory; when that moved the
// accumulator to another binade the remaining chunks of the round are re-derived for the new one.
constexpr int kSeqWarps = 32;
__device__ __forceinline__ void seq_chunk_q(const float (&x)[8], int eb, unsigned& Q, bool& anybad)
{
    bool bad = !(eb >= 30 && eb <= 250);
    int qsum = 0;
    if (!bad) {
        const float scale = __uint_as_float((unsigned)(277 - eb) << 23);   // 1 / ulp(res), a power of two: x * scale is exact
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float t = x[j] * scale;
            bad = bad || !(t >= 0.0f && t < 32768.0f);             // also NaN; 2^15: the Q of a whole round (8192 addends) fits 32 bits
            bad = bad || (t - floorf(t) == 0.5f);                  // a tie: its rounding depends on the parity of S
            qsum += __float2int_rn(t);
        }
    }
    anybad = __any_sync(0xffffffffu, bad);
    Q = (unsigned)__reduce_add_sync(0xffffffffu, anybad ? 0 : qsum);
}
template <class Get>
__device__ __forceinline__ float block_seqsum32(float res, long n, Get get)
{
    __shared__ float sh_x[kSeqWarps][256];
    __shared__ unsigned sh_q[kSeqWarps];
    __shared__ int sh_bad[kSeqWarps];
    __shared__ float sh_res;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long per_round = 256L * kSeqWarps;
    float x[8], xn[8], xnn[8];                               // this round, and the next two already in flight
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const long i = (long)warp * 256 + 32 * j + lane;
        x[j] = i < n ? get(i) : 0.0f;
        xn[j] = i + per_round < n ? get(i + per_round) : 0.0f;
    }
    for (long r0 = 0; r0 < n; r0 += per_round) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const long i = r0 + 2 * per_round + (long)warp * 256 + 32 * j + lane;
            xnn[j] = i < n ? get(i) : 0.0f;                  // + 0.0f never changes a float32 accumulator that started at +0
        }
        const int eb0 = (int)(__float_as_uint(res) >> 23);   // sign bit included: a negative accumulator fails the range test
        {
            unsigned Q; bool anybad;
            seq_chunk_q(x, eb0, Q, anybad);
#pragma unroll
            for (int j = 0; j < 8; ++j) sh_x[warp][32 * j + lane] = x[j];
            if (lane == 0) { sh_q[warp] = Q; sh_bad[warp] = anybad ? 1 : 0; }
        }
        __syncthreads();
        if (warp == 0) {
            // the fold.  Whole round at once when no chunk has a tie and the round stays inside the binade (then so does
            // every prefix: the advances are non-negative) - the usual case: one vote, one REDUX.
            const unsigned bits = __float_as_uint(res);
            const unsigned S0 = (bits & 0x007fffffu) | 0x00800000u;
            const bool anyb = __any_sync(0xffffffffu, sh_bad[lane] != 0);
            const unsigned tot = (unsigned)__reduce_add_sync(0xffffffffu, sh_q[lane]);
            int eb = eb0;
            if (!anyb && S0 + tot < 0x01000000u) {
                res = __uint_as_float(((unsigned)eb << 23) | ((S0 + tot) & 0x007fffffu));
            } else
            for (int w = 0; w < kSeqWarps; ++w) {            // chunk by chunk, in order
                if (r0 + (long)w * 256 >= n) break;
                const unsigned b = __float_as_uint(res);
                const unsigned S = (b & 0x007fffffu) | 0x00800000u;
                if (!sh_bad[w] && (int)(b >> 23) == eb && S + sh_q[w] < 0x01000000u) {
                    res = __uint_as_float(((unsigned)eb << 23) | ((S + sh_q[w]) & 0x007fffffu));
                    continue;
                }
                float r = res;                               // in order, one real float32 addition per element
#pragma unroll 1
                for (int j = 0; j < 8; ++j) {
                    const float xv = sh_x[w][32 * j + lane];
#pragma unroll 8
                    for (int l = 0; l < 32; ++l) r = __fadd_rn(r, __shfl_sync(0xffffffffu, xv, l));
                }
                res = r;
                const int eb_new = (int)(__float_as_uint(res) >> 23);
                if (eb_new != eb) {                          // the later chunks were scaled for the old binade: re-derive them
                    eb = eb_new;
                    for (int v = w + 1; v < kSeqWarps; ++v) {
                        float xv[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) xv[j] = sh_x[v][32 * j + lane];
                        unsigned Q; bool anybad;
                        seq_chunk_q(xv, eb, Q, anybad);
                        if (lane == 0) { sh_q[v] = Q; sh_bad[v] = anybad ? 1 : 0; }
                    }
                    __syncwarp();
                }
            }
            if (lane == 0) sh_res = res;
        }
        __syncthreads();
        res = sh_res;
        __syncthreads();                                     // sh_x / sh_q / sh_res are rewritten next round
#pragma unroll
        for (int j = 0; j < 8; ++j) { x[j] = xn[j]; xn[j] = xnn[j]; }
    }
    return res;
}

// Stop-rule tie-break, part 1: one block per problem.  A problem is UNCERTAIN when the RMSE from the exact (FP64)
// sum of squared changes lies within `band` (relative) of the tolerance - band = the worst-case distance between
// the exact sum and the reference's sequential float32 sum for that many addends (or the caller's override;
// band < 0: every active problem).  For those, the float32 sum is reproduced (block_seqsum32) from the squared changes the step
// kernels left in D2 (quad-major; zero where a site is masked out), starting from carry_in[p]
// (site-sharded runs chain the ranks in site order).  em_decide_kernel then decides on serial[p].
// Half-width (relative, on the RMSE) of the band around the tolerance inside which the FP64 sum is not trusted to
// decide like the reference's sequential float32 sum.  Default: 8x the measured bias of that sum - SURVEY.md 7.1c:
// -0.04 % of the sum at 1 M addends, -0.4 % at 5 M, -2.3 % at 20 M, i.e. 4e-4 (n / 1e6)^1.35, half of that on the square
// root - but never more than the rigorous bound gamma_n / 2 of recursive summation (Higham), plus 1e-4 for the FP32
// per-thread partials of the FP64 sum.  Overrides: -1 = every active problem (no band at all), -2 = the rigorous bound
// alone (every check once n >= 2^23), > 0 = that relative half-width.
__host__ __device__ __forceinline__ double em_band(double cnt, double band_override)
{
    if (band_override == -1.0 || band_override > 0.0) return band_override;
    const double ku = cnt * 5.9604644775390625e-08;        // n 2^-24
    const double rigorous = ku >= 0.5 ? -1.0 : 0.5 * ku / (1.0 - ku) + 1e-4;
    if (band_override == -2.0) return rigorous;
    const double measured = 8.0 * 0.5 * 4e-4 * pow(cnt * 1e-6, 1.35) + 1e-4;
    return (rigorous >= 0.0 && rigorous < measured) ? rigorous : measured;
}
__global__ void __launch_bounds__(kSeqWarps * 32)
em_resolve_kernel(const double* __restrict__ ssq, const double* __restrict__ count, double count_all, int np, double tole,
                  double band_override, const int* __restrict__ active,
                  const float* __restrict__ D2, long M,        // [np / 4][M][4] quad-major (offset to the first problem, a multiple of 4, applied by the caller)
                  const float* __restrict__ carry_in,       // [np] or null (first rank)
                  float* __restrict__ serial,               // [np] out: the running float32 sum after this rank's sites
                  int* __restrict__ uncertain,              // [np] out
                  int* __restrict__ any_uncertain = nullptr)   // optional: set to 1 when some problem is uncertain
{
    for (int p = blockIdx.x; p < np; p += gridDim.x) {      // block-uniform
        int unc = 0;
        if (active[p]) {
            const double cnt = count ? count[p] : count_all;
            float res = (float)ssq[p];
            res = res / (float)cnt;
            const double diff = sqrt((double)res);
            const double band = em_band(cnt, band_override);
            unc = (band < 0.0 || fabs(diff - tole) <= band * tole) ? 1 : 0;
        }
        if (unc && D2) {                                     // D2 == null: flags only
            const float* col = D2 + (long)(p >> 2) * M * 4 + (p & 3);   // quad-major: this problem's addends are 16 bytes apart
            const float r = block_seqsum32(carry_in ? carry_in[p] : 0.0f, M, [&](long i) { return __ldg(col + 4 * i); });
            if (threadIdx.x == 0) serial[p] = r;
        }
        if (threadIdx.x == 0) { uncertain[p] = unc; if (unc && any_uncertain) *any_uncertain = 1; }
    }
}

// the same float32 sum for a contiguous pair of state vectors (population EM: the per-iteration history): x_i = (cur_i - prev_i)^2
__global__ void __launch_bounds__(kSeqWarps * 32)
seqsum_pair_kernel(const float* __restrict__ cur, const float* __restrict__ prev, long M, const float* __restrict__ carry_in,
                   float* __restrict__ out)
{
    const float r = block_seqsum32(carry_in ? carry_in[0] : 0.0f, M, [&](long i) { const float d = __fsub_rn(cur[i], prev[i]); return __fmul_rn(d, d); });
    if (threadIdx.x == 0) out[0] = r;
}
__global__ void __launch_bounds__(kSeqWarps * 32)
seqsum_vec_kernel(const float* __restrict__ x, long n, const float* __restrict__ carry_in, float* __restrict__ out)
{
    const float r = block_seqsum32(carry_in ? carry_in[0] : 0.0f, n, [&](long i) { return x[i]; });
    if (threadIdx.x == 0) out[0] = r;
}

__global__ void rank_sum_i64_kernel(const long long* __restrict__ gathered, int world, int np, long long* __restrict__ out)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
        long long v = 0;
        for (int r = 0; r < world; ++r) v += gathered[(long)r * np + p];
        out[p] = v;
    }
}

// Stop rule of emMAF.py:21-25 with rmse1d's float divide / double sqrt (emMAF_cy.pyx:32-33).
// count[p] = number of sites in problem p's sum.  Single block.  result (mapped pinned host memory):
// [0] = problems still active, [1 + p] = the updated flag of problem p.
__global__ void em_decide_kernel(const double* __restrict__ ssq, const double* __restrict__ count, double count_all,
                                 int np, double tole, int iteration,
                                 int* __restrict__ active, int* __restrict__ iters, int* __restrict__ result,
                                 const int* __restrict__ uncertain = nullptr,   // [np] em_resolve_kernel's flags, or null: FP64 sums only
                                 const float* __restrict__ serial = nullptr,    // [np] the sequential float32 sums of the uncertain problems
                                 int near_slot = -1,                            // result[near_slot]: 1 if a problem still active is within near_factor x of the tolerance
                                 double near_factor = 30.0)
{
    __shared__ int sh_cnt, sh_near, sh_missed;
    if (threadIdx.x == 0) { sh_cnt = 0; sh_near = 0; sh_missed = 0; }
    __syncthreads();
    int mine = 0, near = 0;
    for (int p = threadIdx.x; p < np; p += blockDim.x) {
        int a = active[p];
        if (a) {
            double cnt = count ? count[p] : count_all;
            const bool unc = uncertain && uncertain[p];
            if (unc && !serial) atomicOr(&sh_missed, 1);    // needed the sequential sum, but the rank chain was not queued: the caller restarts
            float res = (unc && serial) ? serial[p] : (float)ssq[p];
            res = res / (float)cnt;
            double diff = sqrt((double)res);
            if (diff < tole) { a = 0; active[p] = 0; iters[p] = iteration; }
            else { ++mine; if (diff <= near_factor * tole) near = 1; }
        }
        result[1 + p] = a;
    }
    if (mine) atomicAdd(&sh_cnt, mine);
    if (near) atomicOr(&sh_near, 1);
    __syncthreads();
    if (threadIdx.x == 0) { result[0] = sh_cnt; if (near_slot >= 0) { result[near_slot] = sh_near; result[near_slot + 1] = sh_missed; } }
}

__global__ void fill_kernel(float* __restrict__ p, long n, float v)
{
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) p[e] = v;
}

// F[s][cols[j]] <- v: the start value of the padding columns only (every real column is written by loo_first)
__global__ void fill_cols_kernel(float* __restrict__ F, int ld, long M, const int* __restrict__ cols, int ncols, float v)
{
    const long total = M * (long)ncols;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long s = e / ncols;
        F[s * (long)ld + cols[(int)(e - s * ncols)]] = v;
    }
}

// F[s][c] <- row[c] for c < ncols: start state of the leave-one-out EM
__global__ void bcast_row_kernel(float* __restrict__ F, int ld, int ncols, long M, const float* __restrict__ row)
{
    long total = M * (long)ncols;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long s = e / ncols;
        int c = (int)(e - s * ncols);
        F[s * (long)ld + c] = row[c];
    }
}

// F[s][c] <- clamp(F[s][c], lo[c], hi[c]) for c < ncols (WGSassign.py:236-240, glassy.py:80-85)
__global__ void clip_cols_kernel(float* __restrict__ F, int ld, int ncols, long M,
                                 const float* __restrict__ lo, const float* __restrict__ hi)
{
    long total = M * (long)ncols;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long s = e / ncols;
        int c = (int)(e - s * ncols);
        float* p = F + s * (long)ld + c;
        float v = *p;
        if (v < lo[c]) v = lo[c];
        if (v > hi[c]) v = hi[c];
        *p = v;
    }
}

// A[s][k] = clamp(FT[k][s], lo[k], hi[k]) (WGSassign.py:236-240): population-major EM state -> [M][K]
__global__ void clip_transpose_kernel(const float* __restrict__ FT, long M, int K, const float* __restrict__ lo,
                                      const float* __restrict__ hi, int do_clip, float* __restrict__ A)
{
    long total = M * (long)K;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long s = e / K;
        int k = (int)(e - s * K);
        float v = FT[(size_t)k * M + s];
        if (do_clip) {
            if (v < lo[k]) v = lo[k];
            if (v > hi[k]) v = hi[k];
        }
        A[e] = v;
    }
}

// dst[s][j] = src[s][cols[j]] (j < nc): column gather between row-major float matrices
// clip_lo / clip_hi (optional, indexed by SOURCE column): clamp on the way, like clip_cols_kernel
__global__ void gather_cols_kernel(const float* __restrict__ src, int lds, const int* __restrict__ cols, int nc,
                                   float* __restrict__ dst, int ldd, int dst_col0, long M,
                                   const float* __restrict__ clip_lo = nullptr, const float* __restrict__ clip_hi = nullptr)
{
    long total = M * (long)nc;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long s = e / nc;
        int j = (int)(e - s * nc);
        int c = cols[j];
        if (c >= 0) {
            float v = src[s * (long)lds + c];
            if (clip_lo) { if (v < clip_lo[c]) v = clip_lo[c]; if (v > clip_hi[c]) v = clip_hi[c]; }
            dst[s * (long)ldd + dst_col0 + j] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------
// fisher: observed Fisher information per (site, population), the effective sample size
// n-tilde, and every individual's own n-tilde sum, in ONE pass over G (fisher_cy.pyx:12-65;
// fisher.py:11-59 makes N+K passes).  Warp per site; lanes stride over each population's
// slab.  HBM-bound: 8 B per (site, individual) in, 8 B per (site, population) out.
// Per-individual sums live in a per-warp private shared-memory row (no atomics), reduced
// over warps then over blocks in fixed order.
// ---------------------------------------------------------------------------------------
// H = ((1-th)^2, 2 th (1-th), th^2) is per (site, population): hoisted out of the individual loop.  The two
// IEEE divisions of the reference (n1/u, n2/u) are one MUFU reciprocal + one Newton step shared by both
// quotients (~1 ulp): the correctly-rounded __fdiv_rn pair was 3/4 of this kernel's instructions and kept an
// HBM-bound pass issue-bound (ncu: 120 instructions per term, 20 % of DRAM bandwidth).
struct FisherH { float h0, h1, h2, th; };
__device__ __forceinline__ FisherH fisher_h(float th, float om) { return FisherH{om * om, 2.0f * th * om, th * th, th}; }
__device__ __forceinline__ float fisher_term(float g0, float g1, const FisherH& H)
{
    float g2 = third_gl(g0, g1);
    float u = fmaf(g0, H.h0, fmaf(g1, H.h1, g2 * H.h2));
    float n1 = 2.0f * ((g0 + g2) - 2.0f * g1);
    float n2 = fmaf(H.th, n1, 2.0f * (g1 - g0));
    float r = fast_rcp(u);
    r = r * fmaf(-u, r, 2.0f);
    float x = n2 * r, y = n1 * r;
    return fmaf(x, x, -y);
}

// Same tile machinery as em_pop_step: grid.y = population, kEmT threads per site row, TMA bulk
// staging.  The per-individual sums are column sums over sites: the 8 rows a warp holds are
// folded with 3 shuffles per individual into a per-warp private shared-memory row (no
// atomics), and warps / blocks are then summed in fixed order.
__global__ void __launch_bounds__(512)
fisher_kernel(const float2* __restrict__ G, int ldg, long M,
              const PopDesc* __restrict__ pops, int K,
              const float* __restrict__ A,                 // [M][K]
              float* __restrict__ f_obs, float* __restrict__ ne_obs,   // [M][K]
              int row16, int accw_ld,
              double* __restrict__ ind_partials)           // [gridDim.x][ldg]
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long mbar[2];
    const int k = blockIdx.y;
    const int R = blockDim.x / kEmT;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31, nwarp = blockDim.x >> 5;
    const int r = t / kEmT, h = t % kEmT;
    const PopDesc pd = pops[k];
    const int cpr = (pd.n + 1) >> 1;
    float4* ring = reinterpret_cast<float4*>(smem_raw);   // [2][R][row16]
    const size_t stage = (size_t)R * row16;
    float* accw = reinterpret_cast<float*>(ring + 2 * stage);   // [nwarp][accw_ld]
    const long ntiles = (M + R - 1) / R;
    float* mine = accw + (size_t)warp * accw_ld;
    for (int c = lane; c < accw_ld; c += 32) mine[c] = 0.f;

    if (t == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_fence_init(); }
    __syncthreads();
    auto issue = [&](long tile, int buf) {
        if (tile < ntiles && t < 32) {
            const long s0 = tile * R;
            const int rows = (int)min((long)R, M - s0);
            if (t == 0) mbar_expect_tx(&mbar[buf], (unsigned)(rows * cpr * 16));
            __syncwarp();
            float4* dst = ring + buf * stage;
            for (int rr = t; rr < rows; rr += 32)
                bulk_g2s(dst + (size_t)rr * row16, G + (s0 + rr) * (long)ldg + pd.col0, (unsigned)(cpr * 16), &mbar[buf]);
        }
    };
    issue(blockIdx.x, 0);
    issue(blockIdx.x + (long)gridDim.x, 1);
    int it = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const long s = tile * R + r;
        const bool live = s < M;
        float th = 0.5f;
        if (live) th = __ldg(&A[s * K + k]);
        const float om = 1.0f - th;
        const float w = live ? 0.5f * th * om : 0.f;
        const FisherH H = fisher_h(th, om);
        mbar_wait(&mbar[buf], (unsigned)((it >> 1) & 1));
        const float4* row = ring + buf * stage + (size_t)r * row16;
        float sum = 0.f;
        for (int q0 = 0; q0 < cpr; q0 += kEmT) {            // warp-uniform trip count: the shuffles below need every lane
            const int q = q0 + h;
            float ta = 0.f, tb = 0.f;
            if (live && q < cpr) {
                float4 v = row[q];
                ta = fisher_term(v.x, v.y, H);
                if (2 * q + 1 < pd.n) tb = fisher_term(v.z, v.w, H);
            }
            sum += ta + tb;
            float ia = ta * w, ib = tb * w;                 // this row's contribution to individuals 2q, 2q+1
#pragma unroll
            for (int o = kEmT; o < 32; o <<= 1) {           // fold the 8 rows of the warp
                ia += __shfl_xor_sync(0xffffffffu, ia, o);
                ib += __shfl_xor_sync(0xffffffffu, ib, o);
            }
            if (lane < kEmT && q < cpr) { mine[2 * q] += ia; mine[2 * q + 1] += ib; }
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        if (live && h == 0) {
            f_obs[s * K + k] = sum;
            ne_obs[s * K + k] = 0.5f * sum * th * om;
        }
        __syncthreads();
        issue(tile + 2 * (long)gridDim.x, buf);
    }
    __syncthreads();
    for (int j = t; j < pd.n; j += blockDim.x) {
        double v = 0.0;
        for (int w2 = 0; w2 < nwarp; ++w2) v += (double)accw[(size_t)w2 * accw_ld + j];
        ind_partials[(long)blockIdx.x * ldg + pd.col0 + j] = v;
    }
}

// ---------------------------------------------------------------------------------------
// fisher2: the same pass without shared-memory staging or block barriers.  The TMA-tile version above
// spent its time at the per-tile __syncthreads (ncu: barrier 2.5 + wait 1.9 stall cycles per issue, 20-30 % of
// HBM) - halving its instruction count did not move it.  Here TPR adjacent lanes own a site row and read it
// straight from global memory, kFisherQ 16-byte loads per thread all in flight before the first use (a row is
// one contiguous, sector-aligned run, so the TPR lanes of one load cover whole sectors); the row sum is a
// shuffle over the TPR lanes, and the per-individual sums are REGISTER accumulators (a thread always serves
// the same 2 x kFisherQ individuals), reduced over the rows of the block once, at the end, in fixed order.
// ---------------------------------------------------------------------------------------
template <int TPR>
__global__ void __launch_bounds__(256)
fisher2_kernel(const float2* __restrict__ G, int ldg, long M,
               const PopDesc* __restrict__ pops, int K,
               const float* __restrict__ A,                 // [M][K]
               float* __restrict__ f_obs, float* __restrict__ ne_obs,   // [M][K]
               double* __restrict__ ind_partials)           // [gridDim.x][ldg]
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* red = reinterpret_cast<float*>(smem_raw);      // [R][2 * TPR * kFisherQ]
    constexpr int R = 256 / TPR;
    constexpr int RW = 2 * TPR * kFisherQ;
    const int k = blockIdx.y;
    const int t = threadIdx.x, r = t / TPR, h = t % TPR;
    const PopDesc pd = pops[k];
    const int cpr = (pd.n + 1) >> 1;                      // 16-byte units (pairs of individuals) per row
    const long ntiles = (M + R - 1) / R;
    float ia[kFisherQ], ib[kFisherQ];
#pragma unroll
    for (int j = 0; j < kFisherQ; ++j) { ia[j] = 0.f; ib[j] = 0.f; }
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long s = tile * R + r;
        const bool live = s < M;
        float4 v[kFisherQ];
        const float4* row = reinterpret_cast<const float4*>(G + (live ? s : 0) * (long)ldg + pd.col0);
#pragma unroll
        for (int j = 0; j < kFisherQ; ++j) {
            const int q = h + j * TPR;
            v[j] = make_float4(1.f, 0.f, 1.f, 0.f);
            if (live && q < cpr) v[j] = ld_stream4(row + q);
        }
        float th = 0.5f;
        if (live) th = __ldg(&A[s * K + k]);
        const float om = 1.0f - th;
        const float w = live ? 0.5f * th * om : 0.f;
        const FisherH H = fisher_h(th, om);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < kFisherQ; ++j) {
            const int q = h + j * TPR;
            float ta = 0.f, tb = 0.f;
            if (live && q < cpr) {
                ta = fisher_term(v[j].x, v[j].y, H);
                if (2 * q + 1 < pd.n) tb = fisher_term(v[j].z, v[j].w, H);
            }
            sum += ta + tb;
            ia[j] = fmaf(ta, w, ia[j]);
            ib[j] = fmaf(tb, w, ib[j]);
        }
#pragma unroll
        for (int o = 1; o < TPR; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (live && h == 0) {
            f_obs[s * K + k] = sum;
            ne_obs[s * K + k] = 0.5f * sum * th * om;
        }
    }
#pragma unroll
    for (int j = 0; j < kFisherQ; ++j) {
        const int q = h + j * TPR;
        red[r * RW + 2 * q] = ia[j];
        red[r * RW + 2 * q + 1] = ib[j];
    }
    __syncthreads();
    for (int j = t; j < pd.n; j += blockDim.x) {
        double vsum = 0.0;
        for (int rr = 0; rr < R; ++rr) vsum += (double)red[rr * RW + j];
        ind_partials[(long)blockIdx.x * ldg + pd.col0 + j] = vsum;
    }
}

// ---------------------------------------------------------------------------------------
// Diagnostics: plain read streams over the resident GL matrix, to separate "what HBM gives for
// this access pattern" from "what a kernel's arithmetic costs" when reading a roofline.
// mode 0: flat 128-bit grid-stride read of the whole matrix; mode 1: one population slab at a
// time (rows of pd.n pairs, 4 KB apart), the pattern of the per-population kernels.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stream_probe_kernel(const float2* __restrict__ G, int ldg, long M, const PopDesc* __restrict__ pops, int K, int mode,
                    float* __restrict__ sink)
{
    float acc = 0.f;
    if (mode == 0) {
        const float4* p = reinterpret_cast<const float4*>(G);
        long n4 = M * (long)ldg / 2;
        for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n4; e += (long)gridDim.x * blockDim.x) {
            float4 v = ld_stream4(p + e);
            acc += v.x + v.y + v.z + v.w;
        }
    } else {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int k = blockIdx.y; k < K; k += gridDim.y) {
            PopDesc pd = pops[k];
            int cpr = (pd.n + 1) >> 1;
            for (long s = (long)blockIdx.x * 8 + warp; s < M; s += (long)gridDim.x * 8) {
                const float4* row = reinterpret_cast<const float4*>(G + s * (long)ldg + pd.col0);
                for (int c = lane; c < cpr; c += 32) { float4 v = ld_stream4(row + c); acc += v.x + v.y + v.z + v.w; }
            }
        }
    }
    if (acc == 123456.789f) sink[0] = acc;               // keeps the loads alive
}

// ---------------------------------------------------------------------------------------
// synthetic data (benchmarks): SURVEY 8d model with a counter-based hash, generated in the
// device layout.  AF of population k at site s is an arcsine-distributed value clipped to
// [0.02, 0.98]; genotype = two Bernoulli draws; depth ~ Poisson(depth) by CDF inversion;
// alt reads = Bernoulli(e | 0.5 | 1-e) per read; GL normalised and rounded to 6 decimals.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(uint64_t h) { return (float)((h >> 40) + 0.5f) * (1.0f / 16777216.0f); }

__device__ __forceinline__ float synth_af(uint64_t seed, long site, int k) {
    float u = u01(mix64(seed ^ mix64((uint64_t)site * 1315423911ull + (uint64_t)k * 2654435761ull + 17)));
    float sn = sinpif(0.5f * u);
    float p = sn * sn;
    return fminf(fmaxf(p, 0.02f), 0.98f);
}

__global__ void synth_kernel(float2* __restrict__ G, uchar2* __restrict__ AD, int ldg, long M,
                             const int* __restrict__ ind_of_col, const int* __restrict__ pop_of_col,
                             uint64_t seed, float depth, long site_offset)
{
    long total = M * (long)ldg;
    const float e = 0.01f;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        long s = idx / ldg;
        int c = (int)(idx - s * ldg);
        int ind = ind_of_col[c];
        float2 gl = make_float2(0.f, 0.f);
        uchar2 ad = make_uchar2(0, 0);
        if (ind >= 0) {
            long gs = site_offset + s;
            float p = synth_af(seed, gs, pop_of_col[c]);
            uint64_t h = mix64(seed ^ mix64((uint64_t)gs * 0x100000001B3ull + (uint64_t)ind));
            int geno = (u01(h) < p) + (u01(mix64(h + 1)) < p);
            // Poisson by inversion
            float u = u01(mix64(h + 2));
            float pk = expf(-depth), cdf = pk;
            int D = 0;
            while (u > cdf && D < 60) { ++D; pk *= depth / D; cdf += pk; }
            float pa = geno == 0 ? e : (geno == 1 ? 0.5f : 1.0f - e);
            int alt = 0;
            for (int rd = 0; rd < D; ++rd) alt += (u01(mix64(h + 3 + rd)) < pa);
            int ref = D - alt;
            double l0 = pow(1.0 - (double)e, ref) * pow((double)e, alt);
            double l1 = pow(0.5, D);
            double l2 = pow(1.0 - (double)e, alt) * pow((double)e, ref);
            double tot = l0 + l1 + l2;
            gl.x = (float)(rint(l0 / tot * 1e6) * 1e-6);
            gl.y = (float)(rint(l1 / tot * 1e6) * 1e-6);
            ad = make_uchar2((unsigned char)ref, (unsigned char)alt);
        }
        G[idx] = gl;
        if (AD) AD[idx] = ad;
    }
}

}  // namespace wgs
