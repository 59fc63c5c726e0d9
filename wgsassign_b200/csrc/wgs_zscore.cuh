// z-score kernels: allele-depth class tally, keep mask, expected / variance moments
// (zscore.py:11-120 and zscore_cy.pyx:10-56; assembled in WGSassign.py:346-381, :425-443).
//
// For every individual the reference (1) groups its sites by the (ref, alt) read-count pair
// and takes the mean GL triple of each group, (2) keeps a group when it is frequent enough
// and every split of its depth is present, (3) keeps a site when its group is kept and its
// GL at the group's most likely genotype is within 0.01 of the group mean, and (4) sums the
// observed log-likelihood and its expectation / variance over all read splits of the site's
// depth.  Steps 1-3 are two pure-Python O(M) loops per individual in the reference (~80 % of
// its z-score wall time); here they are two streaming passes over (GL, AD) for ALL
// individuals at once.
//
// Class tables are dense over (ref, alt) with ref + alt <= kZDepthCap.  Tallies are integers
// (bit-exact).  GL sums are signed 64-bit fixed point at 2^-36 (two 18-bit limbs per value,
// |error| <= 2^-37 per value - members of a class share nearly identical GLs, so a coarser
// grid would bias the class mean), hence order-independent: the same bits whatever the grid
// or the number of GPUs.
#pragma once
#include "wgs_kernels.cuh"

namespace wgs {

constexpr int kZDepthCap = 40;
constexpr int kZClasses = (kZDepthCap + 1) * (kZDepthCap + 2) / 2;   // 861
constexpr int kZHotDepth = 4;
constexpr int kZHot = (kZHotDepth + 1) * (kZHotDepth + 2) / 2;       // 15 classes held in registers (sequential kernel)
constexpr int kZHotDepthX = 3;                                       // exact kernel: 7 accumulators per class, 10 classes
constexpr int kZHotX = (kZHotDepthX + 1) * (kZHotDepthX + 2) / 2;
constexpr float kZLimb = 262144.0f;                                  // 2^18 per limb; sums are in units of 2^-36
constexpr double kZUnit = 1.0 / 68719476736.0;                       // 2^-36

__host__ __device__ inline int zclass_id(int ref, int alt)
{
    int d = ref + alt;
    return d * (d + 1) / 2 + alt;
}

struct ZTally { long long cnt, s0, s1, s2; };

// numpy's `1 - L0 - L1` on float32 scalars (zscore.py:17, :54): two float32 roundings
__device__ __forceinline__ float third_gl_np(float g0, float g1) { return __fsub_rn(__fsub_rn(1.0f, g0), g1); }

// value -> (hi, lo) 18-bit limbs: g = hi * 2^-18 + lo * 2^-36 (+- 2^-37); every step is exact in FP32
__device__ __forceinline__ void fix_limbs(float g, int& hi, int& lo)
{
    float gs = g * kZLimb;
    float fl = floorf(gs);
    hi = (int)fl;
    lo = __float2int_rn((gs - fl) * kZLimb);
}

__device__ __forceinline__ void tally_flush(ZTally* t, int cnt, int h0, int l0, int h1, int l1, int h2, int l2)
{
    if (cnt) {
        atomicAdd((unsigned long long*)&t->cnt, (unsigned long long)(long long)cnt);
        atomicAdd((unsigned long long*)&t->s0, (unsigned long long)(((long long)h0 << 18) + l0));
        atomicAdd((unsigned long long*)&t->s1, (unsigned long long)(((long long)h1 << 18) + l1));
        atomicAdd((unsigned long long*)&t->s2, (unsigned long long)(((long long)h2 << 18) + l2));
    }
}

// ---------------------------------------------------------------------------------------
// ztally: per (individual, class) site count and fixed-point GL sums (zscore.py:13-22).
// Thread = individual, warp = 32 consecutive columns of one site (coalesced 256 B GL +
// 64 B AD requests).  The classes of depth <= 3 (86 % of sites at 2x) live in thread-private
// shared-memory counters - no atomics in the streaming loop; deeper
// classes go straight to the table with 64-bit integer atomics.  A thread sees at most
// 2047 sites per launch so the 32-bit register limb sums cannot overflow.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ztally_kernel(const float2* __restrict__ G, const uchar2* __restrict__ AD, int ldg, long M,
              const unsigned char* __restrict__ sel,       // [ldg] 1 = individual requested
              int wx, long sites_per_block,
              ZTally* __restrict__ table,                  // [ldg][kZClasses]
              unsigned long long* __restrict__ deep)       // [ldg] sites with depth > kZDepthCap
{
    // thread-private histogram rows in shared memory: hot[class * 7 + field][thread] - consecutive threads hit
    // consecutive banks, nobody else touches a thread's column (no atomics, no conflicts).  The first version kept
    // these 70 counters in registers and updated them with 70 predicated adds per site at 2 sites in flight per
    // thread: latency-bound at 10 % of HBM.
    extern __shared__ int zhot[];                            // [kZHotX * 7][256]
    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const int wy_count = 8 / wx;
    const int cgx = warp % wx, wy = warp / wx;
    const int col = (blockIdx.x * wx + cgx) * 32 + lane;
    const bool on = col < ldg && sel[col];
    const long s_begin = (long)blockIdx.y * sites_per_block;
    const long s_end = min(M, s_begin + sites_per_block);
    if (!on) return;                                          // no block-wide barrier below

#pragma unroll
    for (int c = 0; c < kZHotX * 7; ++c) zhot[c * 256 + t] = 0;
    int ndeep = 0;
    ZTally* mine = table + (size_t)col * kZClasses;

    constexpr int PF = 8;                                     // sites in flight per thread
    for (long sb = s_begin + wy; sb < s_end; sb += (long)wy_count * PF) {
        float2 gq[PF];
        uchar2 aq[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            long s = sb + (long)wy_count * u;
            gq[u] = make_float2(0.f, 0.f); aq[u] = make_uchar2(255, 255);       // depth 510: counted nowhere
            if (s < s_end) { gq[u] = ld_stream2(&G[s * (long)ldg + col]); aq[u] = AD[s * (long)ldg + col]; }
        }
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            if (sb + (long)wy_count * u >= s_end) break;
            const float2 g = gq[u];
            int ref = aq[u].x, alt = aq[u].y, d = ref + alt;
            int vh0, vl0, vh1, vl1, vh2, vl2;
            fix_limbs(g.x, vh0, vl0);
            fix_limbs(g.y, vh1, vl1);
            fix_limbs(third_gl_np(g.x, g.y), vh2, vl2);
            if (d <= kZHotDepthX) {
                int* h = zhot + zclass_id(ref, alt) * 7 * 256 + t;
                h[0] += 1;
                h[256] += vh0; h[2 * 256] += vl0; h[3 * 256] += vh1; h[4 * 256] += vl1; h[5 * 256] += vh2; h[6 * 256] += vl2;
            } else if (d <= kZDepthCap) {
                tally_flush(mine + zclass_id(ref, alt), 1, vh0, vl0, vh1, vl1, vh2, vl2);
            } else {
                ++ndeep;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < kZHotX; ++c) {
        const int* h = zhot + c * 7 * 256 + t;
        tally_flush(mine + c, h[0], h[256], h[2 * 256], h[3 * 256], h[4 * 256], h[5 * 256], h[6 * 256]);
    }
    if (ndeep) atomicAdd(&deep[col], (unsigned long long)ndeep);
}

// ---------------------------------------------------------------------------------------
// ztally_seq: the reference-faithful tally.  The reference's class mean is numpy's float32
// mean over the class's rows (zscore.py:22), i.e. a SEQUENTIAL float32 sum in site order
// whose rounding error grows with the class size (already 4e-6 relative at 400 sites) and
// feeds both the 0.01 keep test and the expected log-likelihood.  Reproducing its bits
// needs the same order: one thread per individual walks all sites in order (a warp = 32
// individuals, so loads stay coalesced; 8 sites are prefetched ahead).  Hot classes are
// float accumulators in registers updated by predicated adds; the rest are read-modify-
// written in the individual's own table row (no other thread touches it).
// Used on a single GPU up to 2^18 sites (it is serial in the site index: ~0.8 us per site);
// larger and site-sharded runs use the order-independent ztally_kernel above (at that scale the
// reference's own float32 mean has lost its precision anyway).  WGS_Z_EXACT_MEANS=0/1 overrides.
// ---------------------------------------------------------------------------------------
struct ZTallyF { float s0, s1, s2; int cnt; };

__global__ void __launch_bounds__(32)
ztally_seq_kernel(const float2* __restrict__ G, const uchar2* __restrict__ AD, int ldg, long M,
                  const unsigned char* __restrict__ sel,
                  ZTallyF* __restrict__ table,             // [ldg][kZClasses], zeroed
                  unsigned long long* __restrict__ deep)
{
    const int col = blockIdx.x * 32 + threadIdx.x;
    if (col >= ldg || !sel[col]) return;
    int cnt[kZHot];
    float a0[kZHot], a1[kZHot], a2[kZHot];
#pragma unroll
    for (int c = 0; c < kZHot; ++c) { cnt[c] = 0; a0[c] = a1[c] = a2[c] = 0.f; }
    int ndeep = 0;
    ZTallyF* mine = table + (size_t)col * kZClasses;
    for (long sb = 0; sb < M; sb += 8) {
        float2 g[8];
        uchar2 ad[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            g[u] = make_float2(0.f, 0.f); ad[u] = make_uchar2(255, 255);
            if (sb + u < M) { g[u] = ld_stream2(&G[(sb + u) * (long)ldg + col]); ad[u] = AD[(sb + u) * (long)ldg + col]; }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (sb + u >= M) break;
            int ref = ad[u].x, alt = ad[u].y, d = ref + alt;
            float g0 = g[u].x, g1 = g[u].y, g2 = third_gl_np(g0, g1);
            if (d <= kZHotDepth) {
                int code = zclass_id(ref, alt);
#pragma unroll
                for (int c = 0; c < kZHot; ++c) {
                    if (code == c) {
                        cnt[c] += 1;
                        a0[c] = __fadd_rn(a0[c], g0); a1[c] = __fadd_rn(a1[c], g1); a2[c] = __fadd_rn(a2[c], g2);
                    }
                }
            } else if (d <= kZDepthCap) {
                ZTallyF* t = mine + zclass_id(ref, alt);
                ZTallyF v = *t;
                v.s0 = __fadd_rn(v.s0, g0); v.s1 = __fadd_rn(v.s1, g1); v.s2 = __fadd_rn(v.s2, g2); v.cnt += 1;
                *t = v;
            } else {
                ++ndeep;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < kZHot; ++c) { ZTallyF v; v.s0 = a0[c]; v.s1 = a1[c]; v.s2 = a2[c]; v.cnt = cnt[c]; mine[c] = v; }
    if (ndeep) atomicAdd(&deep[col], (unsigned long long)ndeep);
}

// ---------------------------------------------------------------------------------------
// ztally_seq2: the same sequential float32 sums as ztally_seq - the reference's class means
// (zscore.py:22) bit for bit - restructured so that it is the tally of EVERY run: any number
// of sites, and any number of site-sharded ranks.
//
//  * The table is read as the CARRY-IN state and left as the carry-out: under site sharding
//    the ranks run this kernel one after the other in site order and hand the table on
//    (wgs_zscore: ncclSend / ncclRecv over NVLink, 16 bytes per (individual, class)), so the
//    order of every addition is the reference's over the whole file.
//  * One warp = 32 individuals, lane = individual (coalesced rows).  The (GL, AD) rows of 32
//    sites per stage are streamed into an 8-stage shared-memory ring by the warp's own
//    16-byte / 8-byte LDGSTS copies: ~80 KB in flight per warp, because with N / 32 warps on
//    the whole GPU it is bytes in flight per warp, not resident warps, that cover HBM latency
//    (8 register-prefetched sites per thread left ztally_seq latency-bound at ~1 us per 8 sites).
//  * The accumulators of all classes up to depth 9 (55 classes: > 99.9 % of the sites at 2x)
//    are thread-private float4 cells {s0, s1, s2, count} in shared memory, cell[class][lane]:
//    one conflict-free LDS.128 / STS.128 per site whatever the class mix inside the warp
//    (15 register accumulators with predicated adds cost 75 issue slots per site).  Two sites
//    are in flight per thread: both cells are loaded first and the second site takes the
//    first one's result when the classes coincide - the additions of a class stay in site
//    order, the load latency is paid once per pair.  Deeper classes are read-modify-written
//    in the individual's own table row (no other thread touches it).
// ---------------------------------------------------------------------------------------
constexpr int kZSeqHotDepth = 9;
constexpr int kZSeqHot = (kZSeqHotDepth + 1) * (kZSeqHotDepth + 2) / 2;  // 55
constexpr int kZSeqSB = 32;                                          // sites per ring stage
constexpr int kZSeqStages = 8;
constexpr size_t kZSeqSmem = (size_t)kZSeqHot * 32 * sizeof(float4) + (size_t)kZSeqStages * kZSeqSB * 32 * (sizeof(float2) + sizeof(uchar2));

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(gmem));
}

__global__ void __launch_bounds__(32)
ztally_seq2_kernel(const float2* __restrict__ G, const uchar2* __restrict__ AD, int ldg, long M,
                   const unsigned char* __restrict__ sel,
                   ZTallyF* __restrict__ table,            // [ldg][kZClasses]: carry-in, updated in place
                   unsigned long long* __restrict__ deep)
{
    extern __shared__ __align__(16) unsigned char zs_raw[];
    float4* cell = reinterpret_cast<float4*>(zs_raw);                          // [kZSeqHot][32]
    float2* Gs = reinterpret_cast<float2*>(cell + kZSeqHot * 32);              // [stages][SB][32]
    uchar2* As = reinterpret_cast<uchar2*>(Gs + kZSeqStages * kZSeqSB * 32);   // [stages][SB][32]
    const int lane = threadIdx.x;
    const int col0 = blockIdx.x * 32, col = col0 + lane;
    const bool on = col < ldg && sel[col];
    if (__ballot_sync(0xffffffffu, on) == 0u) return;
    const int ncols = min(32, ldg - col0);                  // a multiple of 4: slabs are padded to 4 individuals
    ZTallyF* mine = table + (size_t)(on ? col : col0) * kZClasses;
    for (int c = 0; c < kZSeqHot; ++c) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on) { const ZTallyF t = mine[c]; v = make_float4(t.s0, t.s1, t.s2, __int_as_float(t.cnt)); }
        cell[c * 32 + lane] = v;
    }
    float4* const mycell = cell + lane;
    const int gc = ncols >> 1, ac = ncols >> 2;             // 16-byte GL chunks / 8-byte AD chunks per site row
    auto stage_load = [&](long s0, int st) {
        const int rows = (int)min((long)kZSeqSB, M - s0);
        if (rows > 0) {
            float2* gd = Gs + (size_t)st * kZSeqSB * 32;
            uchar2* ad = As + (size_t)st * kZSeqSB * 32;
            if (ncols == 32) {
                for (int e = lane; e < rows * 16; e += 32) { const int r = e >> 4, c = e & 15; cp_async16(gd + r * 32 + 2 * c, G + (s0 + r) * (long)ldg + col0 + 2 * c); }
                for (int e = lane; e < rows * 8; e += 32) { const int r = e >> 3, c = e & 7; cp_async8(ad + r * 32 + 4 * c, AD + (s0 + r) * (long)ldg + col0 + 4 * c); }
            } else {
                for (int e = lane; e < rows * gc; e += 32) { const int r = e / gc, c = e - r * gc; cp_async16(gd + r * 32 + 2 * c, G + (s0 + r) * (long)ldg + col0 + 2 * c); }
                for (int e = lane; e < rows * ac; e += 32) { const int r = e / ac, c = e - r * ac; cp_async8(ad + r * 32 + 4 * c, AD + (s0 + r) * (long)ldg + col0 + 4 * c); }
            }
        }
        cp_async_commit();                                  // always: the wait below counts groups
    };
    const long nst = (M + kZSeqSB - 1) / kZSeqSB;
    for (int p = 0; p < kZSeqStages - 1; ++p) stage_load((long)p * kZSeqSB, p);
    int ndeep = 0;
    ZTallyF* cold = table + (size_t)col * kZClasses;
    for (long j = 0; j < nst; ++j) {
        const int st = (int)(j % kZSeqStages);
        cp_async_wait<kZSeqStages - 2>();                   // stage j has landed
        __syncwarp();
        stage_load((j + kZSeqStages - 1) * kZSeqSB, (int)((j + kZSeqStages - 1) % kZSeqStages));   // refills the slot consumed at j-1
        const long s0 = j * kZSeqSB;
        const int rows = (int)min((long)kZSeqSB, M - s0);
        if (on) {
            const float2* gs = Gs + (size_t)st * kZSeqSB * 32 + lane;
            const uchar2* as = As + (size_t)st * kZSeqSB * 32 + lane;
#pragma unroll 2
            for (int u = 0; u < rows; u += 2) {
                const bool two = u + 1 < rows;
                const uchar2 a1 = as[u * 32], a2 = two ? as[(u + 1) * 32] : make_uchar2(255, 255);
                const float2 g1 = gs[u * 32], g2 = two ? gs[(u + 1) * 32] : make_float2(0.f, 0.f);
                const int d1 = a1.x + a1.y, d2 = a2.x + a2.y;
                const int c1 = d1 * (d1 + 1) / 2 + a1.y, c2 = d2 * (d2 + 1) / 2 + a2.y;
                const bool h1 = d1 <= kZSeqHotDepth, h2 = two && d2 <= kZSeqHotDepth;
                float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1;
                if (h1) v1 = mycell[c1 * 32];
                if (h2) v2 = mycell[c2 * 32];
                const float t1 = third_gl_np(g1.x, g1.y), t2 = third_gl_np(g2.x, g2.y);
                if (h1) {
                    v1.x = __fadd_rn(v1.x, g1.x); v1.y = __fadd_rn(v1.y, g1.y); v1.z = __fadd_rn(v1.z, t1);
                    v1.w = __int_as_float(__float_as_int(v1.w) + 1);
                }
                if (h1 && h2 && c1 == c2) v2 = v1;           // same class twice in a row: the second addition sees the first
                if (h2) {
                    v2.x = __fadd_rn(v2.x, g2.x); v2.y = __fadd_rn(v2.y, g2.y); v2.z = __fadd_rn(v2.z, t2);
                    v2.w = __int_as_float(__float_as_int(v2.w) + 1);
                }
                if (h1) mycell[c1 * 32] = v1;                 // program order: when the classes coincide the second store wins
                if (h2) mycell[c2 * 32] = v2;
                if (!h1) {                                  // rare: deeper than the shared-memory cells
                    if (d1 <= kZDepthCap) {
                        ZTallyF v = cold[c1];
                        v.s0 = __fadd_rn(v.s0, g1.x); v.s1 = __fadd_rn(v.s1, g1.y); v.s2 = __fadd_rn(v.s2, t1); v.cnt += 1;
                        cold[c1] = v;
                    } else ++ndeep;
                }
                if (two && !h2) {
                    if (d2 <= kZDepthCap) {
                        ZTallyF v = cold[c2];
                        v.s0 = __fadd_rn(v.s0, g2.x); v.s1 = __fadd_rn(v.s1, g2.y); v.s2 = __fadd_rn(v.s2, t2); v.cnt += 1;
                        cold[c2] = v;
                    } else ++ndeep;
                }
            }
        }
        __syncwarp();                                       // everyone is done with stage j before a later load refills it
    }
    cp_async_wait<0>();
    if (on) {
        for (int c = 0; c < kZSeqHot; ++c) {
            const float4 v = cell[c * 32 + lane];
            ZTallyF t; t.s0 = v.x; t.s1 = v.y; t.s2 = v.z; t.cnt = __float_as_int(v.w);
            mine[c] = t;
        }
        if (ndeep) atomicAdd(&deep[col], (unsigned long long)ndeep);
    }
}

// Deepest read depth with a non-empty class in the tally table: the host then moves and scans only the
// classes up to that depth (66 instead of 861 per individual at 10 reads) - the full-table copies and loops
// were 40 % of a z-score call at 2,000 individuals.
template <class T>
__global__ void zmaxdepth_kernel(const T* __restrict__ table, long n, int* __restrict__ out)
{
    int best = 0;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        if (table[e].cnt > 0) {
            const int id = (int)(e % kZClasses);
            int d = 0;
            while ((d + 1) * (d + 2) / 2 <= id) ++d;
            best = max(best, d);
        }
    }
    if (best > 0) atomicMax(out, best);
}

// Per-(individual, class) decision tables built on the host from the tallies:
//   kmax[id][col]  : -1 class not kept, else argmax of the class mean (zscore.py:53)
//   kmean[id][col] : the class mean at that argmax (float32)
//   zlike/zfac[col][id] : class mean GL triple (AD_like) and binomial read probabilities
//                         (AD_factorial), zscore.py:63-79
// ---------------------------------------------------------------------------------------
// zkeep: keep[s][col] = class kept && !(|mean[max] - GL[max]| > 0.01)  (zscore.py:43-61)
// and the per-individual kept-site count (loci_kept - bit-exact).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
zkeep_kernel(const float2* __restrict__ G, const uchar2* __restrict__ AD, int ldg, long M,
             const unsigned char* __restrict__ sel,
             const signed char* __restrict__ kmax, const float* __restrict__ kmean, int ncls,   // [ncls][ldg]
             int wx, long sites_per_block,
             unsigned char* __restrict__ keep,             // [M][ldg]
             unsigned long long* __restrict__ kept)        // [ldg]
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy_count = 8 / wx;
    const int cgx = warp % wx, wy = warp / wx;
    const int col = (blockIdx.x * wx + cgx) * 32 + lane;
    const bool on = col < ldg && sel[col];
    const long s_begin = (long)blockIdx.y * sites_per_block;
    const long s_end = min(M, s_begin + sites_per_block);
    if (col >= ldg) return;
    const signed char* km = kmax + col;                     // class-major tables [class][ldg]: the lanes of a warp (consecutive
    const float* kv = kmean + col;                          // individuals, mostly the same few classes) share cache lines
    int nk = 0;
    for (long s = s_begin + wy; s < s_end; s += wy_count) {
        unsigned char k = 0;
        if (on) {
            uchar2 ad = AD[s * (long)ldg + col];
            int ref = ad.x, alt = ad.y;
            if (ref + alt <= kZDepthCap) {
                int id = zclass_id(ref, alt);
                int mx = id < ncls ? km[(long)id * ldg] : -1;
                if (mx >= 0) {
                    float2 g = ld_stream2(&G[s * (long)ldg + col]);
                    float gl = mx == 0 ? g.x : (mx == 1 ? g.y : third_gl_np(g.x, g.y));
                    float diff = fabsf(__fsub_rn(kv[(long)id * ldg], gl));
                    k = !(diff > 0.01f);
                }
            }
        }
        keep[s * (long)ldg + col] = k;
        nk += k;
    }
    if (nk) atomicAdd(&kept[col], (unsigned long long)nk);
}

// ---------------------------------------------------------------------------------------
// zmoments: for every kept (site, individual): observed log-likelihood, its expectation
// over all read splits of the site's depth and the variance (zscore_cy.pyx:10-56).
// af = afbase[s * af_ld + afcol[col]]: a column of the caller's AF matrix (assignment
// mode) or the individual's own leave-one-out state column (reference mode).
// Per-thread float64 sums -> partials[split][col][3] -> fixed-order reduction.
// ---------------------------------------------------------------------------------------
// Table layout: the class means are CLASS-major, zlike_t[id][ldg] - the lanes of a warp are consecutive
// individuals and mostly sit in the same few classes, so one load touches a handful of 512-byte rows instead of
// 32 scattered lines (the individual-major table made this kernel L1-transaction bound: 4 scattered 16-byte
// loads per class and pass = 41 ms at 500 k x 2,000).  The binomial read probabilities depend on the class
// only: one [ncls] table staged in shared memory.  The per-class log terms of a site are kept in registers
// for the variance pass (depths up to kZCacheDepth; deeper sites recompute them).
constexpr int kZCacheDepth = 7;
__global__ void __launch_bounds__(256)
zmoments_kernel(const float2* __restrict__ G, const uchar2* __restrict__ AD, int ldg, long M,
                const unsigned char* __restrict__ keep,
                const float* __restrict__ afbase, int af_ld, const int* __restrict__ afcol,
                const float4* __restrict__ zlike_t,        // [ncls][ldg] class means (GL triple), class-major
                const float4* __restrict__ zfac_c, int ncls,   // [ncls] binomial read probabilities per class
                int wx, long sites_per_block,
                double* __restrict__ partials)             // [gridDim.y][ldg][3]
{
    extern __shared__ __align__(16) unsigned char zsmem[];
    float4* fac_s = reinterpret_cast<float4*>(zsmem);     // [ncls]
    __shared__ double red[8][32][3];
    for (int e = threadIdx.x; e < ncls; e += blockDim.x) fac_s[e] = zfac_c[e];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wy_count = 8 / wx;
    const int cgx = warp % wx, wy = warp / wx;
    const int col = (blockIdx.x * wx + cgx) * 32 + lane;
    const bool col_ok = col < ldg;
    const long s_begin = (long)blockIdx.y * sites_per_block;
    const long s_end = min(M, s_begin + sites_per_block);
    double w_obs = 0.0, w_mu = 0.0, w_var = 0.0;
    if (col_ok) {
        const int ac = afcol[col];
        const float4* lk = zlike_t + col;
        for (long s = s_begin + wy; s < s_end; s += wy_count) {
            if (!keep[s * (long)ldg + col]) continue;
            float2 g = ld_stream2(&G[s * (long)ldg + col]);
            uchar2 ad = AD[s * (long)ldg + col];
            float a = __ldg(&afbase[s * (long)af_ld + ac]);
            float om = 1.0f - a;
            float P0 = om * om, P1 = 2.0f * om * a, P2 = a * a;                 // zscore_cy.pyx:19-21
            float lobs = logf(fmaf(g.x, P0, fmaf(g.y, P1, third_gl(g.x, g.y) * P2)));   // :22-25
            int Dl = ad.x + ad.y;
            int base = Dl * (Dl + 1) / 2;
            float wl = 0.f, vr = 0.f;
            if (Dl <= kZCacheDepth) {
                float ev[kZCacheDepth + 1];
#pragma unroll
                for (int x = 0; x <= kZCacheDepth; ++x) {   // class (ref = x, alt = Dl - x), the order of :28-30
                    ev[x] = 0.f;
                    if (x <= Dl) {
                        int id = base + (Dl - x);
                        float4 l = __ldg(&lk[(long)id * ldg]), f = fac_s[id];
                        float e = logf(fmaf(l.x, P0, fmaf(l.y, P1, l.z * P2)));       // :31
                        ev[x] = e;
                        wl = __fadd_rn(wl, e * P0 * f.x);                              // :32-34, float32 after every add
                        wl = __fadd_rn(wl, e * P1 * f.y);
                        wl = __fadd_rn(wl, e * P2 * f.z);
                    }
                }
#pragma unroll
                for (int x = 0; x <= kZCacheDepth; ++x) {
                    if (x <= Dl) {
                        float4 f = fac_s[base + (Dl - x)];
                        float dd = wl - ev[x];
                        dd = dd * dd;
                        vr = __fadd_rn(vr, dd * P0 * f.x);                             // :54-56
                        vr = __fadd_rn(vr, dd * P1 * f.y);
                        vr = __fadd_rn(vr, dd * P2 * f.z);
                    }
                }
            } else {
                for (int x = 0; x <= Dl; ++x) {
                    int id = base + (Dl - x);
                    float4 l = __ldg(&lk[(long)id * ldg]), f = fac_s[id];
                    float e = logf(fmaf(l.x, P0, fmaf(l.y, P1, l.z * P2)));
                    wl = __fadd_rn(wl, e * P0 * f.x);
                    wl = __fadd_rn(wl, e * P1 * f.y);
                    wl = __fadd_rn(wl, e * P2 * f.z);
                }
                for (int x = 0; x <= Dl; ++x) {
                    int id = base + (Dl - x);
                    float4 l = __ldg(&lk[(long)id * ldg]), f = fac_s[id];
                    float e = logf(fmaf(l.x, P0, fmaf(l.y, P1, l.z * P2)));
                    float dd = wl - e;
                    dd = dd * dd;
                    vr = __fadd_rn(vr, dd * P0 * f.x);
                    vr = __fadd_rn(vr, dd * P1 * f.y);
                    vr = __fadd_rn(vr, dd * P2 * f.z);
                }
            }
            w_obs += (double)lobs; w_mu += (double)wl; w_var += (double)vr;
        }
    }
    red[warp][lane][0] = w_obs; red[warp][lane][1] = w_mu; red[warp][lane][2] = w_var;
    __syncthreads();
    if (wy == 0 && col_ok) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            double v = 0.0;
            for (int w = 0; w < wy_count; ++w) v += red[w * wx + cgx][lane][q];
            partials[((long)blockIdx.y * ldg + col) * 3 + q] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------
// zmoments_list: the per-site arrays of zscore_cy.expected_W_l / variance_W_l for ONE
// individual over an explicit list of kept sites, with the caller's own class tables
// (zscore_cy.pyx:10-56, including the AD_index[Aa, Ar] lookup convention of :30).  One thread
// per kept site; outputs are the float32 per-site vectors the Cython functions fill.
// ---------------------------------------------------------------------------------------
__global__ void zmoments_list_kernel(const float2* __restrict__ G, const uchar2* __restrict__ AD, int ldg, int col,
                                     const int* __restrict__ L_keep, long mk, const float* __restrict__ Avec,
                                     const float* __restrict__ AD_factorial, const float* __restrict__ AD_like,
                                     const int* __restrict__ AD_index, int idx_rows, int idx_cols,
                                     float* __restrict__ W_obs, float* __restrict__ W_l, float* __restrict__ W_var)
{
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < mk; e += (long)gridDim.x * blockDim.x) {
        const long s = L_keep[e];
        const float2 g = G[s * (long)ldg + col];
        const uchar2 ad = AD[s * (long)ldg + col];
        const float a = Avec[e];
        const float om = 1.0f - a;
        const float P0 = om * om, P1 = 2.0f * om * a, P2 = a * a;
        W_obs[e] = logf(fmaf(g.x, P0, fmaf(g.y, P1, third_gl(g.x, g.y) * P2)));
        const int Dl = ad.x + ad.y;
        float wl = 0.f;
        for (int Aa = 0; Aa <= Dl; ++Aa) {
            const int Ar = Dl - Aa;
            int c = 0;
            if (Aa < idx_rows && Ar < idx_cols) c = AD_index[Aa * idx_cols + Ar];
            const float* l = AD_like + 3 * c;
            const float* f = AD_factorial + 3 * c;
            const float ee = logf(fmaf(l[0], P0, fmaf(l[1], P1, l[2] * P2)));
            wl = __fadd_rn(wl, ee * P0 * f[0]);
            wl = __fadd_rn(wl, ee * P1 * f[1]);
            wl = __fadd_rn(wl, ee * P2 * f[2]);
        }
        float vr = 0.f;
        for (int Aa = 0; Aa <= Dl; ++Aa) {
            const int Ar = Dl - Aa;
            int c = 0;
            if (Aa < idx_rows && Ar < idx_cols) c = AD_index[Aa * idx_cols + Ar];
            const float* l = AD_like + 3 * c;
            const float* f = AD_factorial + 3 * c;
            const float ee = logf(fmaf(l[0], P0, fmaf(l[1], P1, l[2] * P2)));
            float dd = wl - ee;
            dd = dd * dd;
            vr = __fadd_rn(vr, dd * P0 * f[0]);
            vr = __fadd_rn(vr, dd * P1 * f[1]);
            vr = __fadd_rn(vr, dd * P2 * f[2]);
        }
        W_l[e] = wl;
        W_var[e] = vr;
    }
}

}  // namespace wgs
