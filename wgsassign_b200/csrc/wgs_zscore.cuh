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
// ztally_ord: the reference-faithful tally.  The reference's class mean is numpy's float32 mean over the class's
// rows (zscore.py:22), i.e. a SEQUENTIAL float32 sum in site order whose rounding error grows with the class size
// (already 4e-6 relative at 400 sites) and feeds both the 0.01 keep test and the expected log-likelihood.
// Reproducing its bits needs its order - but only WITHIN a class: sums of different classes, components and
// individuals are independent chains.
//
//  * One WARP = one individual, LANE = SITE: a batch is 32 consecutive sites.  Lanes whose sites fall in the same
//    (ref, alt) class find each other with MATCH.ANY; the class's first lane (the leader) takes the class's running
//    sums {s0, s1, s2, count} from the warp's shared-memory cells and adds its members' GL triples in lane (= site)
//    order from the warp's strip, in branch-free rounds (trip count = the largest class of the batch, ~10 of 32 at
//    2x).  The float32 additions of one class are the only dependent chain; different classes run in different lanes.
//    Measured designs, 1M sites x 2,000 individuals: one thread per individual walking the sites 49 ms; lane = class
//    with register accumulators 17..20 ms (every lane walks all 32 codes); MATCH + divergent leader loops 20 ms;
//    MATCH + branch-free rounds 15.3 ms (this one) - the pass is instruction-bound (9.3 warp instructions per site
//    and individual before, ~6 now), not HBM-bound: profiles/ncu_summary_r2_zscore_v7.txt.
//  * The kZOrdCells = 66 classes of depth <= 10 have cells in shared memory (> 99.99 % of the sites at 2x).
//  * Deeper classes (up to the depth cap) are read-modify-written in the individual's own table row by their leader.
//  * A block = kZOrdWarps adjacent individuals: their rows of a site tile arrive with coalesced 16-byte / 8-byte LDGSTS
//    copies into a ring of kZOrdStages tiles.
//  * The table is read as the CARRY-IN state and left as the carry-out: under site sharding the ranks run this
//    kernel one after the other in site order and hand the table on (wgs_zscore: ncclSend / ncclRecv over NVLink,
//    16 bytes per (individual, class)), so the order of every addition is the reference's over the whole file.
//    [col_lo, col_hi) restricts a launch to a column group: the groups pipeline through the ranks.
// ---------------------------------------------------------------------------------------
struct ZTallyF { float s0, s1, s2; int cnt; };
constexpr int kZOrdWarps = 16;                                           // individuals per block: 128-byte GL rows (one line), 32 bytes of depths
constexpr int kZOrdTile = 128;                                           // sites per staged tile (4 batches of 32)
constexpr int kZOrdStages = 4;                                           // ring depth: 3 tiles (60 KB) in flight per block - with two, the pass ran at
                                                                         //   0.55 TB/s whatever the arithmetic: bytes in flight, not instructions
constexpr int kZOrdGS = 18;                                              // tile row strides (float2 / uchar2 units): 16- / 8-byte aligned rows whose
constexpr int kZOrdAS = 20;                                              //   column reads (lane = row) spread over the banks
#ifndef WGS_ZORD_UNROLL
#define WGS_ZORD_UNROLL 2
#endif
constexpr int kZOrdUnroll = WGS_ZORD_UNROLL;                             // additions per trip of a leader's loop
constexpr int kZOrdCells = 66;                                           // classes with shared-memory cells: every depth up to 10
constexpr size_t kZOrdSmem = (size_t)kZOrdWarps * (32 + kZOrdCells) * sizeof(float4) +
                             (size_t)kZOrdStages * kZOrdTile * (kZOrdGS * sizeof(float2) + kZOrdAS * sizeof(uchar2));

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(gmem));
}

// One batch = 32 consecutive sites of one individual, lane = site.  Lanes of the same (ref, alt) class find each other
// with MATCH.ANY; the class's first member (the LEADER) reads the class's running sums from the warp's cells and adds
// the members' GL triples in lane (= site) order, reading them from the warp's strip - the only dependent chain is the
// float32 additions themselves.  The trip count is the size of the largest class in the batch (warp-uniform, ~10 of 32
// at 2x): ~4 warp instructions per site.
__device__ __forceinline__ void ztally_ord_batch(float2 g, uchar2 a, bool valid, int lane, float4* __restrict__ strip,
                                                 float4* __restrict__ cell, ZTallyF* __restrict__ mine, int& ndeep)
{
    const int d = a.x + a.y;
    const int code = d * (d + 1) / 2 + a.y;
    const bool deepf = valid && d > kZDepthCap;
    const bool act = valid && !deepf;
    strip[lane] = make_float4(g.x, g.y, third_gl_np(g.x, g.y), 0.f);
    ndeep += __popc(__ballot_sync(0xffffffffu, deepf));      // same value in every lane; lane 0 reports it
    const unsigned peers = __match_any_sync(0xffffffffu, act ? code : -1 - lane);   // lanes of the same class (inactive lanes: alone)
    const bool leader = act && (peers & ((1u << lane) - 1u)) == 0u;
    const int maxn = __reduce_max_sync(0xffffffffu, act ? __popc(peers) : 0);
    const bool hot = code < kZOrdCells;
    f32x2 v01 = pack2(0.f, 0.f);
    float v2 = 0.f;
    int cnt = 0;
    unsigned rem = leader ? peers : 0u;                      // members still to add, the leader itself first
    if (leader) {
        if (hot) { const float4 c = cell[code]; v01 = pack2(c.x, c.y); v2 = c.z; cnt = __float_as_int(c.w); }
        else { const ZTallyF t = mine[code]; v01 = pack2(t.s0, t.s1); v2 = t.s2; cnt = t.cnt; }
        cnt += __popc(peers);
    }
    __syncwarp();                                            // the strip is complete
    // branch-free rounds (a divergent `if` per round cost a convergence barrier pair and doubled the instruction count):
    // a lane with nothing left re-reads its own strip entry and adds +0 to nothing
    // kZOrdUnroll rounds per trip: the strip reads of a trip are issued together (their addresses come from `rem` alone),
    // so one shared-memory latency covers the trip instead of one per addition (measured at 1M x 2,000, both launches of
    // a z-score call: 30.7 ms with 1 round per trip, 28.6 with 2, 29.0 with 4, 29.5 with 8)
#pragma unroll 1
    for (int r = 0; r < maxn; r += kZOrdUnroll) {            // warp-uniform trip count
        float4 t[kZOrdUnroll];
        bool more[kZOrdUnroll];
#pragma unroll
        for (int u = 0; u < kZOrdUnroll; ++u) {
            more[u] = rem != 0u;
            t[u] = strip[more[u] ? __ffs(rem) - 1 : lane];
            rem &= rem - 1u;                                 // 0 stays 0
        }
#pragma unroll
        for (int u = 0; u < kZOrdUnroll; ++u) {
            const f32x2 n01 = fadd2(v01, pack2(t[u].x, t[u].y));
            const float n2 = __fadd_rn(v2, t[u].z);
            v01 = more[u] ? n01 : v01;
            v2 = more[u] ? n2 : v2;
        }
    }
    if (leader) {
        const float2 v = unpack2(v01);
        if (hot) cell[code] = make_float4(v.x, v.y, v2, __int_as_float(cnt));
        else { ZTallyF t; t.s0 = v.x; t.s1 = v.y; t.s2 = v2; t.cnt = cnt; mine[code] = t; }
    }
    __syncwarp();                                            // cells and strip are reused by the next batch
}

__global__ void __launch_bounds__(kZOrdWarps * 32)
ztally_ord_kernel(const float2* __restrict__ G, const uchar2* __restrict__ AD, int ldg, long M,
                  const unsigned char* __restrict__ sel, int col_lo, int col_hi,
                  ZTallyF* __restrict__ table,             // [ldg][kZClasses]: carry-in, updated in place
                  unsigned long long* __restrict__ deep)
{
    extern __shared__ __align__(16) unsigned char zs_raw[];
    float4* gbuf = reinterpret_cast<float4*>(zs_raw);                                          // [warps][32] GL triples of a batch
    float4* cells = gbuf + kZOrdWarps * 32;                                                    // [warps][kZOrdCells] running sums
    float2* Gt = reinterpret_cast<float2*>(cells + kZOrdWarps * kZOrdCells);                   // [stages][tile][kZOrdGS]
    uchar2* At = reinterpret_cast<uchar2*>(Gt + (size_t)kZOrdStages * kZOrdTile * kZOrdGS);    // [stages][tile][kZOrdAS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col0 = col_lo + blockIdx.x * kZOrdWarps;
    const int ncols = min(kZOrdWarps, col_hi - col0);        // a multiple of 4: slabs are padded to 4 individuals
    const int col = col0 + warp;
    const bool on = warp < ncols && sel[col];                // warp-uniform
    ZTallyF* mine = table + (size_t)(warp < ncols ? col : col0) * kZClasses;
    float4* cell = cells + warp * kZOrdCells;
    if (on) for (int c = lane; c < kZOrdCells; c += 32) { const ZTallyF t = mine[c]; cell[c] = make_float4(t.s0, t.s1, t.s2, __int_as_float(t.cnt)); }
    // the block's columns of sites [128 t, 128 t + 128): coalesced rows, 16-byte GL chunks and 8-byte depth chunks
    const int gc = ncols >> 1, ac = ncols >> 2;
    auto stage = [&](long t) {
        const int buf = (int)(t % kZOrdStages);
        const long s0 = t * kZOrdTile;
        const int rows = (int)max(0L, min((long)kZOrdTile, M - s0));
        float2* gd = Gt + (size_t)buf * kZOrdTile * kZOrdGS;
        uchar2* ad = At + (size_t)buf * kZOrdTile * kZOrdAS;
        if (ncols == kZOrdWarps) {
            for (int e = tid; e < rows * 8; e += kZOrdWarps * 32) { const int r = e >> 3, c = e & 7; cp_async16(gd + r * kZOrdGS + 2 * c, G + (s0 + r) * (long)ldg + col0 + 2 * c); }
            for (int e = tid; e < rows * 4; e += kZOrdWarps * 32) { const int r = e >> 2, c = e & 3; cp_async8(ad + r * kZOrdAS + 4 * c, AD + (s0 + r) * (long)ldg + col0 + 4 * c); }
        } else {
            for (int e = tid; e < rows * gc; e += kZOrdWarps * 32) { const int r = e / gc, c = e - r * gc; cp_async16(gd + r * kZOrdGS + 2 * c, G + (s0 + r) * (long)ldg + col0 + 2 * c); }
            for (int e = tid; e < rows * ac; e += kZOrdWarps * 32) { const int r = e / ac, c = e - r * ac; cp_async8(ad + r * kZOrdAS + 4 * c, AD + (s0 + r) * (long)ldg + col0 + 4 * c); }
        }
        cp_async_commit();                                  // always: the wait below counts groups
    };
    const long ntiles = (M + kZOrdTile - 1) / kZOrdTile;
    int ndeep = 0;
    float4* strip = gbuf + warp * 32;
    for (int p = 0; p < kZOrdStages - 1; ++p) stage(p);
    for (long t = 0; t < ntiles; ++t) {
        const int buf = (int)(t % kZOrdStages);
        stage(t + kZOrdStages - 1);                         // into the slot tile t-1 used: released by the barrier that ended it
        cp_async_wait<kZOrdStages - 1>();
        __syncthreads();                                    // tile t complete, for every thread's copies
        if (on) {
            const long s0 = t * kZOrdTile;
#pragma unroll 1
            for (int p = 0; p < kZOrdTile / 32; ++p) {
                const long b = s0 + p * 32;
                if (b >= M) break;                           // warp-uniform
                const int row = p * 32 + lane;
                const bool valid = b + lane < M;
                const float2 g = valid ? Gt[((size_t)buf * kZOrdTile + row) * kZOrdGS + warp] : make_float2(0.f, 0.f);
                const uchar2 a = valid ? At[((size_t)buf * kZOrdTile + row) * kZOrdAS + warp] : make_uchar2(0, 0);
                ztally_ord_batch(g, a, valid, lane, strip, cell, mine, ndeep);
            }
        }
        __syncthreads();                                    // everyone is done with tile t before its buffer is refilled
    }
    cp_async_wait<0>();
    if (on) {
        __syncwarp();
        for (int c = lane; c < kZOrdCells; c += 32) {
            const float4 v = cell[c];
            ZTallyF tt; tt.s0 = v.x; tt.s1 = v.y; tt.s2 = v.z; tt.cnt = __float_as_int(v.w);
            mine[c] = tt;
        }
        if (lane == 0 && ndeep) atomicAdd(&deep[col], (unsigned long long)ndeep);
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
