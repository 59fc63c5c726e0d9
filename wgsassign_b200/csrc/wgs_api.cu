// wgsassign_b200 - C ABI (include/wgsassign_b200.h) over the sm_100a kernels.
// Host-side orchestration only: residency, launch configuration, the EM stop-rule loop and
// fixed-order reductions.  There is no CPU implementation of any operator in this file.
#include "../../include/wgsassign_b200.h"
#include "wgs_kernels.cuh"
#include "wgs_zscore.cuh"

#include <algorithm>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <limits>
#include <functional>
#include <map>
#include <string>
#include <vector>

using namespace wgs;

namespace {
std::string g_create_error;
constexpr int kNumSM_fallback = 148;

struct TimedLaunch { int name_id; cudaEvent_t a, b; cudaStream_t st; };
struct FamilyStat { double ms = 0, bytes = 0, units = 0; long launches = 0; };

}  // namespace

struct wgs_ctx {
    int device = 0;
    int num_sm = kNumSM_fallback;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    std::string err;

    // structure
    int N = 0, K = 0, ldg = 0;
    bool pops_set = false;
    std::vector<int> pop_of_ind, col_of_ind, ind_of_col, pop_of_col;
    std::vector<PopDesc> pops;
    int *d_ind_of_col = nullptr, *d_col_of_ind = nullptr, *d_pop_of_col = nullptr;
    PopDesc* d_pops = nullptr;

    // resident data
    float2* G[2] = {nullptr, nullptr};
    long Mg[2] = {0, 0};
    uchar2* AD = nullptr;
    long M_ad = 0;

    // asynchronous upload (wgs_upload_gl_async): one event per population slab on stream2
    bool upload_pending = false;
    long upload_capacity = 0;                    // rows the GL buffer was allocated for (wgs_upload_gl_begin)
    std::vector<cudaEvent_t> ev_pop;
    std::vector<int> upload_order;               // populations in the order their slabs were queued (smallest first)

    // EM decisions come back through two mapped pinned slots ([0] = problems still active, [1..] = flags)
    int* em_pin[2] = {nullptr, nullptr};
    int* em_pin_dev[2] = {nullptr, nullptr};
    size_t em_pin_cap = 0;
    cudaEvent_t em_ev[2] = {nullptr, nullptr};

    // NCCL communicator for the EM stop rule under site sharding (wgs_nccl_init): the squared changes are
    // all-gathered on the compute stream, so the sharded EM keeps the single-GPU look-ahead schedule
    void* nccl_comm = nullptr;
    int nccl_rank = 0, nccl_world = 1;

    // sharding
    int rank = 0, world = 1;                     // wgs_set_rank (or wgs_nccl_init): position in the site order of the ranks
    long M_total = -1, site_offset = 0;
    wgs_allreduce_fn fn = nullptr;
    void* user = nullptr;

    // caching allocator
    std::multimap<size_t, void*> pool_free;
    std::map<void*, size_t> pool_size;

    // allele frequencies kept on the device between wgs_ref_af and wgs_loo_partial (NULL host pointers)
    float* d_af = nullptr;
    long af_rows = 0; int af_cols = 0;

    // z-score class tables of the last call
    std::vector<std::vector<int>> zclasses;
    std::vector<std::vector<float>> ztable;      // per individual: rows of (ref, alt, n_loci, mean0, mean1, mean2, kept) of every observed class
    long z_deep_sites = 0;
    long ad_saturated = 0;                       // (site, individual) pairs of the resident depths with a count above 254

    // run-time options (wgs_set_option): experiment / fallback switches, read at call entry.  Environment variables
    // are NOT consulted on any call path; with WGS_DEBUG set, wgs_create imports WGS_<NAME>=<int> once.
    std::map<std::string, int> opts;

    // device_free_bytes(): free device memory at the pool's last change
    long pool_epoch = 0, mem_epoch_seen = -1;
    size_t mem_free_seen = 0;

    // instrumentation
    long launches = 0;
    bool timing = false;
    std::vector<std::string> tnames;
    std::vector<TimedLaunch> timed;
    std::map<std::string, FamilyStat> tdone;

    long M() const { return Mg[0]; }
    long Mtot() const { return M_total >= 0 ? M_total : Mg[0]; }
};

namespace {

int opt(const wgs_ctx* c, const char* name, int def = 0)
{
    auto it = c->opts.find(name);
    return it == c->opts.end() ? def : it->second;
}

// every option the library reads, with its meaning (wgs_set_option rejects anything else)
const char* const kOptionNames[] = {
    "trace",              // 1: wall-clock of the host-side phases of every call on stderr
    "pl2_wx",             // pop_like2: warps per block (1..4), 0 = least padding
    "poplike_v1",         // 1: direct-form likelihood kernel even where the ratio form applies
    "loolike_v1",         // 1: gather-through-L1 leave-one-out likelihood kernel
    "loolike_v2",         // 1: staged state-row kernel (loo_like2) even where loo_like3 applies
    "loolike_smallblock", // 1: loo_like2 in two small blocks per SM
    "fisher_v1",          // 1: TMA-tile Fisher kernel
    "em_step",            // 1: population EM with one iteration per launch
    "em_multi1",          // 1: shared-memory-tile multi-iteration population EM
    "em_no_lookahead",    // 1: read every stop decision before queueing the next iteration
    "loo_v4", "loo_nofirst", "loo_fullfill", "loo_block", "loo_stages", "loo_passes", "loo_dbg", "prepack_v1",
    "loo_first_bpsm",     // blocks per SM of the first leave-one-out iteration (default 2)
    "loo_big_margin_pm",  // per mille of thread utilisation a 512-thread block must gain over two 256-thread ones to be chosen (default 0: chosen on a tie)
    "loo_by_pop",         // 1: leave-one-out EM population by population (one packed-row buffer at a time); -1: never
    "upload_sync",        // 1: wgs_upload_gl_async falls back to the chunked synchronous upload
    "ztally_groups",      // column groups of the rank-chained class tally (default 1: a group launch is latency-bound and takes as long as the full one)
    "z_exact_means",      // 1: order-independent fixed-point class means (NOT the reference's float32 means)
    "rmse_exact",         // 0: stop rule on the float64 sum only (no sequential float32 tie-break)
    "rmse_band_ppm",      // tie-break band around the tolerance: 0 = 8x the measured float32 bias (default), > 0 = ppm of the RMSE,
                          //   -1 = every check is resolved sequentially, -2 = the rigorous summation bound
    "nccl_sums",          // 0: final sums through the host callback even when a communicator is attached
};

int fail(wgs_ctx* c, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return 1;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ---- NCCL, bound at run time (the process usually has torch's libnccl.so.2 loaded already) ----------------
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /* ncclUniqueId by value */ struct NcclId128, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
struct NcclId128 { char internal[128]; };
NcclApi g_nccl;
bool nccl_load()
{
    if (g_nccl.lib) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, NcclId128, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclSend");
    g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclRecv");
    g_nccl.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclBroadcast");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy || !g_nccl.Send || !g_nccl.Recv ||
        !g_nccl.Broadcast) return false;
    g_nccl.lib = h;
    return true;
}
constexpr int kNcclFloat64 = 8;                                 // ncclDataType_t::ncclFloat64 (nccl.h)
constexpr int kNcclInt64 = 4, kNcclChar = 0;

int name_id(wgs_ctx* ctx, const char* name)
{
    for (size_t i = 0; i < ctx->tnames.size(); ++i) if (ctx->tnames[i] == name) return (int)i;
    ctx->tnames.push_back(name);
    return (int)ctx->tnames.size() - 1;
}

struct LaunchScope {
    wgs_ctx* ctx; cudaStream_t st; TimedLaunch tl; bool on;
    LaunchScope(wgs_ctx* c, const char* name, cudaStream_t s) : ctx(c), st(s), on(c->timing) {
        ++ctx->launches;
        if (on) {
            tl.name_id = name_id(ctx, name);
            tl.st = st;
            cudaEventCreate(&tl.a); cudaEventCreate(&tl.b);
            cudaEventRecord(tl.a, st);
        }
    }
    ~LaunchScope() { if (on) { cudaEventRecord(tl.b, st); ctx->timed.push_back(tl); } }
};
#define LAUNCH(name, kern, grid, block, smem, st, ...)                 \
    do {                                                               \
        LaunchScope ls_(ctx, name, st);                                \
        kern<<<grid, block, smem, st>>>(__VA_ARGS__);                  \
    } while (0)

void add_work(wgs_ctx* ctx, const char* name, double bytes, double units)
{
    if (!ctx->timing) return;
    auto& slot = ctx->tdone[name];
    slot.bytes += bytes; slot.units += units;
}

// Besides the per-family kernel time, the idle time of the stream between consecutive timed launches is kept under
// "gap" (all of it) and "gap_long" (the gaps above 50 us: host round trips, allocations, copies in between).
void fold_timing(wgs_ctx* ctx)
{
    for (size_t i = 0; i < ctx->timed.size(); ++i) {
        auto& t = ctx->timed[i];
        cudaEventSynchronize(t.b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t.a, t.b);
        auto& slot = ctx->tdone[ctx->tnames[t.name_id]];
        slot.ms += ms; slot.launches += 1;
        if (i > 0 && ctx->timed[i - 1].st == t.st) {
            float gap = 0.f;
            if (cudaEventElapsedTime(&gap, ctx->timed[i - 1].b, t.a) == cudaSuccess && gap > 0.f) {
                auto& g = ctx->tdone["gap"];
                g.ms += gap; g.launches += 1;
                if (gap > 0.05f) { auto& gl = ctx->tdone["gap_long"]; gl.ms += gap; gl.launches += 1; }
            }
        }
    }
    for (auto& t : ctx->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    ctx->timed.clear();
}

// option "trace": wall-clock of the host-side phases of every API call on stderr
struct Trace {
    const char* what; std::chrono::steady_clock::time_point t0; bool on;
    Trace(const wgs_ctx* c, const char* w) : what(w), t0(std::chrono::steady_clock::now()), on(opt(c, "trace") != 0) {}
    void lap(const char* phase) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[wgs trace] %-14s %-18s %9.3f ms\n", what, phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// Device memory comes from a per-context caching pool: cudaMalloc / cudaFree of multi-GB
// temporaries on every call cost tens of milliseconds (cudaFree synchronises the device),
// more than the kernels of a small problem.  All work is stream-ordered on ctx->stream, so
// a block can be handed to the next user as soon as the host releases it.
size_t pool_round(size_t bytes) { return bytes < (1u << 20) ? ((bytes + 511) / 512) * 512 : ((bytes + (1u << 20) - 1) >> 20) << 20; }

void pool_trim(wgs_ctx* ctx)
{
    for (auto& kv : ctx->pool_free) cudaFree(kv.second);
    if (!ctx->pool_free.empty()) ++ctx->pool_epoch;
    ctx->pool_free.clear();
}

// Free device memory as of the pool's last cudaMalloc / cudaFree.  cudaMemGetInfo is a driver round trip that takes
// 0.5 .. 12 ms (measured inside the leave-one-out call, the stream idle meanwhile), so it is asked again only after
// the pool itself changed what is allocated; memory another allocator of the process takes in between shows up as a
// failed allocation, which every caller handles.
size_t device_free_bytes(wgs_ctx* ctx)
{
    if (ctx->mem_epoch_seen != ctx->pool_epoch) {
        size_t freeb = 0, totalb = 0;
        cudaMemGetInfo(&freeb, &totalb);
        ctx->mem_free_seen = freeb;
        ctx->mem_epoch_seen = ctx->pool_epoch;
    }
    return ctx->mem_free_seen;
}

void* pool_take(wgs_ctx* ctx, size_t bytes)
{
    bytes = pool_round(std::max<size_t>(bytes, 16));
    auto it = ctx->pool_free.lower_bound(bytes);
    if (it != ctx->pool_free.end() && it->first <= bytes + bytes / 4 + (1u << 20)) {
        void* p = it->second;
        ctx->pool_size[p] = it->first;
        ctx->pool_free.erase(it);
        return p;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        pool_trim(ctx);                                       // give everything cached back and retry once
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    ctx->pool_size[p] = bytes;
    ++ctx->pool_epoch;
    return p;
}

void pool_give(wgs_ctx* ctx, void* p)
{
    if (!p) return;
    auto it = ctx->pool_size.find(p);
    if (it == ctx->pool_size.end()) { cudaFree(p); ++ctx->pool_epoch; return; }
    ctx->pool_free.insert({it->second, p});
    ctx->pool_size.erase(it);
}

template <class T> int dev_alloc(wgs_ctx* ctx, T** p, size_t count)
{
    *p = (T*)pool_take(ctx, std::max<size_t>(count, 1) * sizeof(T));
    if (!*p) return fail(ctx, "out of device memory allocating %zu bytes", std::max<size_t>(count, 1) * sizeof(T));
    return 0;
}
template <class T> void dev_free(wgs_ctx* ctx, T*& p) { pool_give(ctx, p); p = nullptr; }

struct DevBuf {   // RAII for temporaries (returned to the pool)
    void* p = nullptr;
    wgs_ctx* owner = nullptr;
    ~DevBuf() { if (p) pool_give(owner, p); }
    template <class T> T* as() { return (T*)p; }
};
struct RawBuf {   // non-owning view with the DevBuf accessors
    void* p;
    template <class T> T* as() { return (T*)p; }
};
void buf_release(DevBuf& b) { if (b.p) pool_give(b.owner, b.p); b.p = nullptr; }
int buf_alloc(wgs_ctx* ctx, DevBuf& b, size_t bytes)
{
    if (b.p) { pool_give(b.owner, b.p); b.p = nullptr; }
    b.owner = ctx;
    b.p = pool_take(ctx, bytes);
    if (!b.p) return fail(ctx, "out of device memory allocating %zu bytes", bytes);
    return 0;
}

int grid_for(long n, int block, int cap) { return (int)std::max<long>(1, std::min<long>((n + block - 1) / block, cap)); }

// ---- structure ---------------------------------------------------------------------------
int build_structure(wgs_ctx* ctx, const int32_t* pop_of_ind, int N, int K)
{
    ctx->N = N; ctx->K = K;
    ctx->pop_of_ind.assign(N, 0);
    int Ks = std::max(K, 1);
    if (K > 0) for (int i = 0; i < N; ++i) {
        if (pop_of_ind[i] < 0 || pop_of_ind[i] >= K) return fail(ctx, "pop_of_ind[%d]=%d outside [0,%d)", i, pop_of_ind[i], K);
        ctx->pop_of_ind[i] = pop_of_ind[i];
    }
    std::vector<int> cnt(Ks, 0);
    for (int i = 0; i < N; ++i) ++cnt[ctx->pop_of_ind[i]];
    ctx->pops.assign(Ks, PopDesc{0, 0});
    int col = 0;
    for (int k = 0; k < Ks; ++k) {
        ctx->pops[k].col0 = col; ctx->pops[k].n = cnt[k];
        col += (cnt[k] + 3) / 4 * 4;                 // 32-byte aligned slabs
    }
    ctx->ldg = std::max(col, 4);
    ctx->ind_of_col.assign(ctx->ldg, -1);
    ctx->pop_of_col.assign(ctx->ldg, 0);
    ctx->col_of_ind.assign(N, 0);
    std::vector<int> fillp(Ks, 0);
    for (int i = 0; i < N; ++i) {
        int k = ctx->pop_of_ind[i];
        int c = ctx->pops[k].col0 + fillp[k]++;
        ctx->col_of_ind[i] = c; ctx->ind_of_col[c] = i;
    }
    for (int k = 0; k < Ks; ++k) {
        int end = (k + 1 < Ks) ? ctx->pops[k + 1].col0 : ctx->ldg;
        for (int c = ctx->pops[k].col0; c < end; ++c) ctx->pop_of_col[c] = k;
    }
    dev_free(ctx, ctx->d_ind_of_col); dev_free(ctx, ctx->d_col_of_ind); dev_free(ctx, ctx->d_pop_of_col); dev_free(ctx, ctx->d_pops);
    if (dev_alloc(ctx, &ctx->d_ind_of_col, ctx->ldg) || dev_alloc(ctx, &ctx->d_col_of_ind, N) ||
        dev_alloc(ctx, &ctx->d_pop_of_col, ctx->ldg) || dev_alloc(ctx, &ctx->d_pops, Ks)) return 1;
    CU(cudaMemcpy(ctx->d_ind_of_col, ctx->ind_of_col.data(), ctx->ldg * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_col_of_ind, ctx->col_of_ind.data(), N * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_pop_of_col, ctx->pop_of_col.data(), ctx->ldg * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_pops, ctx->pops.data(), Ks * sizeof(PopDesc), cudaMemcpyHostToDevice));
    return 0;
}

// Host-blocking end of an asynchronous upload: every operator but the pipelined one starts with it.
void upload_fence(wgs_ctx* ctx)
{
    if (!ctx->upload_pending) return;
    cudaStreamSynchronize(ctx->stream2);
    ctx->upload_pending = false;
}

void drop_data(wgs_ctx* ctx)
{
    upload_fence(ctx);
    dev_free(ctx, ctx->G[0]); dev_free(ctx, ctx->G[1]); dev_free(ctx, ctx->AD); dev_free(ctx, ctx->d_af);
    ctx->Mg[0] = ctx->Mg[1] = ctx->M_ad = 0; ctx->af_rows = 0; ctx->af_cols = 0;
}

int ensure_structure(wgs_ctx* ctx, int N)
{
    if (ctx->N == N && ctx->ldg > 0) return 0;
    if (ctx->pops_set) return fail(ctx, "matrix has %d individuals but wgs_set_pops was given %d", N, ctx->N);
    drop_data(ctx);
    return build_structure(ctx, nullptr, N, 0);
}

// chunked host -> device with a column permutation kernel per chunk (two staging buffers)
template <class Elem, class Fn>
int upload_rows(wgs_ctx* ctx, const Elem* host, long M, int N, Fn&& repack)
{
    const size_t row_bytes = (size_t)N * sizeof(Elem);
    long chunk = std::max<long>(1, (long)((size_t)(96u << 20) / row_bytes));
    chunk = std::min(chunk, std::max<long>(M, 1));
    Elem* stage[2] = {nullptr, nullptr};
    cudaStream_t st[2] = {ctx->stream, ctx->stream2};
    for (int b = 0; b < 2; ++b) {
        stage[b] = (Elem*)pool_take(ctx, (size_t)chunk * row_bytes);
        if (!stage[b]) return fail(ctx, "out of device memory for the upload staging buffers");
    }
    int b = 0;
    int rc = 0;
    for (long r0 = 0; r0 < M && !rc; r0 += chunk, b ^= 1) {
        long rows = std::min(chunk, M - r0);
        cudaError_t e = cudaMemcpyAsync(stage[b], host + (size_t)r0 * N, (size_t)rows * row_bytes, cudaMemcpyHostToDevice, st[b]);
        if (e != cudaSuccess) { rc = fail(ctx, "H2D copy failed: %s", cudaGetErrorString(e)); break; }
        repack(stage[b], r0, rows, st[b]);
    }
    cudaStreamSynchronize(st[0]); cudaStreamSynchronize(st[1]);
    pool_give(ctx, stage[0]); pool_give(ctx, stage[1]);
    if (rc) return rc;
    CU(cudaGetLastError());
    return 0;
}

// ---- likelihood launch configuration --------------------------------------------------------
struct LikeCfg { int wx, gx, gy; long sites_per_block; };
LikeCfg like_cfg(wgs_ctx* ctx, long M, int blocks_per_sm, int inds_per_thread = 1)
{
    LikeCfg c;
    int groups = (ctx->ldg + 32 * inds_per_thread - 1) / (32 * inds_per_thread);
    // column groups per block: the largest power of two that wastes < 7 % of the warp slots on columns past ldg
    c.wx = 1;
    for (int w = 2; w <= 8; w *= 2)
        if ((groups + w - 1) / w * w <= groups + groups * 7 / 100) c.wx = w;
    c.gx = (groups + c.wx - 1) / c.wx;
    long target = (long)ctx->num_sm * blocks_per_sm * 2;
    long gy = std::max<long>(1, target / c.gx);
    long max_gy = std::max<long>(1, (M + kPopLikeTS - 1) / kPopLikeTS);
    gy = std::min(gy, max_gy);
    long spb = (M + gy - 1) / gy;
    spb = (spb + kPopLikeTS - 1) / kPopLikeTS * kPopLikeTS;
    if (spb < kPopLikeTS) spb = kPopLikeTS;
    c.sites_per_block = spb;
    c.gy = (int)std::max<long>(1, (M + spb - 1) / spb);
    return c;
}

int pick_R(float margin)
{
    // like >= ~margin^2; R factors on a mantissa in [1,2) must stay above 2^-120 (FP32 normal range ends at 2^-126)
    if (!(margin > 0.f)) return 1;
    double bits = -2.0 * std::log2((double)margin);
    if (bits < 1.0) bits = 1.0;
    int r = (int)(120.0 / bits);
    if (r >= 8) return 8;
    if (r >= 4) return 4;
    if (r >= 2) return 2;
    return 1;
}

const int kKT[] = {2, 4, 5, 8, 10, 16, 20};
int pick_KT(int Krem)
{
    for (int kt : kKT) if (kt >= Krem) return kt;
    return 20;
}

template <int KT, int R, int I>
int launch_pop_like_i(wgs_ctx* ctx, const float2* G, long M, const float* dA, int K, int k0, const LikeCfg& c, double* partials)
{
    constexpr int KP = (KT + 1) / 2;
    size_t smem = (size_t)kPopLikeTS * KP * (sizeof(ulonglong2) + sizeof(f32x2)) + 8 * 32 * 4 * sizeof(double);
    auto kern = pop_like_kernel<KT, R, I>;
    if (smem > 40 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH("pop_like", kern, dim3(c.gx, c.gy), kPopLikeThreads, smem, ctx->stream,
           G, ctx->ldg, M, dA, K, k0, c.wx, c.sites_per_block, partials);
    {   // algorithmic: every (g0,g1) pair once per pass + this pass's AF columns; one evaluation per (site, ind, pop)
        int kt = std::min(KT, K - k0);
        add_work(ctx, "pop_like", (double)M * ctx->N * 8.0 + (double)M * kt * 4.0, (double)M * ctx->N * kt);
    }
    return 0;
}
// individuals per thread: as many as the register file allows for the tile width, 1 when the matrix is narrow
constexpr int pop_like_imax(int KT) { return KT <= 10 ? 2 : 1; }
int pop_like_inds(int KT, int ldg)
{
    return ldg > 64 ? pop_like_imax(KT) : 1;
}
template <int KT, int R>
int launch_pop_like_t(wgs_ctx* ctx, const float2* G, long M, const float* dA, int K, int k0, const LikeCfg& c, double* partials, int I)
{
    if (I == 1) return launch_pop_like_i<KT, R, 1>(ctx, G, M, dA, K, k0, c, partials);
    return launch_pop_like_i<KT, R, pop_like_imax(KT)>(ctx, G, M, dA, K, k0, c, partials);
}
template <int KT, int R>
int launch_loo_like_t(wgs_ctx* ctx, const float2* G, long M, const float* Fx, int ldf, const int* rc, int K, int k0,
                      const LikeCfg& c, long pm, long pr, double* partials)
{
    size_t smem = 8 * 32 * 4 * sizeof(double);
    LAUNCH("loo_like", (loo_like_kernel<KT, R>), dim3(c.gx, c.gy), kPopLikeThreads, smem, ctx->stream,
           G, ctx->ldg, M, Fx, ldf, rc, K, k0, c.wx, c.sites_per_block, pm, pr, ctx->site_offset, partials);
    {   // GL pairs once + the LOO state row (one float per individual) + full-data AF columns
        int kt = std::min(KT, K - k0);
        add_work(ctx, "loo_like", (double)M * ctx->N * 12.0 + (double)M * kt * 4.0, (double)M * ctx->N * kt);
    }
    return 0;
}

// ---- ratio-form likelihood kernel (pop_like2) --------------------------------------------------
struct Like2Cfg { int R, gy; long spb; };
const int kKT2[] = {4, 8, 10, 16, 20};
int pick_KT2(int Krem)
{
    for (int kt : kKT2) if (kt >= Krem) return kt;
    return 20;
}
// R sites between renormalisations: every per-site factor lies in [m/2, 1/(2m)]; R of them on a
// mantissa in [1,2) must stay inside 2^+-120.  0 = margin too small for the ratio form.
int pick_R2(float margin)
{
    if (!(margin >= 9.3132257e-10f)) return 0;                 // 2^-30 (also NaN / <= 0)
    double bits = std::log2(1.0 / (double)margin) - 1.0;
    if (bits < 1.0) bits = 1.0;
    if (16.0 * bits <= 120.0) return 16;
    if (8.0 * bits <= 120.0) return 8;
    return 0;
}
Like2Cfg like2_cfg(wgs_ctx* ctx, long M, int R)
{
    Like2Cfg c;
    c.R = R;
    long target = (long)ctx->num_sm * 4;                          // ~2 waves of 2 resident blocks per SM
    long spb = (M + target - 1) / target;
    spb = std::max<long>(kPL2TS, (spb + kPL2TS - 1) / kPL2TS * kPL2TS);
    spb = std::min<long>(spb, 256L * R);                          // 16-bit exponent fields: at most 256 renormalisations per thread
    c.spb = spb;
    c.gy = (int)std::max<long>(1, (M + spb - 1) / spb);
    return c;
}
template <int KT, int I, int R>
int launch_pop_like2_i(wgs_ctx* ctx, const float2* G, long M, const float* dA, int K, int k0, const Like2Cfg& c, double* partials)
{
    constexpr int KP = (KT + 1) / 2;
    const int groups = (ctx->ldg + 32 * I - 1) / (32 * I);
    // warps (column groups) per block: an exited warp keeps its registers until the whole block retires, so
    // padding warps cost occupancy (9 groups in blocks of 4: 1.35 ms; in blocks of 1: 1.06 ms) - least padding wins
    int wx = 1, best_waste = 1 << 30;
    for (int w = 1; w <= 4; ++w) {
        int waste = (groups + w - 1) / w * w - groups;
        if (waste <= best_waste) { best_waste = waste; wx = w; }
    }
    if (int o = opt(ctx, "pl2_wx", 0)) wx = std::max(1, std::min(4, o));
    const int gx = (groups + wx - 1) / wx;
    // per-pass coefficient rows, padded to whole tiles so that the staging copies never run past the end
    const long Mpad = (M + kPL2TS - 1) / kPL2TS * kPL2TS;
    DevBuf xy;
    if (buf_alloc(ctx, xy, (size_t)std::max<long>(Mpad, kPL2TS) * KP * sizeof(ulonglong2))) return 1;
    LAUNCH("pop_like_aux", xy_precompute_kernel, grid_for(M * KP, 256, ctx->num_sm * 8), 256, 0, ctx->stream, dA, M, K, k0, KP, xy.as<ulonglong2>());
    const size_t smem = (size_t)wx * 2 * kPL2TS * KP * sizeof(ulonglong2);
    auto kern = pop_like2_kernel<KT, I, R>;
    if (smem > 40 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    LAUNCH("pop_like", kern, dim3(gx, c.gy), wx * 32, smem, ctx->stream,
           G, ctx->ldg, M, xy.as<ulonglong2>(), K, k0, c.spb, partials);
    {   // algorithmic: every (g0,g1) pair once per pass + this pass's AF columns; one evaluation per (site, ind, pop)
        int kt = std::min(KT, K - k0);
        add_work(ctx, "pop_like", (double)M * ctx->N * 8.0 + (double)M * kt * 4.0, (double)M * ctx->N * kt);
    }
    CU(cudaStreamSynchronize(ctx->stream));               // xy is released on return
    return 0;
}
template <int KT, int I>
int launch_pop_like2_r(wgs_ctx* ctx, const float2* G, long M, const float* dA, int K, int k0, const Like2Cfg& c, double* partials)
{
    if (c.R == 16) return launch_pop_like2_i<KT, I, 16>(ctx, G, M, dA, K, k0, c, partials);
    return launch_pop_like2_i<KT, I, 8>(ctx, G, M, dA, K, k0, c, partials);
}
int launch_pop_like2(wgs_ctx* ctx, int KT, const float2* G, long M, const float* dA, int K, int k0, const Like2Cfg& c, double* partials)
{
    switch (KT) {
        case 4: return launch_pop_like2_r<4, 2>(ctx, G, M, dA, K, k0, c, partials);
        case 8: return launch_pop_like2_r<8, 2>(ctx, G, M, dA, K, k0, c, partials);
        case 10: return launch_pop_like2_r<10, 2>(ctx, G, M, dA, K, k0, c, partials);
        case 16: return launch_pop_like2_r<16, 1>(ctx, G, M, dA, K, k0, c, partials);
        default: return launch_pop_like2_r<20, 1>(ctx, G, M, dA, K, k0, c, partials);
    }
}
// sums[col][k] += sum over this rank's sites of log(2 a (1-a)) for population k
int add_af_logsum(wgs_ctx* ctx, const float* dA, long M, int K, double* sums, long np)
{
    // blocks of a whole multiple of K threads: thread t only ever visits population t % K
    const int block = K <= 256 ? K * (256 / K) : K;
    const int grid = ctx->num_sm * 4;
    DevBuf per, C;
    if (buf_alloc(ctx, per, (size_t)grid * K * sizeof(double)) || buf_alloc(ctx, C, (size_t)K * sizeof(double))) return 1;
    LAUNCH("pop_like_aux", af_logsum_kernel, grid, block, block * sizeof(double), ctx->stream, dA, M * K, K, per.as<double>());
    LAUNCH("pop_like_aux", af_logsum_reduce_kernel, (K + 127) / 128, 128, 0, ctx->stream, per.as<double>(), grid, K, C.as<double>());
    LAUNCH("pop_like_aux", add_pop_const_kernel, grid_for(np, 256, ctx->num_sm * 4), 256, 0, ctx->stream, sums, np, K, C.as<double>());
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- leave-one-out likelihoods with the state rows staged in shared memory (loo_like2) -----------
struct LooLike2Cfg { int W, gx, gy, TS; long spb; size_t smem; bool wide, big; };
bool loo_like2_cfg(wgs_ctx* ctx, long M, int ldf, int K, LooLike2Cfg* c)
{
    const int groups = (ctx->ldg + 31) / 32;
    c->wide = K > 10;                                             // one block per SM, <= 4 sites per tile (see the kernel)
    // more than 10 column groups and a narrow population tile: big blocks (up to 18 warps, one per SM) build the cells of
    // a state row once per 18 warps of individuals instead of once per 10
    c->big = !c->wide && groups > kLL2MaxW && !opt(ctx, "loolike_smallblock");
    const int maxw = c->wide ? kLL2WideW : (c->big ? kLL2BigW : kLL2MaxW);
    const int nb = (groups + maxw - 1) / maxw;                    // blocks per site split
    c->W = (groups + nb - 1) / nb;
    c->gx = nb;
    const size_t budget = (c->wide || c->big) ? 210 * 1024 : 100 * 1024;
    // per site of a tile: 16 B plane cell + 4 B landing row per state value, 2 x 8 B per GL column of the block
    const size_t per_site = (size_t)ldf * 20 + (size_t)c->W * 32 * 16;
    const size_t fixed = (size_t)ldf * 8;                         // clamp bounds
    if (budget <= fixed) return false;
    int TS = (int)std::min<size_t>(c->wide ? 4 : 8, (budget - fixed) / per_site);
    if (TS < 1) return false;
    c->TS = TS;
    c->smem = (size_t)TS * per_site + fixed;
    long target = std::max<long>(1, (long)ctx->num_sm * ((c->wide || c->big) ? 1 : 2) / nb);    // one wave of resident blocks
    long spb = (M + target - 1) / target;
    spb = std::max<long>(TS, (spb + TS - 1) / TS * TS);
    c->spb = spb;
    c->gy = (int)std::max<long>(1, (M + spb - 1) / spb);
    return true;
}
template <int KT, int TSMAX, int MINB, int MAXW = kLL2MaxW>
int launch_loo_like2_t(wgs_ctx* ctx, const float2* G, long M, const float* Fx, int ldf, const float* clo, const float* chi,
                       const int* rc, int K, int k0, const LooLike2Cfg& c, long pm, long pr, int R, double* partials)
{
    auto kern = loo_like2_kernel<KT, TSMAX, MINB, MAXW>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    LAUNCH("loo_like", kern, dim3(c.gx, c.gy), c.W * 32, c.smem, ctx->stream,
           G, ctx->ldg, M, Fx, ldf, clo, chi, rc, K, k0, c.TS, c.spb, pm, pr, ctx->site_offset, R, partials);
    {   // GL pairs once + the LOO state row (one float per individual) + full-data AF columns
        int kt = std::min(KT, K - k0);
        add_work(ctx, "loo_like", (double)M * ctx->N * 12.0 + (double)M * kt * 4.0, (double)M * ctx->N * kt);
    }
    return 0;
}
// population tile width of one pass: the narrow kernels take pick_KT's widths, the wide ones 16 or 20 (masked past K)
int loo_like2_KT(const LooLike2Cfg& c, int remaining) { return c.wide ? (remaining > 16 ? 20 : 16) : pick_KT(remaining); }
int launch_loo_like2(wgs_ctx* ctx, int KT, const float2* G, long M, const float* Fx, int ldf, const float* clo, const float* chi,
                     const int* rc, int K, int k0, const LooLike2Cfg& c, long pm, long pr, int R, double* partials)
{
    if (c.wide) {
        if (KT == 16) return launch_loo_like2_t<16, 4, 1, kLL2WideW>(ctx, G, M, Fx, ldf, clo, chi, rc, K, k0, c, pm, pr, R, partials);
        return launch_loo_like2_t<20, 4, 1, kLL2WideW>(ctx, G, M, Fx, ldf, clo, chi, rc, K, k0, c, pm, pr, R, partials);
    }
    if (c.big) {
        switch (KT) {
            case 2: return launch_loo_like2_t<2, 8, 1, kLL2BigW>(ctx, G, M, Fx, ldf, clo, chi, rc, K, k0, c, pm, pr, R, partials);
            case 4: return launch_loo_like2_t<4, 8, 1, kLL2BigW>(ctx, G, M, Fx, ldf, clo, chi, rc, K, k0, c, pm, pr, R, partials);
            case 5: return launch_loo_like2_t<5, 8, 1, kLL2BigW>(ctx, G, M, Fx, ldf, clo, chi, rc, K, k0, c, pm, pr, R, partials);
            case 8: return launch_loo_like2_t<8, 8, 1, kLL2BigW>(ctx, G, M, Fx, ldf, clo, chi, rc, K, k0, c, pm, pr, R, partials);
            default: return launch_loo_like2_t<10, 8, 1, kLL2BigW>(ctx, G, M, Fx, ldf, clo, chi, rc, K, k0, c, pm, pr, R, partials);
        }
    }
    switch (KT) {
        case 2: return launch_loo_like2_t<2, 8, 2>(ctx, G, M, Fx, ldf, clo, chi, rc, K, k0, c, pm, pr, R, partials);
        case 4: return launch_loo_like2_t<4, 8, 2>(ctx, G, M, Fx, ldf, clo, chi, rc, K, k0, c, pm, pr, R, partials);
        case 5: return launch_loo_like2_t<5, 8, 2>(ctx, G, M, Fx, ldf, clo, chi, rc, K, k0, c, pm, pr, R, partials);
        case 8: return launch_loo_like2_t<8, 8, 2>(ctx, G, M, Fx, ldf, clo, chi, rc, K, k0, c, pm, pr, R, partials);
        default: return launch_loo_like2_t<10, 8, 2>(ctx, G, M, Fx, ldf, clo, chi, rc, K, k0, c, pm, pr, R, partials);
    }
}

// ---- leave-one-out likelihoods of a run-ordered panel (loo_like3) -----------------------------------------
struct LooLike3Plan {
    int KT = 0, KP = 0, R = 0, r_own = 1, wx = 1, gx = 1;
    Like2Cfg c2{};
    std::vector<int> slot_of_col, pop_of_slot, lastcol_of_slot;
};
// populations in contiguous runs of the ID file, at least two members each, K <= 20, and every frequency the kernel
// can read far enough from 0 and 1 for the ratio form (margin): else the general kernels take the call
bool plan_loo_like3(wgs_ctx* ctx, long M, float margin, int r_own, LooLike3Plan* p)
{
    const int K = ctx->K, N = ctx->N, ldg = ctx->ldg;
    if (K < 2 || K > 20 || opt(ctx, "loolike_v1") || opt(ctx, "loolike_v2")) return false;
    std::vector<int> slot_of_pop(K, -1), last_ind(K, -1);
    int nslots = 0;
    for (int i = 0; i < N; ++i) {
        const int k = ctx->pop_of_ind[i];
        if (slot_of_pop[k] < 0) slot_of_pop[k] = nslots++;
        else if (ctx->pop_of_ind[i - 1] != k) return false;       // population k resumes after another one: not run-ordered
        last_ind[k] = i;
    }
    if (nslots != K) return false;
    for (int k = 0; k < K; ++k) if (ctx->pops[k].n < 2) return false;
    p->R = pick_R2(margin);
    if (p->R == 0) return false;
    p->r_own = std::max(1, r_own);
    p->KT = pick_KT2(K); p->KP = (p->KT + 1) / 2;
    p->pop_of_slot.assign(K, 0); p->lastcol_of_slot.assign(K, 0);
    for (int k = 0; k < K; ++k) { p->pop_of_slot[slot_of_pop[k]] = k; p->lastcol_of_slot[slot_of_pop[k]] = ctx->col_of_ind[last_ind[k]]; }
    p->slot_of_col.assign(ldg, 0);
    for (int c = 0; c < ldg; ++c) p->slot_of_col[c] = slot_of_pop[ctx->pop_of_col[c]];
    const int groups = (ldg + 63) / 64;                           // 64 columns per warp (two adjacent columns per thread)
    const int wmax = p->KT > 10 ? 2 : 4;                          // 20 KB of coefficient tiles per warp at KT = 20
    int best_waste = 1 << 30;
    for (int w = 1; w <= wmax; ++w) {
        const int waste = (groups + w - 1) / w * w - groups;
        if (waste <= best_waste) { best_waste = waste; p->wx = w; }
    }
    p->gx = (groups + p->wx - 1) / p->wx;
    p->c2 = like2_cfg(ctx, M, p->R);
    return true;
}
template <int KT, int R>
int launch_loo_like3_t(wgs_ctx* ctx, const LooLike3Plan& p, const float2* G, long M, const float* F, int ldf, const float* clo, const float* chi,
                       const ulonglong2* XY, const int* d_slot_of_col, const int* d_pop_of_slot, long pm, long pr, double* partials)
{
    constexpr int KP = (KT + 1) / 2;
    const size_t smem = (size_t)p.wx * 2 * kPL2TS * 2 * KP * sizeof(ulonglong2);
    auto kern = loo_like3_kernel<KT, R>;
    if (smem > 40 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    LAUNCH("loo_like", kern, dim3(p.gx, p.c2.gy), p.wx * 32, smem, ctx->stream, G, ctx->ldg, M, F, ldf, clo, chi, XY, d_slot_of_col,
           d_pop_of_slot, ctx->K, p.c2.spb, pm, pr, ctx->site_offset, p.r_own, partials);
    // GL pairs once + the leave-one-out state (one float per individual) + the 2K shared frequencies of the site
    add_work(ctx, "loo_like", (double)M * ctx->N * 12.0 + (double)M * ctx->K * 8.0, (double)M * ctx->N * ctx->K);
    return 0;
}
int launch_loo_like3(wgs_ctx* ctx, const LooLike3Plan& p, const float2* G, long M, const float* F, int ldf, const float* clo, const float* chi,
                     const ulonglong2* XY, const int* dsc, const int* dps, long pm, long pr, double* partials)
{
#define LL3(KTV)                                                                                                               \
    return p.R == 16 ? launch_loo_like3_t<KTV, 16>(ctx, p, G, M, F, ldf, clo, chi, XY, dsc, dps, pm, pr, partials)             \
                     : launch_loo_like3_t<KTV, 8>(ctx, p, G, M, F, ldf, clo, chi, XY, dsc, dps, pm, pr, partials)
    switch (p.KT) {
        case 4: LL3(4);
        case 8: LL3(8);
        case 10: LL3(10);
        case 16: LL3(16);
        default: LL3(20);
    }
#undef LL3
}

#define DISPATCH_R(FN, KTV, R, ...)                                      \
    switch (R) {                                                         \
        case 8: rc_ = FN<KTV, 8>(__VA_ARGS__); break;                    \
        case 4: rc_ = FN<KTV, 4>(__VA_ARGS__); break;                    \
        case 2: rc_ = FN<KTV, 2>(__VA_ARGS__); break;                    \
        default: rc_ = FN<KTV, 1>(__VA_ARGS__); break;                   \
    }
#define DISPATCH_KT(FN, KT, R, ...)                                      \
    do {                                                                 \
        switch (KT) {                                                    \
            case 2: DISPATCH_R(FN, 2, R, __VA_ARGS__) break;             \
            case 4: DISPATCH_R(FN, 4, R, __VA_ARGS__) break;             \
            case 5: DISPATCH_R(FN, 5, R, __VA_ARGS__) break;             \
            case 8: DISPATCH_R(FN, 8, R, __VA_ARGS__) break;             \
            case 10: DISPATCH_R(FN, 10, R, __VA_ARGS__) break;           \
            case 16: DISPATCH_R(FN, 16, R, __VA_ARGS__) break;           \
            default: DISPATCH_R(FN, 20, R, __VA_ARGS__) break;           \
        }                                                                \
    } while (0)

// margin of a device AF matrix -> renormalisation interval
int af_R(wgs_ctx* ctx, const float* dA, long n, int* R_out, float* margin_out = nullptr)
{
    DevBuf b;
    if (buf_alloc(ctx, b, sizeof(int))) return 1;
    int init = 0x3f000000;   // 0.5f
    CU(cudaMemcpyAsync(b.p, &init, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH("af_margin", af_margin_kernel, grid_for(n, 256, ctx->num_sm * 8), 256, 0, ctx->stream, dA, n, b.as<int>());
    int bits = 0;
    CU(cudaMemcpyAsync(&bits, b.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float m;
    memcpy(&m, &bits, 4);
    *R_out = pick_R(m);
    if (margin_out) *margin_out = m;
    return 0;
}

// Cross-rank sums of the small per-individual / per-population results, on the device: the ranks' local values are
// all-gathered over the context's own communicator (NVLink) and added IN RANK ORDER, so every rank holds the same
// bits and they do not depend on NCCL's reduction schedule.  Without a communicator (or with option nccl_sums = 0)
// the operators return local partial sums and the caller combines them (wgs_partials_combined() tells which).
bool nccl_sums(const wgs_ctx* ctx) { return ctx->fn && ctx->nccl_comm && opt(ctx, "nccl_sums", 1) != 0; }
int dev_all_sum(wgs_ctx* ctx, void* dbuf, size_t n, int dtype)
{
    DevBuf g;
    if (buf_alloc(ctx, g, (size_t)ctx->nccl_world * n * 8)) return 1;
    int rc_ = g_nccl.AllGather(dbuf, g.p, n, dtype == WGS_F64 ? kNcclFloat64 : kNcclInt64, ctx->nccl_comm, ctx->stream);
    if (rc_) return fail(ctx, "ncclAllGather failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc_) : "?");
    if (dtype == WGS_F64)
        LAUNCH("rank_sum", em_rank_sum_kernel, grid_for((long)n, 128, ctx->num_sm * 4), 128, 0, ctx->stream, g.as<double>(), ctx->nccl_world, (int)n, (double*)dbuf);
    else
        LAUNCH("rank_sum", rank_sum_i64_kernel, grid_for((long)n, 128, ctx->num_sm * 4), 128, 0, ctx->stream, g.as<long long>(), ctx->nccl_world, (int)n, (long long*)dbuf);
    return 0;
}

// partial sums [ldg][K] on device -> host [N][K] in Beagle order (summed over the ranks first when the context has a communicator)
int cols_to_host(wgs_ctx* ctx, double* d_colsums, int K, double* out)
{
    if (nccl_sums(ctx) && dev_all_sum(ctx, d_colsums, (size_t)ctx->ldg * K, WGS_F64)) return 1;
    std::vector<double> h((size_t)ctx->ldg * K);
    CU(cudaMemcpyAsync(h.data(), d_colsums, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < ctx->N; ++i)
        memcpy(out + (size_t)i * K, h.data() + (size_t)ctx->col_of_ind[i] * K, K * sizeof(double));
    return 0;
}

// ---- EM loop ------------------------------------------------------------------------------
struct EmState {
    int np = 0;                  // problems (columns of the partials)
    int ld = 0;                  // leading dimension of partials
    int nblocks = 0;
    DevBuf partials, ssq, count, active, iters, gathered;
    std::vector<int> h_active, h_iters;
    int slot_c0[2] = {0, 0}, slot_nc[2] = {0, 0};
    // exact stop rule (the reference's sequential float32 sum, emMAF_cy.pyx:26-33) for checks that land inside the
    // band in which the exact FP64 sum cannot decide: D2 [ldg / 4][M][4] (quad-major) = the squared changes of the last iteration
    DevBuf d2, serial, carry, uncertain;
    bool exact = false;
    bool chain_on = false, chain_always = false, missed = false;   // site-sharded runs: the rank chain is queued only near convergence
    double band_override = 0.0;
    double near_factor = 30.0;   // "near": within this factor of the tolerance (two iterations of a 0.3x contraction above the band)
};

// The reference's stop sum is sequential float32; `exact` runs keep what is needed to reproduce it (option
// rmse_exact = 0 decides on the FP64 sums alone).  rmse_band_ppm selects the band inside which a check is resolved
// sequentially (em_band, wgs_kernels.cuh): 0 = 8x the measured bias of the float32 sum, > 0 = ppm of the tolerance,
// -1 = every check (tests), -2 = the rigorous summation bound.
double band_override_of(const wgs_ctx* ctx)
{
    const int ppm = opt(ctx, "rmse_band_ppm", 0);
    return ppm == 0 ? 0.0 : (ppm < 0 ? (double)ppm : ppm * 1e-6);   // -1: every check, -2: the rigorous bound (see em_band)
}
bool em_uncertain(double ssq, double count, double tole, double band_override)
{
    float res = (float)ssq;
    res = res / (float)count;
    const double diff = std::sqrt((double)res);
    const double band = em_band(count, band_override);
    return band < 0.0 || std::fabs(diff - tole) <= band * tole;
}

// The decision kernels write straight into mapped pinned host memory (no copy to queue, nothing pageable on the
// path); the slots live in the context because cudaHostAlloc / cudaFreeHost synchronise the device.
int em_pin_reserve(wgs_ctx* ctx, size_t ints)
{
    if (ints <= ctx->em_pin_cap) return 0;
    CU(cudaDeviceSynchronize());
    size_t cap = std::max<size_t>(ints, 1 << 16);
    for (int b = 0; b < 2; ++b) {
        if (ctx->em_pin[b]) cudaFreeHost(ctx->em_pin[b]);
        ctx->em_pin[b] = nullptr;
        CU(cudaHostAlloc((void**)&ctx->em_pin[b], cap * sizeof(int), cudaHostAllocMapped));
        CU(cudaHostGetDevicePointer((void**)&ctx->em_pin_dev[b], ctx->em_pin[b], 0));
        if (!ctx->em_ev[b]) CU(cudaEventCreateWithFlags(&ctx->em_ev[b], cudaEventDisableTiming));
    }
    ctx->em_pin_cap = cap;
    return 0;
}

int em_state_init(wgs_ctx* ctx, EmState& st, int np, int ld, int nblocks, const std::vector<int>& active0, bool exact = false, bool zero_d2 = true)
{
    st.np = np; st.ld = ld; st.nblocks = nblocks;
    if (buf_alloc(ctx, st.partials, (size_t)nblocks * ld * sizeof(double)) || buf_alloc(ctx, st.ssq, (size_t)np * sizeof(double)) ||
        buf_alloc(ctx, st.active, (size_t)np * sizeof(int)) || buf_alloc(ctx, st.iters, (size_t)np * sizeof(int))) return 1;
    if (em_pin_reserve(ctx, (size_t)np + 3)) return 1;
    CU(cudaMemsetAsync(st.partials.p, 0, (size_t)nblocks * ld * sizeof(double), ctx->stream));
    CU(cudaMemsetAsync(st.iters.p, 0, (size_t)np * sizeof(int), ctx->stream));
    st.h_active = active0;
    st.h_iters.assign(np, 0);
    CU(cudaMemcpyAsync(st.active.p, st.h_active.data(), (size_t)np * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    st.exact = exact;
    st.band_override = band_override_of(ctx);
    {   // the band for the whole-file site count decides how early the rank chain has to be queued; "every check" means
        // the chain is on from the first iteration
        const double band = em_band((double)ctx->Mtot(), st.band_override);
        if (band < 0.0) st.chain_always = true;
        else st.near_factor = 12.0 * (1.0 + band);
    }
    if (exact) {
        const long M = ctx->M();
        if (buf_alloc(ctx, st.d2, (size_t)std::max<long>(M, 1) * ctx->ldg * sizeof(float)) || buf_alloc(ctx, st.serial, (size_t)np * sizeof(float)) ||
            buf_alloc(ctx, st.carry, (size_t)np * sizeof(float)) || buf_alloc(ctx, st.uncertain, (size_t)np * sizeof(int))) return 1;
        if (zero_d2)                                             // masked sites are never written and must read as + 0; without a
            CU(cudaMemsetAsync(st.d2.p, 0, (size_t)std::max<long>(M, 1) * ctx->ldg * sizeof(float), ctx->stream));   // mask every entry of an active problem is rewritten each iteration
        CU(cudaMemsetAsync(st.serial.p, 0, (size_t)np * sizeof(float), ctx->stream));
    }
    return 0;
}

constexpr int kNcclFloat32 = 7;

// Hand `n` floats from rank to rank in site order: every rank runs launch(carry_in) - which must leave its running
// values in `vals` - after it has received the previous rank's; the last rank's values are then broadcast, so every
// rank returns with the whole-file values in `vals`.  With a communicator everything is queued on the compute
// stream (no host synchronisation); through the host callback the ranks take turns (one blocking round each).
template <class Launch>
int chain_floats(wgs_ctx* ctx, float* vals, float* carry, size_t n, Launch&& launch)
{
    if (!ctx->fn || ctx->world <= 1) return launch((const float*)nullptr);
    const int r = ctx->rank, W = ctx->world;
    if (ctx->nccl_comm) {
        int rc_ = 0;
        if (r > 0) rc_ = g_nccl.Recv(carry, n, kNcclFloat32, r - 1, ctx->nccl_comm, ctx->stream);
        if (!rc_ && launch(r > 0 ? (const float*)carry : (const float*)nullptr)) return 1;
        if (!rc_ && r < W - 1) rc_ = g_nccl.Send(vals, n, kNcclFloat32, r + 1, ctx->nccl_comm, ctx->stream);
        if (!rc_) rc_ = g_nccl.Broadcast(vals, vals, n, kNcclFloat32, W - 1, ctx->nccl_comm, ctx->stream);
        if (rc_) return fail(ctx, "NCCL hand-over of the sequential sums failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc_) : "?");
        return 0;
    }
    std::vector<long long> h(n);                                 // float bits travel as integers: a sum with one contributor is exact
    std::vector<float> hv(n);
    for (int q = 0; q < W; ++q) {
        if (q == r) {
            if (launch(q > 0 ? (const float*)carry : (const float*)nullptr)) return 1;
            CU(cudaMemcpyAsync(hv.data(), vals, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            for (size_t e = 0; e < n; ++e) { unsigned b; memcpy(&b, &hv[e], 4); h[e] = (long long)b; }
        } else std::fill(h.begin(), h.end(), 0LL);
        ctx->fn(h.data(), (int64_t)n, WGS_I64, ctx->user);
        for (size_t e = 0; e < n; ++e) { unsigned b = (unsigned)h[e]; memcpy(&hv[e], &b, 4); }
        float* dst = (q + 1 == W) ? vals : carry;
        if (q + 1 == r || q + 1 == W) CU(cudaMemcpyAsync(dst, hv.data(), n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));                  // hv is reused next round
    }
    return 0;
}

// After a step kernel has written st.partials: reduce, (all-reduce across ranks), decide.
// d_count: optional per-problem site counts (already global); else count_all.
// [c0, c0 + nc) restricts the reduction, the decision and the active count to those problems (nc < 0: all) -
// the pipelined leave-one-out EM advances one population at a time.
// Two halves: em_after_step_queue puts reduce + decide + the read-back of the decision on the stream (into result
// slot `slot`), em_after_step_wait blocks until that read-back has landed.  The caller may queue the NEXT
// iteration's kernels in between: a problem that this decision freezes is skipped on the device (the step kernels
// read the `active` flags), so a speculative launch never changes a result.
int em_after_step_queue(wgs_ctx* ctx, EmState& st, double tole, int iteration, const double* d_count, double count_all,
                        int slot, int c0 = 0, int nc = -1)
{
    if (nc < 0) nc = st.np;
    LAUNCH("em_ssq_reduce", em_ssq_reduce_kernel, (nc + 7) / 8, 256, 0, ctx->stream,
           st.partials.as<double>() + c0, st.nblocks, nc, st.ld, st.ssq.as<double>() + c0);
    if (ctx->fn && ctx->nccl_comm) {
        // all ranks' local sums, gathered on the stream and added in rank order: no host in the loop
        if (!st.gathered.p && buf_alloc(ctx, st.gathered, (size_t)ctx->nccl_world * st.np * sizeof(double))) return 1;
        int rc_ = g_nccl.AllGather(st.ssq.as<double>() + c0, st.gathered.p, (size_t)nc, kNcclFloat64, ctx->nccl_comm, ctx->stream);
        if (rc_) return fail(ctx, "ncclAllGather failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc_) : "?");
        LAUNCH("em_ssq_reduce", em_rank_sum_kernel, grid_for(nc, 128, 64), 128, 0, ctx->stream, st.gathered.as<double>(), ctx->nccl_world, nc,
               st.ssq.as<double>() + c0);
    } else if (ctx->fn) {
        std::vector<double> h(nc);
        CU(cudaMemcpyAsync(h.data(), st.ssq.as<double>() + c0, nc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->fn(h.data(), nc, WGS_F64, ctx->user);
        CU(cudaMemcpyAsync(st.ssq.as<double>() + c0, h.data(), nc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));                  // h is a local
    }
    const int* d_unc = nullptr;
    const float* d_serial = nullptr;
    if (st.exact) {
        // checks inside the band where the FP64 sum cannot decide are resolved with the reference's own sequential
        // float32 sum, rebuilt from D2 (block_seqsum32); under site sharding the running sums pass from rank to rank
        const double* cnt = d_count ? d_count + c0 : nullptr;
        auto resolve = [&](const float* carry_in, bool sums, int* any_flag = nullptr) -> int {
            LAUNCH("em_resolve", em_resolve_kernel, std::max(1, std::min(nc, ctx->num_sm * 16)), kSeqWarps * 32, 0, ctx->stream,
                   st.ssq.as<double>() + c0, cnt, count_all, nc, tole, st.band_override, st.active.as<int>() + c0,
                   sums ? st.d2.as<float>() + (size_t)c0 * std::max<long>(ctx->M(), 1) : (const float*)nullptr, ctx->M(), carry_in, st.serial.as<float>() + c0,
                   st.uncertain.as<int>() + c0, any_flag);
            return 0;
        };
        const bool sharded = ctx->fn && ctx->world > 1;
        bool have_sums = true;
        if (!sharded) {
            if (resolve(nullptr, true)) return 1;
        } else if (ctx->nccl_comm && !(st.chain_on || st.chain_always)) {
            // far from convergence: flags only, nothing leaves the device (a flag raised here restarts the EM, see run_em_loo)
            if (resolve(nullptr, false)) return 1;
            have_sums = false;
        } else {
            // Near convergence (or through the host callback, where the sums pass through the host anyway): look at the flags
            // first - they are the same on every rank, being derived from the whole-file FP64 sums - and run the rank chain
            // only for a check that needs it.  The read-back costs one stream synchronisation (the look-ahead launch is
            // already queued behind it); a chain costs a hop per rank.
            int* any_dev = ctx->em_pin_dev[slot] + (ctx->em_pin_cap - 1);         // last word of the slot: never used by a decision
            int* any_host = ctx->em_pin[slot] + (ctx->em_pin_cap - 1);
            *any_host = 0;
            if (resolve(nullptr, false, any_dev)) return 1;
            CU(cudaStreamSynchronize(ctx->stream));
            if (*any_host && chain_floats(ctx, st.serial.as<float>() + c0, st.carry.as<float>() + c0, (size_t)nc,
                                          [&](const float* cin) { return resolve(cin, true); })) return 1;
        }
        d_unc = st.uncertain.as<int>() + c0;
        d_serial = have_sums ? st.serial.as<float>() + c0 : nullptr;
    }
    LAUNCH("em_decide", em_decide_kernel, 1, 1024, 0, ctx->stream, st.ssq.as<double>() + c0, d_count ? d_count + c0 : nullptr, count_all, nc, tole,
           iteration, st.active.as<int>() + c0, st.iters.as<int>() + c0, ctx->em_pin_dev[slot], d_unc, d_serial, 1 + nc, st.near_factor);
    CU(cudaEventRecord(ctx->em_ev[slot], ctx->stream));
    st.slot_c0[slot] = c0; st.slot_nc[slot] = nc;
    return 0;
}
int em_after_step_wait(wgs_ctx* ctx, EmState& st, int slot, int* n_active)
{
    CU(cudaEventSynchronize(ctx->em_ev[slot]));
    *n_active = ctx->em_pin[slot][0];
    const int nc = st.slot_nc[slot];
    memcpy(st.h_active.data() + st.slot_c0[slot], ctx->em_pin[slot] + 1, (size_t)nc * sizeof(int));
    st.chain_on = ctx->em_pin[slot][1 + nc] != 0;                 // some problem is within 30x of the tolerance: queue the rank chain from now on
    if (ctx->em_pin[slot][2 + nc]) st.missed = true;              // a check needed the chain before it was on
    return 0;
}
int em_after_step(wgs_ctx* ctx, EmState& st, double tole, int iteration, const double* d_count, double count_all, int* n_active,
                  int c0 = 0, int nc = -1)
{
    if (em_after_step_queue(ctx, st, tole, iteration, d_count, count_all, 0, c0, nc)) return 1;
    return em_after_step_wait(ctx, st, 0, n_active);
}

// Stop rule of emMAF.py:21-25 with rmse1d's float divide / double sqrt (emMAF_cy.pyx:32-33).
inline bool em_converged(double ssq, double count, double tole)
{
    float res = (float)ssq;
    res = res / (float)count;
    return std::sqrt((double)res) < tole;
}

// Per-population EM on the resident G: FT [K][M] (device, population-major) <- converged, UNclipped f.
// Up to kEmChunk iterations per read of G (em_pop_multi_kernel); WGS_EM_STEP=1 selects the
// one-iteration-per-launch kernel (same bits, 14 reads of G instead of 3).
int run_em_pop(wgs_ctx* ctx, int iter, double tole, float* FT, std::vector<int>& iters_out, int k0 = 0, int kn = -1)
{
    // populations [k0, k0 + kn) only (kn < 0: all): FT then points at THEIR block of the population-major state
    const long M = ctx->M();
    const int K = kn < 0 ? std::max(ctx->K, 1) : kn;
    const PopDesc* hpops = ctx->pops.data() + k0;
    const PopDesc* dpops = ctx->d_pops + k0;
    if (K > kMaxKq * 32) return fail(ctx, "more than %d populations not supported", kMaxKq * 32);
    int nmax = 1;
    for (int k = 0; k < K; ++k) nmax = std::max(nmax, hpops[k].n);
    int row16 = (nmax + 3) / 4 * 2;                             // widest slab in 16-byte units ...
    while (row16 % 8 != kEmT % 8) ++row16;                      // ... padded so that kEmT threads x 2 rows hit 8 distinct bank groups
    int R = 128;
    while (R > 8 && 2 * (size_t)R * row16 * 16 > 110 * 1024) R /= 2;
    size_t smem = 2 * (size_t)R * row16 * 16;
    if (smem > 220 * 1024) return fail(ctx, "population of %d individuals exceeds the EM shared-memory tile", nmax);
    const bool multi = !opt(ctx, "em_step");
    // register-tile variant (em_pop_multi2): TPR threads per row, up to 4 quads of individuals each (n <= 512)
    int tpr = 4, qpt = 4;
    {
        const int nqmax = (nmax + 3) / 4;
        if (nqmax <= 4) qpt = 1; else if (nqmax <= 8) qpt = 2; else qpt = 4;
        while (tpr < 32 && tpr * qpt < nqmax) tpr *= 2;
    }
    const bool packed = multi && !opt(ctx, "em_multi1") && tpr * qpt * 4 >= nmax;
    int raw16 = (nmax + 3) / 4 * 2;                             // pairs per slab row, whole quads ...
    if (raw16 % 4 == 0) raw16 += 2;                             // ... rows two apart land on different bank groups
    const int R2 = 256 / tpr;
    const size_t smem2 = 2 * (size_t)R2 * raw16 * 16;
    // always opt in: static + dynamic shared memory together may cross the 48 KB default limit
    int occ = 1;
#define EM2_CASE(TPRV, QPTV)                                                                                             \
    do {                                                                                                                 \
        CU(cudaFuncSetAttribute(em_pop_multi2_kernel<TPRV, QPTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024))); \
        CU(cudaFuncSetAttribute(em_pop_multi2_kernel<TPRV, QPTV>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, em_pop_multi2_kernel<TPRV, QPTV>, 256, smem));             \
    } while (0)
    if (packed) {
        R = R2; smem = smem2;
        if (tpr == 4 && qpt == 1) EM2_CASE(4, 1); else if (tpr == 4 && qpt == 2) EM2_CASE(4, 2); else if (tpr == 4) EM2_CASE(4, 4);
        else if (tpr == 8) EM2_CASE(8, 4); else if (tpr == 16) EM2_CASE(16, 4); else EM2_CASE(32, 4);
#undef EM2_CASE
    } else if (multi) {
        CU(cudaFuncSetAttribute(em_pop_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, em_pop_multi_kernel, R * kEmT, smem));
    } else {
        CU(cudaFuncSetAttribute(em_pop_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, em_pop_step_kernel, R * kEmT, smem));
    }
    long ntiles = (M + R - 1) / R;
    // one full wave of persistent blocks: gx * K <= resident slots
    int gx = (int)std::max<long>(1, std::min<long>(ntiles, ((long)ctx->num_sm * std::max(occ, 1)) / K));
    LAUNCH("fill", fill_kernel, grid_for(M * K, 256, ctx->num_sm * 8), 256, 0, ctx->stream, FT, M * K, 0.25f);
    iters_out.assign(K, 0);
    if (!multi) {
        EmState st;
        if (em_state_init(ctx, st, K, K, gx, std::vector<int>(K, 1))) return 1;
        int n_active = K;
        for (int it = 1; it <= iter && n_active > 0; ++it) {
            LAUNCH("em_pop", em_pop_step_kernel, dim3(gx, K), R * kEmT, smem, ctx->stream, ctx->G[0], ctx->ldg, M, dpops, K, FT,
                   st.active.as<int>(), row16, st.partials.as<double>());
            {   // 8 B per (site, individual of an active population) + f read/write
                double inds = 0, act = 0;
                for (int k = 0; k < K; ++k) if (st.h_active[k]) { inds += hpops[k].n; act += 1; }
                add_work(ctx, "em_pop", (double)M * inds * 8.0 + (double)M * act * 8.0, (double)M * inds);
            }
            if (em_after_step(ctx, st, tole, it, nullptr, (double)ctx->Mtot(), &n_active)) return 1;
        }
        CU(cudaMemcpy(iters_out.data(), st.iters.p, K * sizeof(int), cudaMemcpyDeviceToHost));
        CU(cudaGetLastError());
        return 0;
    }

    const int np = K * kEmChunk;
    // the register-tile kernel also records the state after EVERY iteration of a pass (8 x 4 bytes per site and
    // population - nothing next to the pass itself), so a stop inside a chunk needs no replay pass
    DevBuf FT1, partials, ssq, dcur, diters, hist;
    if (packed && buf_alloc(ctx, hist, (size_t)kEmChunk * std::max<long>(M, 1) * K * sizeof(float))) return 1;
    if (buf_alloc(ctx, FT1, (size_t)std::max<long>(M, 1) * K * sizeof(float)) || buf_alloc(ctx, partials, (size_t)gx * np * sizeof(double)) ||
        buf_alloc(ctx, ssq, (size_t)np * sizeof(double)) || buf_alloc(ctx, dcur, K * sizeof(int)) || buf_alloc(ctx, diters, K * sizeof(int))) return 1;
    std::vector<int> cur(K, 0), done(K, 0), replay(K, 0), run(K, 0);
    std::vector<char> fin(K, 0);
    std::vector<double> h(np);
    // exact stop rule (see EmState): the register-tile kernel's per-iteration history holds what the sequential sum needs
    const bool exact = packed && opt(ctx, "rmse_exact", 1) != 0;
    const double band_ov = band_override_of(ctx);
    DevBuf dser, dcarry;
    if (exact && (buf_alloc(ctx, dser, sizeof(float)) || buf_alloc(ctx, dcarry, sizeof(float)))) return 1;
    for (;;) {
        bool any = false;
        for (int k = 0; k < K; ++k) {
            run[k] = 0;
            if (fin[k]) continue;
            run[k] = replay[k] > 0 ? replay[k] : std::min(kEmChunk, iter - done[k]);
            if (run[k] <= 0) { fin[k] = 1; run[k] = 0; continue; }           // iteration limit reached without convergence
            any = true;
        }
        if (!any) break;
        CU(cudaMemcpyAsync(dcur.p, cur.data(), K * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(diters.p, run.data(), K * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
#define EM2_LAUNCH(TPRV, QPTV)                                                                                          \
    LAUNCH("em_pop", (em_pop_multi2_kernel<TPRV, QPTV>), dim3(gx, K), 256, smem, ctx->stream, ctx->G[0], ctx->ldg, M, dpops, K, FT, \
           FT1.as<float>(), dcur.as<int>(), diters.as<int>(), raw16, partials.as<double>(), hist.as<float>())
        if (packed) {
            if (tpr == 4 && qpt == 1) EM2_LAUNCH(4, 1); else if (tpr == 4 && qpt == 2) EM2_LAUNCH(4, 2); else if (tpr == 4) EM2_LAUNCH(4, 4);
            else if (tpr == 8) EM2_LAUNCH(8, 4); else if (tpr == 16) EM2_LAUNCH(16, 4); else EM2_LAUNCH(32, 4);
        }
#undef EM2_LAUNCH
        else
            LAUNCH("em_pop", em_pop_multi_kernel, dim3(gx, K), R * kEmT, smem, ctx->stream, ctx->G[0], ctx->ldg, M, dpops, K, FT,
                   FT1.as<float>(), dcur.as<int>(), diters.as<int>(), row16, partials.as<double>());
        {   // 8 B per (site, individual of a running population) + f read/write, once per pass; one unit per (site, individual, iteration)
            double inds = 0, act = 0, units = 0;
            for (int k = 0; k < K; ++k) if (run[k] > 0) { inds += hpops[k].n; act += 1; units += (double)hpops[k].n * run[k]; }
            add_work(ctx, "em_pop", (double)M * inds * 8.0 + (double)M * act * 8.0, (double)M * units);
        }
        LAUNCH("em_ssq_reduce", em_ssq_reduce_kernel, (np + 7) / 8, 256, 0, ctx->stream, partials.as<double>(), gx, np, np, ssq.as<double>());
        if (nccl_sums(ctx) && dev_all_sum(ctx, ssq.p, np, WGS_F64)) return 1;
        CU(cudaMemcpyAsync(h.data(), ssq.p, np * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->fn && !nccl_sums(ctx)) ctx->fn(h.data(), np, WGS_F64, ctx->user);
        for (int k = 0; k < K; ++k) {
            if (run[k] <= 0) continue;
            if (replay[k] > 0) {                                            // exactly t* iterations from the chunk's start state
                cur[k] ^= 1; done[k] += replay[k]; replay[k] = 0; fin[k] = 1; iters_out[k] = done[k];
                continue;
            }
            int tstar = 0;
            for (int u = 0; u < run[k] && !tstar; ++u) {
                const double ssq_u = h[(size_t)k * kEmChunk + u];
                bool conv = em_converged(ssq_u, (double)ctx->Mtot(), tole);
                if (exact && em_uncertain(ssq_u, (double)ctx->Mtot(), tole, band_ov)) {
                    // inside the band where the exact sum cannot decide: the reference's sequential float32 sum of this
                    // iteration's squared changes, from the state history (ranks chained in site order)
                    const float* curp = hist.as<float>() + ((size_t)u * K + k) * M;
                    const float* prevp = u == 0 ? (cur[k] ? FT1.as<float>() : FT) + (size_t)k * M : hist.as<float>() + ((size_t)(u - 1) * K + k) * M;
                    if (chain_floats(ctx, dser.as<float>(), dcarry.as<float>(), 1, [&](const float* cin) {
                            LAUNCH("em_resolve", seqsum_pair_kernel, 1, kSeqWarps * 32, 0, ctx->stream, curp, prevp, M, cin, dser.as<float>());
                            return 0;
                        })) return 1;
                    float ser = 0.f;
                    CU(cudaMemcpyAsync(&ser, dser.p, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
                    CU(cudaStreamSynchronize(ctx->stream));
                    float res = ser / (float)ctx->Mtot();
                    conv = std::sqrt((double)res) < tole;
                }
                if (conv) tstar = u + 1;
            }
            if (tstar == 0 || tstar == run[k]) {                            // the state written by this pass is the one to keep
                cur[k] ^= 1; done[k] += run[k];
                if (tstar) { fin[k] = 1; iters_out[k] = done[k]; }
            } else if (packed) {                                            // stop inside the chunk: that iteration's state is in the history
                float* dst = (cur[k] ? FT : FT1.as<float>()) + (size_t)k * M;
                CU(cudaMemcpyAsync(dst, hist.as<float>() + ((size_t)(tstar - 1) * K + k) * M, (size_t)M * sizeof(float),
                                   cudaMemcpyDeviceToDevice, ctx->stream));
                cur[k] ^= 1; done[k] += tstar; fin[k] = 1; iters_out[k] = done[k];
            } else {
                replay[k] = tstar;                                          // stop inside the chunk: replay from FT[cur]
            }
        }
    }
    for (int k = 0; k < K; ++k)
        if (cur[k]) CU(cudaMemcpyAsync(FT + (size_t)k * M, FT1.as<float>() + (size_t)k * M, (size_t)M * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return 0;
}

struct LooLaunch { int block, rows_per_pass, passes, grid; bool big; size_t smem; bool packed; int nc; int stages; };

// Launch geometry of the leave-one-out EM step kernels (four problems per thread).  packed = the
// pre-packed pair-polynomial kernel (loo_em_step5), else the in-kernel packing quad kernel (loo_em_step4).
// The packed kernel runs one block of 512 threads per SM (116 registers, no spills, ring of up to 190 KB) unless two
// blocks of 256 idle fewer threads; populations of > 1024 individuals always take one block of 512.
int loo_cfg(wgs_ctx* ctx, int n, bool packed, LooLaunch* out)
{
    LooLaunch best{0, 0, 1, 0, false, 0, packed, 0, 3};
    const int nq = (n + 3) / 4;                                   // threads per site row (four problems each)
    best.nc = (nq + 1) / 2;
    double best_u = -1;
    // whole multiples of 4 warps only: 7 warps per block leave one scheduler of the SM with less work
    // than the others (measured: 12 % slower than 8 warps, scripts/microbench/loo_quad_rate.cu)
    const int opt_bd = opt(ctx, "loo_block", 0);
    for (int bd = 128; bd <= 512; bd += 128) {
        if (opt_bd && opt_bd != bd) continue;
        int rpp = bd / nq;
        if (rpp < 1) continue;
        double u = (double)(rpp * nq) / bd;
        if (bd > 256 && best_u >= 0.88) break;                    // blocks above 256 threads (one per SM) only when the smaller ones idle > 12 % of their threads
        if (u > best_u + 1e-9) { best_u = u; best.block = bd; best.rows_per_pass = rpp; }
    }
    if (best.block == 0) return fail(ctx, "population of %d individuals exceeds the LOO-EM block limit (2048)", n);
    // one block of 16 warps (one ring of up to 190 KB) instead of two of 8 unless that idles more threads: measured
    // 0.517 against 0.554 ms per launch at n = 50 (507 of 512 against 247 of 256 working threads) and 0.383 against
    // 0.395 ms at n = 100 (the same 500 of 512 / 250 of 256)
    if (packed && !opt_bd && best.block <= 256 && 512 / nq >= 1) {
        const double u512 = (double)((512 / nq) * nq) / 512;
        if (u512 + 1e-9 >= best_u + opt(ctx, "loo_big_margin_pm", 0) * 1e-3) { best_u = u512; best.block = 512; best.rows_per_pass = 512 / nq; }
    }
    best.big = best.block > 256;
    // per tile row - quad kernel: two packed buffers (odd 16-byte stride) + the raw TMA landing row;
    //              - packed kernel: double-buffered packed cells (odd 16-byte stride) + double-buffered raw row
    if (packed) {
        // ring of row groups: as many stages as fit in ~100 KB (two resident blocks per SM), at least 3
        const size_t group_bytes = (size_t)best.rows_per_pass * loo5_row_units(n) * 16;
        const size_t ring_budget = (best.big ? 190 : 100) * 1024;   // one / two resident blocks per SM
        int stages = (int)std::min<size_t>(kLoo5MaxStages, ring_budget / std::max<size_t>(group_bytes, 1));
        if (int o = opt(ctx, "loo_stages", 0)) stages = o;
        best.stages = std::max(3, std::min(stages, kLoo5MaxStages));
        best.passes = 1;
        best.smem = best.stages * group_bytes + (size_t)best.block * sizeof(float4);
    } else {
        // per tile row: two packed buffers (odd 16-byte stride) + the raw TMA landing row
        const size_t row_bytes = 2 * (size_t)((3 * nq) | 1) * 16 + (size_t)nq * 32;
        int passes = kLoo4MaxPasses;
        while (passes > 1 && (size_t)best.rows_per_pass * passes * row_bytes > 68 * 1024) --passes;
        if (int o = opt(ctx, "loo_passes", 0)) passes = std::max(1, std::min(passes, o));
        best.passes = passes;
        best.smem = (size_t)best.rows_per_pass * passes * row_bytes + (size_t)best.block * sizeof(float4);
    }
    if (best.smem > 200 * 1024) return fail(ctx, "population of %d individuals exceeds the LOO-EM shared-memory tile", n);
    int occ = 1;
#define LOO_PREP(KERN)                                                                                               \
    do {                                                                                                             \
        CU(cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));                     \
        CU(cudaFuncSetAttribute(KERN, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, KERN, best.block, best.smem));                        \
    } while (0)
    if (packed) {
        if (best.big) LOO_PREP((loo_em_step5_kernel<512, 1>));
        else LOO_PREP((loo_em_step5_kernel<256, 2>));
    }
    else        { if (best.big) LOO_PREP((loo_em_step4_kernel<512, 1>)); else LOO_PREP((loo_em_step4_kernel<256, 3>)); }
#undef LOO_PREP
    best.grid = ctx->num_sm * std::max(occ, 1);
    *out = best;
    return 0;
}

// Leave-one-out EM for every individual on the resident G.  F [M][ldf] (device): columns
// [0,ldg) <- converged UNclipped LOO estimates.  mask/d_count optional (reference z-score).
// pop_ready (pipelined mode, wgs_ref_af_loo after wgs_upload_gl_async): one event per population, recorded when
// that population's slab of G has arrived.  The populations then run one after the other - pre-pack and all
// iterations of population k while the slabs of k+1.. are still in flight - instead of all together per
// iteration.  Every problem sees exactly the same sequence of updates and decisions either way.
int run_em_loo(wgs_ctx* ctx, int iter, double tole, float* F, int ldf, const unsigned char* mask, const double* d_count,
               std::vector<int>& iters_cols, const std::vector<unsigned char>* sel = nullptr,
               const std::vector<cudaEvent_t>* pop_ready = nullptr, const std::function<int(int)>* on_pop = nullptr,
               bool force_chain = false)
{
    const long M = ctx->M();
    const int ldg = ctx->ldg, K = ctx->K;
    std::vector<int> active0(ldg, 0);
    for (int c = 0; c < ldg; ++c)
        if (ctx->ind_of_col[c] >= 0 && ctx->pops[ctx->pop_of_col[c]].n > 1 && (!sel || (*sel)[c])) active0[c] = 1;
    std::vector<LooLaunch> cfgs(K);
    // packed rows of population k: through a shared-memory tile of whole rows when at least one row fits a block
    auto launch_prepack = [&](int k, ulonglong2* PKk) -> int {
        const PopDesc pd = ctx->pops[k];
        const int nc = cfgs[k].nc;
        const size_t sm = (size_t)(256 / std::max(nc, 1)) * loo5_row_units(pd.n) * sizeof(ulonglong2);
        if (nc <= 256 && sm <= 64 * 1024 && !opt(ctx, "prepack_v1")) {
            CU(cudaFuncSetAttribute(loo_prepack2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            const long ntl = (M + 256 / nc - 1) / (256 / nc);
            LAUNCH("loo_pack", loo_prepack2_kernel, (int)std::max<long>(1, std::min<long>(ntl, (long)ctx->num_sm * 5)), 256, sm, ctx->stream,
                   ctx->G[0], ldg, M, pd.col0, pd.n, nc, PKk);
        } else {
            LAUNCH("loo_pack", loo_prepack_kernel, grid_for(M * nc, 256, ctx->num_sm * 16), 256, 0, ctx->stream, ctx->G[0], ldg, M,
                   pd.col0, pd.n, nc, PKk);
        }
        add_work(ctx, "loo_pack", (double)M * pd.n * 8.0 + (double)M * loo5_row_units(pd.n) * 16.0, (double)M * pd.n);
        return 0;
    };
    int nblocks = 1;
    for (int k = 0; k < K; ++k) {
        if (ctx->pops[k].n <= 1) continue;
        nblocks = std::max(nblocks, 1);
    }
    // pre-packed pair polynomials (loo_em_step5): 80 bytes per 8 individuals and site, written once per call;
    // when the pool cannot provide them (or WGS_LOO_V4 is set) the in-kernel packing quad kernel runs instead
    // by_pop: ONE packed-row buffer, the populations run one after the other (pre-pack k, all iterations of k, then
    // k+1) - what the pipelined mode does anyway, and what a matrix too large for all packed rows at once (a 2.5 M
    // x 2,000 shard would need 92 GB of them) falls back to before giving up the packed kernel.  Every problem sees
    // the same updates and decisions either way.
    Trace tr(ctx, "em_loo");
    std::vector<DevBuf> pk(K);
    const int loo_dbg = opt(ctx, "loo_dbg", 0);
    bool packed = !opt(ctx, "loo_v4");
    bool by_pop = pop_ready != nullptr;
    if (packed) {
        size_t total_pk = 0, max_pk = 0, cached = 0;
        for (int k = 0; k < K; ++k) {
            if (ctx->pops[k].n <= 1) continue;
            const size_t b = (size_t)std::max<long>(M, 1) * loo5_row_units(ctx->pops[k].n) * sizeof(ulonglong2);
            total_pk += b; max_pk = std::max(max_pk, b);
        }
        for (auto& kv : ctx->pool_free) cached += kv.first;
        const size_t freeb = device_free_bytes(ctx);
        tr.lap("meminfo");
        const size_t d2b = opt(ctx, "rmse_exact", 1) ? (size_t)std::max<long>(M, 1) * ldg * sizeof(float) : 0;
        const size_t avail = freeb + cached, slack = (size_t)2 << 30;
        const int o = opt(ctx, "loo_by_pop", 0);
        if (!pop_ready && o >= 0 && (o > 0 || total_pk + d2b + slack > avail)) by_pop = true;
        if (by_pop && !pop_ready) {
            pk[0].owner = ctx;
            pk[0].p = pool_take(ctx, std::max<size_t>(max_pk, 16));
            if (!pk[0].p) packed = false;
        } else if (!by_pop) {
            for (int k = 0; k < K && packed; ++k) {
                if (ctx->pops[k].n <= 1) continue;
                pk[k].owner = ctx;
                pk[k].p = pool_take(ctx, (size_t)std::max<long>(M, 1) * loo5_row_units(ctx->pops[k].n) * sizeof(ulonglong2));
                if (!pk[k].p) packed = false;
            }
        } else {                                                  // pipelined upload: all buffers (the populations overlap the transfer)
            for (int k = 0; k < K && packed; ++k) {
                if (ctx->pops[k].n <= 1) continue;
                pk[k].owner = ctx;
                pk[k].p = pool_take(ctx, (size_t)std::max<long>(M, 1) * loo5_row_units(ctx->pops[k].n) * sizeof(ulonglong2));
                if (!pk[k].p) packed = false;
            }
        }
        if (!packed) for (int k = 0; k < K; ++k) buf_release(pk[k]);
    }
    // iteration 1 (loo_first, 256-thread blocks) has its own grid; rows of the per-block squared-change table that only
    // it wrote are cleared before iteration 2's reduction (launch_iter)
    const int first_grid = ctx->num_sm * std::max(1, opt(ctx, "loo_first_bpsm", 2));
    tr.lap("packed alloc");
    const bool shared_pk = packed && by_pop && !pop_ready;       // every population uses pk[0]
    auto pk_of = [&](int k) { return (shared_pk ? pk[0] : pk[k]).as<ulonglong2>(); };
    for (int k = 0; k < K; ++k) {
        if (ctx->pops[k].n <= 1) continue;
        if (loo_cfg(ctx, ctx->pops[k].n, packed, &cfgs[k])) return 1;
        nblocks = std::max(nblocks, std::max(cfgs[k].grid, first_grid));
        if (packed && !by_pop) {
            if (launch_prepack(k, pk_of(k))) return 1;
        }
    }
    tr.lap("cfg+prepack");
    EmState st;
    if (em_state_init(ctx, st, ldg, ldg, nblocks, active0, opt(ctx, "rmse_exact", 1) != 0, mask != nullptr)) return 1;
    if (force_chain) st.chain_always = true;
    tr.lap("state init");
    float* const D2 = st.d2.as<float>();                         // null when the exact stop rule is switched off
    // When every real column is an active problem of a population that loo_first serves, iteration 1 writes all of
    // them without reading the start state: only the padding columns need a value (the step kernels load whole quads).
    bool pads_only = !mask && !sel && iter >= 1 && !opt(ctx, "loo_nofirst") && !opt(ctx, "loo_fullfill");
    for (int k = 0; k < K; ++k)
        if (ctx->pops[k].n <= 1 || (ctx->pops[k].n + 1) / 2 > 32 * kFisherQ) pads_only = false;
    if (pads_only) {
        std::vector<int> pads;
        for (int c = 0; c < ldg; ++c) if (ctx->ind_of_col[c] < 0) pads.push_back(c);
        if (!pads.empty() && M > 0) {
            DevBuf dp;
            if (buf_alloc(ctx, dp, pads.size() * sizeof(int))) return 1;
            CU(cudaMemcpyAsync(dp.p, pads.data(), pads.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            LAUNCH("fill", fill_cols_kernel, grid_for(M * (long)pads.size(), 256, ctx->num_sm * 8), 256, 0, ctx->stream, F, ldf, M,
                   dp.as<int>(), (int)pads.size(), 0.25f);
            CU(cudaStreamSynchronize(ctx->stream));
        }
    } else {   // f = 0.25 everywhere; NaN for the (degenerate) single-member populations, like 0/0 in the reference
        std::vector<float> row(ldg, 0.25f);
        for (int c = 0; c < ldg; ++c)
            if (ctx->ind_of_col[c] >= 0 && ctx->pops[ctx->pop_of_col[c]].n <= 1) row[c] = std::numeric_limits<float>::quiet_NaN();
        DevBuf drow;
        if (buf_alloc(ctx, drow, ldg * sizeof(float))) return 1;
        CU(cudaMemcpyAsync(drow.p, row.data(), ldg * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH("fill", bcast_row_kernel, grid_for(M * ldg, 256, ctx->num_sm * 8), 256, 0, ctx->stream, F, ldf, ldg, M, drow.as<float>());
        CU(cudaStreamSynchronize(ctx->stream));
    }
    tr.lap("start fill");
    auto launch_step = [&](int k) -> int {
        PopDesc pd = ctx->pops[k];
        const LooLaunch& lc = cfgs[k];
        int TS = lc.rows_per_pass * lc.passes;
        long ntiles = (M + TS - 1) / TS;
        if (lc.packed) {
            if (lc.big)
                LAUNCH("loo_em", (loo_em_step5_kernel<512, 1>), lc.grid, lc.block, lc.smem, ctx->stream, pk_of(k), ldg, M,
                       pd.col0, pd.n, lc.rows_per_pass, F, ldf, st.active.as<int>(), mask, st.partials.as<double>(), ntiles, lc.stages, loo_dbg, D2);
            else
                LAUNCH("loo_em", (loo_em_step5_kernel<256, 2>), lc.grid, lc.block, lc.smem, ctx->stream, pk_of(k), ldg, M,
                       pd.col0, pd.n, lc.rows_per_pass, F, ldf, st.active.as<int>(), mask, st.partials.as<double>(), ntiles, lc.stages, loo_dbg, D2);
        } else {
            if (lc.big)
                LAUNCH("loo_em", (loo_em_step4_kernel<512, 1>), lc.grid, lc.block, lc.smem, ctx->stream, ctx->G[0], ldg, M, pd.col0, pd.n,
                       lc.rows_per_pass, lc.passes, F, ldf, st.active.as<int>(), mask, st.partials.as<double>(), ntiles, D2);
            else
                LAUNCH("loo_em", (loo_em_step4_kernel<256, 3>), lc.grid, lc.block, lc.smem, ctx->stream, ctx->G[0], ldg, M, pd.col0, pd.n,
                       lc.rows_per_pass, lc.passes, F, ldf, st.active.as<int>(), mask, st.partials.as<double>(), ntiles, D2);
        }
        return 0;
    };
    // Iteration 1 starts every problem of a population from the same f: n evaluations per site serve all n problems
    // (loo_first_kernel) instead of n^2.  Populations above 512 individuals take the general step kernel.
    const bool use_first = !opt(ctx, "loo_nofirst");
    auto first_tpr = [&](int k) {
        int tpr = 2;
        const int cpr = (ctx->pops[k].n + 1) / 2;
        while (tpr < 32 && tpr * kFisherQ < cpr) tpr *= 2;
        return tpr * kFisherQ >= cpr ? tpr : 0;
    };
    auto launch_iter = [&](int k, int it) -> int {
        const int tpr = (it == 1 && use_first) ? first_tpr(k) : 0;
        PopDesc pd = ctx->pops[k];
        if (!tpr) {
            if (it == 2 && use_first && first_tpr(k) && first_grid > cfgs[k].grid)
                CU(cudaMemset2DAsync(st.partials.as<double>() + (size_t)cfgs[k].grid * ldg + pd.col0, (size_t)ldg * sizeof(double), 0,
                                     (size_t)pd.n * sizeof(double), first_grid - cfgs[k].grid, ctx->stream));
            return launch_step(k);
        }
        const size_t sm = (size_t)256 * 2 * kFisherQ * sizeof(float);
#define LOO_FIRST(T)                                                                                             \
    LAUNCH("loo_first", loo_first_kernel<T>, first_grid, 256, sm, ctx->stream, ctx->G[0], ldg, M, pd.col0, pd.n, F, ldf, \
           st.active.as<int>(), mask, st.partials.as<double>(), D2)
        if (tpr == 2) LOO_FIRST(2); else if (tpr == 4) LOO_FIRST(4); else if (tpr == 8) LOO_FIRST(8);
        else if (tpr == 16) LOO_FIRST(16); else LOO_FIRST(32);
#undef LOO_FIRST
        add_work(ctx, "loo_first", (double)M * pd.n * 8.0 + (double)M * pd.n * (D2 ? 8.0 : 4.0), (double)M * pd.n);
        return 0;
    };
    // algorithmic work of one iteration of population k, from the flags that were in force when it ran: the
    // population's packed rows once + read/write of every active problem's f (+ its squared change, written for the
    // exact stop rule); n evaluations per active (site, problem)
    auto account = [&](int k, int it) {
        if (it == 1 && use_first && first_tpr(k)) return;           // accounted as "loo_first" at launch
        PopDesc pd = ctx->pops[k];
        double act = 0;
        for (int j = 0; j < pd.n; ++j) act += st.h_active[pd.col0 + j] ? 1 : 0;
        if (act > 0)
            add_work(ctx, "loo_em", (double)M * (cfgs[k].packed ? loo5_row_units(pd.n) * 16.0 : pd.n * 8.0) + (double)M * act * (D2 ? 12.0 : 8.0), (double)M * act * pd.n);
    };
    auto pop_active = [&](int k) {
        bool any = false;
        for (int j = 0; j < ctx->pops[k].n; ++j) any = any || st.h_active[ctx->pops[k].col0 + j];
        return any;
    };
    // Look-ahead: iteration t+1 is queued before the host has read the decision of iteration t (the device skips
    // whatever that decision froze), so the GPU never idles on the round trip.  Not under site sharding, where the
    // squared changes pass through the host for the all-reduce anyway.
    const bool ahead = (ctx->fn == nullptr || ctx->nccl_comm != nullptr) && !opt(ctx, "em_no_lookahead");
    const double count_all = (double)ctx->Mtot();
    if (by_pop) {
        for (int ko = 0; ko < K; ++ko) {
            const int k = (pop_ready && (int)ctx->upload_order.size() == K) ? ctx->upload_order[ko] : ko;   // the order the slabs arrive in
            PopDesc pd = ctx->pops[k];
            if (pop_ready) CU(cudaStreamWaitEvent(ctx->stream, (*pop_ready)[k], 0));   // device-side: the host keeps queueing
            if (on_pop && (*on_pop)(k)) return 1;                           // the fused operator's full-data EM of this population
            if (pd.n <= 1) continue;
            if (packed) {
                if (launch_prepack(k, pk_of(k))) return 1;
            }
            if (!pop_active(k) || iter < 1) continue;
            if (launch_iter(k, 1) || em_after_step_queue(ctx, st, tole, 1, d_count, count_all, 1, pd.col0, pd.n)) return 1;
            for (int it = 1; it <= iter; ++it) {
                if (ahead && it < iter)
                    if (launch_iter(k, it + 1) || em_after_step_queue(ctx, st, tole, it + 1, d_count, count_all, (it + 1) & 1, pd.col0, pd.n)) return 1;
                int n_act = 0;
                account(k, it);
                if (em_after_step_wait(ctx, st, it & 1, &n_act)) return 1;
                if (n_act == 0) break;
                if (!ahead && it < iter)
                    if (launch_iter(k, it + 1) || em_after_step_queue(ctx, st, tole, it + 1, d_count, count_all, (it + 1) & 1, pd.col0, pd.n)) return 1;
            }
        }
    } else {
        auto queue_round = [&](int it) -> int {
            for (int k = 0; k < K; ++k) {
                if (ctx->pops[k].n <= 1 || !pop_active(k)) continue;
                if (launch_iter(k, it)) return 1;
            }
            return em_after_step_queue(ctx, st, tole, it, d_count, count_all, it & 1);
        };
        int n_active = 0;
        for (int a : active0) n_active += a;
        if (n_active > 0 && iter >= 1) {
            if (queue_round(1)) return 1;
            for (int it = 1; it <= iter; ++it) {
                if (ahead && it < iter && queue_round(it + 1)) return 1;
                for (int k = 0; k < K; ++k) if (ctx->pops[k].n > 1) account(k, it);
                if (em_after_step_wait(ctx, st, it & 1, &n_active)) return 1;
                if (n_active == 0) break;
                if (!ahead && it < iter && queue_round(it + 1)) return 1;
            }
        }
    }
    tr.lap("iterations");
    CU(cudaStreamSynchronize(ctx->stream));                      // a speculative round may still be running
    tr.lap("drain");
    if (st.missed && !force_chain) {
        // A stop check landed inside the undecidable band while the rank chain was not queued (a population that
        // converges by more than 30x within two iterations): run again with the chain on from the first iteration.
        // The EM is deterministic, so nothing but time is lost.
        for (auto& b : pk) buf_release(b);
        buf_release(st.d2);
        return run_em_loo(ctx, iter, tole, F, ldf, mask, d_count, iters_cols, sel, pop_ready, on_pop, true);
    }
    iters_cols.resize(ldg);
    CU(cudaMemcpy(iters_cols.data(), st.iters.p, ldg * sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaGetLastError());
    return 0;
}

// clip bounds per sorted column for n_eff = n_pop - minus individuals (float like numpy's compare/assign)
void clip_bounds(wgs_ctx* ctx, int minus, std::vector<float>& lo, std::vector<float>& hi)
{
    lo.assign(ctx->ldg, 0.f); hi.assign(ctx->ldg, 1.f);
    for (int c = 0; c < ctx->ldg; ++c) {
        int n = ctx->pops[ctx->pop_of_col[c]].n - minus;
        double l = 1.0 / (2.0 * (n + 1));
        lo[c] = (float)l; hi[c] = (float)(1.0 - l);
    }
}

}  // namespace

// =============================================================================================
extern "C" {

int32_t wgs_abi_version(void) { return WGS_ABI_VERSION; }

int32_t wgs_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int32_t wgs_device_pci_bus_id(int32_t device, char* out, int32_t len)
{
    if (!out || len < 16) return 1;
    out[0] = 0;
    return cudaDeviceGetPCIBusId(out, len, device) == cudaSuccess ? 0 : 1;
}

const char* wgs_last_error(const wgs_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int32_t wgs_create(int32_t device, wgs_ctx** out)
{
    wgs_ctx* ctx = nullptr;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(nullptr, "no CUDA device available (%s) - wgsassign_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, "device %d out of range (%d devices)", device, n);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major < 10) return fail(nullptr, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    ctx = new wgs_ctx();
    ctx->device = device;
    ctx->num_sm = prop.multiProcessorCount;
    cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
    if (em_pin_reserve(ctx, 1)) { g_create_error = ctx->err; wgs_destroy(ctx); return 1; }
    if (getenv("WGS_DEBUG"))                                     // debugging only: WGS_<OPTION>=<int> seeds the option table once
        for (const char* nm : kOptionNames) {
            std::string env = "WGS_";
            for (const char* p = nm; *p; ++p) env += (char)toupper(*p);
            if (const char* v = getenv(env.c_str())) ctx->opts[nm] = atoi(v);
        }
    *out = ctx;
    return 0;
}

void wgs_destroy(wgs_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    fold_timing(ctx);
    drop_data(ctx);
    dev_free(ctx, ctx->d_ind_of_col); dev_free(ctx, ctx->d_col_of_ind); dev_free(ctx, ctx->d_pop_of_col); dev_free(ctx, ctx->d_pops);
    pool_trim(ctx);
    if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->nccl_comm);
    for (cudaEvent_t e : ctx->ev_pop) cudaEventDestroy(e);
    for (int b = 0; b < 2; ++b) { if (ctx->em_pin[b]) cudaFreeHost(ctx->em_pin[b]); if (ctx->em_ev[b]) cudaEventDestroy(ctx->em_ev[b]); }
    cudaStreamDestroy(ctx->stream); cudaStreamDestroy(ctx->stream2);
    delete ctx;
}

int32_t wgs_set_option(wgs_ctx* ctx, const char* name, int32_t value)
{
    for (const char* nm : kOptionNames)
        if (!strcmp(nm, name)) { ctx->opts[name] = value; return 0; }
    return fail(ctx, "unknown option '%s'", name);
}

int32_t wgs_get_option(const wgs_ctx* ctx, const char* name, int32_t* value)
{
    for (const char* nm : kOptionNames)
        if (!strcmp(nm, name)) { *value = opt(ctx, name, 0); return 0; }
    return 1;
}

void* wgs_host_alloc(int64_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void wgs_host_free(void* p) { if (p) cudaFreeHost(p); }

int32_t wgs_set_pops(wgs_ctx* ctx, const int32_t* pop_of_ind, int32_t N, int32_t K)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    if (N <= 0) return fail(ctx, "N must be positive");
    drop_data(ctx);
    ctx->pops_set = K > 0;
    return build_structure(ctx, pop_of_ind, N, K);
}

int32_t wgs_nccl_unique_id(void* id_out)
{
    if (!nccl_load()) return fail(nullptr, "libnccl.so.2 could not be loaded");
    int rc = g_nccl.GetUniqueId(id_out);
    if (rc) return fail(nullptr, "ncclGetUniqueId failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    return 0;
}

int32_t wgs_nccl_init(wgs_ctx* ctx, const void* id, int32_t rank, int32_t world)
{
    cudaSetDevice(ctx->device);                                  // no upload fence: the communicator does not touch the matrix
    if (!nccl_load()) return fail(ctx, "libnccl.so.2 could not be loaded");
    if (ctx->nccl_comm) { g_nccl.CommDestroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    if (world <= 1) return 0;
    NcclId128 uid;
    memcpy(&uid, id, sizeof uid);
    void* comm = nullptr;
    int rc = g_nccl.CommInitRank(&comm, world, uid, rank);
    if (rc) return fail(ctx, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    ctx->nccl_comm = comm; ctx->nccl_rank = rank; ctx->nccl_world = world;
    ctx->rank = rank; ctx->world = world;
    return 0;
}

int32_t wgs_set_rank(wgs_ctx* ctx, int32_t rank, int32_t world)
{
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, "rank %d outside [0,%d)", rank, world);
    ctx->rank = rank; ctx->world = world;
    return 0;
}

int32_t wgs_set_shard(wgs_ctx* ctx, int64_t M_total, int64_t site_offset, wgs_allreduce_fn fn, void* user)
{
    ctx->M_total = M_total; ctx->site_offset = site_offset; ctx->fn = fn; ctx->user = user;
    return 0;
}

int32_t wgs_upload_gl(wgs_ctx* ctx, const float* L, int64_t M, int32_t N, int32_t which)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    if (which < 0 || which > 1) return fail(ctx, "which must be 0 or 1");
    if (M < 0 || N <= 0) return fail(ctx, "bad shape");
    if (which == 0) { if (ensure_structure(ctx, N)) return 1; }
    else if (N != ctx->N || M != ctx->Mg[0]) return fail(ctx, "down-sampled matrix must match the GL matrix shape");
    dev_free(ctx, ctx->G[which]);
    if (dev_alloc(ctx, &ctx->G[which], (size_t)std::max<long>(M, 1) * ctx->ldg)) return 1;
    ctx->Mg[which] = M;
    float2* G = ctx->G[which];
    return upload_rows<float2>(ctx, (const float2*)L, M, N, [&](float2* stage, long r0, long rows, cudaStream_t st) {
        LAUNCH("repack", repack_kernel, grid_for(rows * ctx->ldg, 256, ctx->num_sm * 16), 256, 0, st, stage, N,
               G + (size_t)r0 * ctx->ldg, ctx->ldg, ctx->d_ind_of_col, rows);
    });
}

int32_t wgs_upload_gl_async(wgs_ctx* ctx, const float* L, int64_t M, int32_t N)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    if (M < 0 || N <= 0) return fail(ctx, "bad shape");
    if (ensure_structure(ctx, N)) return 1;
    const int K = std::max(ctx->K, 1), ldg = ctx->ldg;
    // runs of consecutive Beagle columns inside each population slab: one strided DMA each
    struct Run { int col, ind, len; };
    std::vector<std::vector<Run>> runs(K);
    size_t nruns = 0;
    for (int k = 0; k < K; ++k) {
        const PopDesc pd = ctx->pops[k];
        for (int j = 0; j < pd.n; ++j) {
            const int c = pd.col0 + j, i = ctx->ind_of_col[c];
            if (!runs[k].empty() && runs[k].back().ind + runs[k].back().len == i) ++runs[k].back().len;
            else { runs[k].push_back(Run{c, i, 1}); ++nruns; }
        }
    }
    // interleaved ID files (many short runs) are better served by the chunked upload + permutation kernel
    if (nruns > (size_t)8 * K + 64 || opt(ctx, "upload_sync")) return wgs_upload_gl(ctx, L, M, N, 0);
    dev_free(ctx, ctx->G[0]);
    if (dev_alloc(ctx, &ctx->G[0], (size_t)std::max<long>(M, 1) * ldg)) return 1;
    ctx->Mg[0] = M;
    float2* G = ctx->G[0];
    while ((int)ctx->ev_pop.size() < K) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_pop.push_back(e);
    }
    std::vector<int> pads;
    for (int c = 0; c < ldg; ++c) if (ctx->ind_of_col[c] < 0) pads.push_back(c);
    if (!pads.empty() && M > 0) {
        DevBuf dp;
        if (buf_alloc(ctx, dp, pads.size() * sizeof(int))) return 1;
        CU(cudaMemcpyAsync(dp.p, pads.data(), pads.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH("repack", pad_fill_kernel, grid_for(M * (long)pads.size(), 256, ctx->num_sm * 8), 256, 0, ctx->stream, G, ldg, (long)M,
               dp.as<int>(), (int)pads.size());
        CU(cudaStreamSynchronize(ctx->stream));                 // pads in place (and dp released) before any consumer is queued
    }
    // smallest population first: the pipelined operator starts working after the shortest possible wait
    ctx->upload_order.resize(K);
    for (int k = 0; k < K; ++k) ctx->upload_order[k] = k;
    std::stable_sort(ctx->upload_order.begin(), ctx->upload_order.end(), [&](int a, int b) { return ctx->pops[a].n < ctx->pops[b].n; });
    for (int k : ctx->upload_order) {
        if (M > 0)
            for (const Run& r : runs[k])
                CU(cudaMemcpy2DAsync(G + r.col, (size_t)ldg * sizeof(float2), L + 2 * (size_t)r.ind, (size_t)N * sizeof(float2),
                                     (size_t)r.len * sizeof(float2), (size_t)M, cudaMemcpyHostToDevice, ctx->stream2));
        CU(cudaEventRecord(ctx->ev_pop[k], ctx->stream2));
    }
    ctx->upload_pending = true;
    return 0;
}

// Row-block form of the asynchronous upload, for a matrix that is still being parsed: begin (allocate for up to
// M_capacity rows), rows (queue the strided DMAs of a finished row block - one per population slab run - and return
// at once), end (the final row count; from then on the context behaves as after wgs_upload_gl_async).
int32_t wgs_upload_gl_begin(wgs_ctx* ctx, int64_t M_capacity, int32_t N)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    if (M_capacity < 0 || N <= 0) return fail(ctx, "bad shape");
    if (ensure_structure(ctx, N)) return 1;
    const int ldg = ctx->ldg;
    dev_free(ctx, ctx->G[0]);
    if (dev_alloc(ctx, &ctx->G[0], (size_t)std::max<long>(M_capacity, 1) * ldg)) return 1;
    ctx->Mg[0] = 0;
    ctx->upload_capacity = M_capacity;
    // the padding columns of the slabs are not part of any DMA: zero the whole buffer once (device-side, cheap)
    CU(cudaMemsetAsync(ctx->G[0], 0, (size_t)std::max<long>(M_capacity, 1) * ldg * sizeof(float2), ctx->stream2));
    return 0;
}

int32_t wgs_upload_gl_rows(wgs_ctx* ctx, const float* L_rows, int64_t row0, int64_t nrows)
{
    cudaSetDevice(ctx->device);
    if (!ctx->G[0] || row0 < 0 || row0 + nrows > ctx->upload_capacity) return fail(ctx, "row block outside the capacity given to wgs_upload_gl_begin");
    const int K = std::max(ctx->K, 1), ldg = ctx->ldg, N = ctx->N;
    float2* G = ctx->G[0] + (size_t)row0 * ldg;
    for (int k = 0; k < K && nrows > 0; ++k) {
        const PopDesc pd = ctx->pops[k];
        int j = 0;
        while (j < pd.n) {                                       // runs of consecutive Beagle columns inside the slab
            const int c = pd.col0 + j, i = ctx->ind_of_col[c];
            int len = 1;
            while (j + len < pd.n && ctx->ind_of_col[c + len] == i + len) ++len;
            CU(cudaMemcpy2DAsync(G + c, (size_t)ldg * sizeof(float2), L_rows + 2 * (size_t)i, (size_t)N * sizeof(float2),
                                 (size_t)len * sizeof(float2), (size_t)nrows, cudaMemcpyHostToDevice, ctx->stream2));
            j += len;
        }
    }
    ctx->upload_pending = true;
    return 0;
}

int32_t wgs_upload_gl_end(wgs_ctx* ctx, int64_t M_final)
{
    cudaSetDevice(ctx->device);
    if (!ctx->G[0] || M_final < 0 || M_final > ctx->upload_capacity) return fail(ctx, "final row count outside the capacity given to wgs_upload_gl_begin");
    const int K = std::max(ctx->K, 1);
    ctx->Mg[0] = M_final;
    while ((int)ctx->ev_pop.size() < K) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_pop.push_back(e);
    }
    ctx->upload_order.resize(K);
    for (int k = 0; k < K; ++k) { ctx->upload_order[k] = k; CU(cudaEventRecord(ctx->ev_pop[k], ctx->stream2)); }
    ctx->upload_pending = true;
    return 0;
}

int32_t wgs_upload_wait(wgs_ctx* ctx)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    CU(cudaGetLastError());
    return 0;
}

int32_t wgs_upload_ad(wgs_ctx* ctx, const int32_t* AD, int64_t M, int32_t N)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    if (N != ctx->N || M != ctx->Mg[0]) return fail(ctx, "allele-depth matrix must match the GL matrix shape (upload GL first)");
    dev_free(ctx, ctx->AD);
    if (dev_alloc(ctx, &ctx->AD, (size_t)std::max<long>(M, 1) * ctx->ldg)) return 1;
    ctx->M_ad = M;
    DevBuf flag;
    if (buf_alloc(ctx, flag, 2 * sizeof(int))) return 1;
    CU(cudaMemsetAsync(flag.p, 0, 2 * sizeof(int), ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    uchar2* dst = ctx->AD;
    int rc = upload_rows<int2>(ctx, (const int2*)AD, M, N, [&](int2* stage, long r0, long rows, cudaStream_t st) {
        LAUNCH("repack", repack_ad_kernel, grid_for(rows * ctx->ldg, 256, ctx->num_sm * 16), 256, 0, st, stage, N,
               dst + (size_t)r0 * ctx->ldg, ctx->ldg, ctx->d_ind_of_col, rows, flag.as<int>());
    });
    if (rc) return rc;
    int bad[2] = {0, 0};
    CU(cudaMemcpy(bad, flag.p, 2 * sizeof(int), cudaMemcpyDeviceToHost));
    if (bad[0]) return fail(ctx, "negative allele depth in the depth matrix");
    ctx->ad_saturated = bad[1];                                  // stored as the "deeper than any class" sentinel
    return 0;
}

int32_t wgs_upload_ad_u8(wgs_ctx* ctx, const uint8_t* AD, int64_t M, int32_t N)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    if (N != ctx->N || M != ctx->Mg[0]) return fail(ctx, "allele-depth matrix must match the GL matrix shape (upload GL first)");
    dev_free(ctx, ctx->AD);
    if (dev_alloc(ctx, &ctx->AD, (size_t)std::max<long>(M, 1) * ctx->ldg)) return 1;
    ctx->M_ad = M;
    DevBuf flag;
    if (buf_alloc(ctx, flag, 2 * sizeof(int))) return 1;
    CU(cudaMemsetAsync(flag.p, 0, 2 * sizeof(int), ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    uchar2* dst = ctx->AD;
    int rc = upload_rows<uchar2>(ctx, (const uchar2*)AD, M, N, [&](uchar2* stage, long r0, long rows, cudaStream_t st) {
        LAUNCH("repack", repack_ad_u8_kernel, grid_for(rows * ctx->ldg, 256, ctx->num_sm * 16), 256, 0, st, stage, N,
               dst + (size_t)r0 * ctx->ldg, ctx->ldg, ctx->d_ind_of_col, rows, flag.as<int>());
    });
    if (rc) return rc;
    int bad[2] = {0, 0};
    CU(cudaMemcpy(bad, flag.p, 2 * sizeof(int), cudaMemcpyDeviceToHost));
    ctx->ad_saturated = bad[1];
    return 0;
}

int32_t wgs_synth(wgs_ctx* ctx, int64_t M, int32_t N, uint64_t seed, float depth, int32_t with_ad)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    if (ensure_structure(ctx, N)) return 1;
    drop_data(ctx);
    if (dev_alloc(ctx, &ctx->G[0], (size_t)M * ctx->ldg)) return 1;
    ctx->Mg[0] = M;
    if (with_ad) { if (dev_alloc(ctx, &ctx->AD, (size_t)M * ctx->ldg)) return 1; ctx->M_ad = M; }
    LAUNCH("synth", synth_kernel, grid_for(M * ctx->ldg, 256, ctx->num_sm * 16), 256, 0, ctx->stream, ctx->G[0], ctx->AD,
           ctx->ldg, (long)M, ctx->d_ind_of_col, ctx->d_pop_of_col, seed, depth, ctx->site_offset);
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return 0;
}

int32_t wgs_download(wgs_ctx* ctx, int64_t site0, int64_t nsites, float* L_out, int32_t* AD_out)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    if (site0 < 0 || site0 + nsites > ctx->Mg[0]) return fail(ctx, "site range outside the resident matrix");
    const int N = ctx->N;
    if (L_out) {
        DevBuf tmp;
        if (buf_alloc(ctx, tmp, (size_t)nsites * N * sizeof(float2))) return 1;
        LAUNCH("unpack", unpack_kernel, grid_for(nsites * N, 256, ctx->num_sm * 16), 256, 0, ctx->stream,
               ctx->G[0] + (size_t)site0 * ctx->ldg, ctx->ldg, ctx->d_col_of_ind, N, tmp.as<float2>(), (long)nsites);
        CU(cudaMemcpyAsync(L_out, tmp.p, (size_t)nsites * N * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (AD_out) {
        if (!ctx->AD) return fail(ctx, "no allele depths resident");
        DevBuf tmp;
        if (buf_alloc(ctx, tmp, (size_t)nsites * N * sizeof(int2))) return 1;
        LAUNCH("unpack", unpack_ad_kernel, grid_for(nsites * N, 256, ctx->num_sm * 16), 256, 0, ctx->stream,
               ctx->AD + (size_t)site0 * ctx->ldg, ctx->ldg, ctx->d_col_of_ind, N, tmp.as<int2>(), (long)nsites);
        CU(cudaMemcpyAsync(AD_out, tmp.p, (size_t)nsites * N * sizeof(int2), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    CU(cudaGetLastError());
    return 0;
}

// Per-population EM + clipping in three steps so that the fused operator can run the EM population by population
// as the slabs of an asynchronous upload arrive: begin (buffers), em (all populations or one), finish (clip,
// transpose to [M][K], copy out).
struct RefAfRun { DevBuf FT; std::vector<int> its; };
static int ref_af_begin(wgs_ctx* ctx, RefAfRun& run)
{
    if (!ctx->G[0]) return fail(ctx, "no GL matrix resident");
    if (!ctx->pops_set) return fail(ctx, "wgs_set_pops must be called before wgs_upload_gl for reference-panel operators");
    const long M = ctx->M();
    const int K = ctx->K;
    dev_free(ctx, ctx->d_af);
    if (dev_alloc(ctx, &ctx->d_af, (size_t)std::max<long>(M, 1) * K)) return 1;
    ctx->af_rows = M; ctx->af_cols = K;
    if (buf_alloc(ctx, run.FT, (size_t)std::max<long>(M, 1) * K * sizeof(float))) return 1;
    run.its.assign(K, 0);
    return 0;
}
static int ref_af_em_one(wgs_ctx* ctx, RefAfRun& run, int iter, double tole, int k)
{
    std::vector<int> it1;
    if (run_em_pop(ctx, iter, tole, run.FT.as<float>() + (size_t)k * ctx->M(), it1, k, 1)) return 1;
    run.its[k] = it1[0];
    return 0;
}
static int ref_af_finish(wgs_ctx* ctx, RefAfRun& run, float* af_out, int32_t* iters_out)
{
    const long M = ctx->M();
    const int K = ctx->K;
    float* F = ctx->d_af;
    std::vector<float> lo(K), hi(K);
    for (int k = 0; k < K; ++k) { double l = 1.0 / (2.0 * (ctx->pops[k].n + 1)); lo[k] = (float)l; hi[k] = (float)(1.0 - l); }
    DevBuf dlo, dhi;
    if (buf_alloc(ctx, dlo, K * sizeof(float)) || buf_alloc(ctx, dhi, K * sizeof(float))) return 1;
    CU(cudaMemcpyAsync(dlo.p, lo.data(), K * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dhi.p, hi.data(), K * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH("clip", clip_transpose_kernel, grid_for(M * K, 256, ctx->num_sm * 8), 256, 0, ctx->stream, run.FT.as<float>(), M, K,
           dlo.as<float>(), dhi.as<float>(), 1, F);
    if (af_out) CU(cudaMemcpyAsync(af_out, F, (size_t)M * K * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    for (int k = 0; k < K; ++k) iters_out[k] = run.its[k];
    return 0;
}
static int ref_af_impl(wgs_ctx* ctx, int32_t iter, double tole, float* af_out, int32_t* iters_out)
{
    Trace tr(ctx, "ref_af");
    RefAfRun run;
    if (ref_af_begin(ctx, run)) return 1;
    tr.lap("alloc");
    if (run_em_pop(ctx, iter, tole, run.FT.as<float>(), run.its)) return 1;
    tr.lap("em");
    if (ref_af_finish(ctx, run, af_out, iters_out)) return 1;
    tr.lap("clip+d2h");
    return 0;
}

int32_t wgs_ref_af(wgs_ctx* ctx, int32_t iter, double tole, float* af_out, int32_t* iters_out)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    return ref_af_impl(ctx, iter, tole, af_out, iters_out);
}

int32_t wgs_emMAF(wgs_ctx* ctx, const float* L_pop, int64_t M, int32_t n, int32_t iter, double tole, float* f_out, int32_t* iters_out)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    wgs_ctx* sub = nullptr;
    if (wgs_create(ctx->device, &sub)) return fail(ctx, "%s", wgs_last_error(nullptr));
    sub->timing = false;
    std::vector<int32_t> zeros(n, 0);
    int rc = wgs_set_pops(sub, zeros.data(), n, 1);
    if (!rc) rc = wgs_upload_gl(sub, L_pop, M, n, 0);
    if (!rc) {
        sub->M_total = ctx->M_total; sub->fn = ctx->fn; sub->user = ctx->user;
        DevBuf F;
        std::vector<int> its;
        auto body = [&](wgs_ctx* ctx) -> int {
            if (buf_alloc(ctx, F, (size_t)std::max<long>(M, 1) * sizeof(float))) return 1;
            if (run_em_pop(ctx, iter, tole, F.as<float>(), its)) return 1;
            CU(cudaMemcpy(f_out, F.p, (size_t)M * sizeof(float), cudaMemcpyDeviceToHost));
            return 0;
        };
        rc = body(sub);
        if (!rc) *iters_out = its[0];
    }
    ctx->launches += sub->launches;
    if (rc) ctx->err = sub->err;
    wgs_destroy(sub);
    return rc;
}

int32_t wgs_pop_like_partial(wgs_ctx* ctx, const float* af, int32_t K, double* out)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    if (!ctx->G[0]) return fail(ctx, "no GL matrix resident");
    if (K <= 0) return fail(ctx, "K must be positive");
    const long M = ctx->M();
    DevBuf dA, partials, sums;
    if (buf_alloc(ctx, dA, (size_t)M * K * sizeof(float))) return 1;
    CU(cudaMemcpyAsync(dA.p, af, (size_t)M * K * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    int R = 1;
    float margin = 0.f;
    if (af_R(ctx, dA.as<float>(), M * K, &R, &margin)) return 1;
    const int R2 = opt(ctx, "poplike_v1") ? 0 : pick_R2(margin);
    if (R2 > 0) {
        // ratio form (pop_like2): every AF inside [2^-16, 1-2^-16] - always true after the reference's clipping
        Like2Cfg c2 = like2_cfg(ctx, M, R2);
        size_t np2 = (size_t)ctx->ldg * K;
        if (buf_alloc(ctx, partials, (size_t)c2.gy * np2 * sizeof(double)) || buf_alloc(ctx, sums, np2 * sizeof(double))) return 1;
        for (int k0 = 0; k0 < K;) {
            int KT = pick_KT2(K - k0);
            if (launch_pop_like2(ctx, KT, ctx->G[0], M, dA.as<float>(), K, k0, c2, partials.as<double>())) return 1;
            k0 += KT;
        }
        LAUNCH("reduce", reduce_partials_kernel, grid_for(np2, 256, ctx->num_sm * 4), 256, 0, ctx->stream, partials.as<double>(), c2.gy,
               (long)np2, sums.as<double>());
        if (add_af_logsum(ctx, dA.as<float>(), M, K, sums.as<double>(), (long)np2)) return 1;
        if (cols_to_host(ctx, sums.as<double>(), K, out)) return 1;
        CU(cudaGetLastError());
        return 0;
    }
    // one launch geometry for all passes: the widest tile decides the individuals per thread
    int I = pop_like_inds(pick_KT(K), ctx->ldg);
    LikeCfg c = like_cfg(ctx, M, 2, I);
    size_t np = (size_t)ctx->ldg * K;
    if (buf_alloc(ctx, partials, (size_t)c.gy * np * sizeof(double)) || buf_alloc(ctx, sums, np * sizeof(double))) return 1;
    for (int k0 = 0; k0 < K;) {
        int KT = pick_KT(K - k0);
        if (I > 1 && pop_like_inds(KT, ctx->ldg) != I) KT = 20;     // keep I constant across passes (only when K > 20)
        int rc_ = 0;
        DISPATCH_KT(launch_pop_like_t, KT, R, ctx, ctx->G[0], M, dA.as<float>(), K, k0, c, partials.as<double>(), I);
        if (rc_) return rc_;
        k0 += KT;
    }
    LAUNCH("reduce", reduce_partials_kernel, grid_for(np, 256, ctx->num_sm * 4), 256, 0, ctx->stream, partials.as<double>(), c.gy,
           (long)np, sums.as<double>());
    if (cols_to_host(ctx, sums.as<double>(), K, out)) return 1;
    CU(cudaGetLastError());
    return 0;
}

// fused (wgs_ref_af_loo): the leave-one-out EM runs FIRST, population by population as the slabs of an
// asynchronous upload arrive; the full-data EM + clipping follow once the matrix is complete, and only then
// are the full-data columns of the LOO state filled in (the EM itself never reads them).
struct FusedRef { float* af_out; int32_t* iters_out; };
static int loo_impl(wgs_ctx* ctx, float* af_inout, int32_t iter, double tole, int32_t use_ds, int32_t parts,
                    double* ll, double* ll_parts, int32_t* iters_out, const FusedRef* fused)
{
    if (!ctx->G[0]) return fail(ctx, "no GL matrix resident");
    if (!ctx->pops_set) return fail(ctx, "wgs_set_pops must be called before wgs_upload_gl for reference-panel operators");
    if (use_ds && !ctx->G[1]) return fail(ctx, "down-sampled GL matrix not resident (wgs_upload_gl which=1)");
    if (parts < 1) return fail(ctx, "parts must be >= 1");
    const long M = ctx->M();
    const int K = ctx->K, ldg = ctx->ldg, N = ctx->N;
    const int ldf = ldg + (K + 3) / 4 * 4;
    Trace tr(ctx, "loo");
    DevBuf F, dcols;
    if (!fused && !af_inout && !(ctx->d_af && ctx->af_rows == M && ctx->af_cols == K))
        return fail(ctx, "af_inout is NULL but no allele-frequency matrix is resident (call wgs_ref_af first)");
    if (af_inout && !fused) {
        dev_free(ctx, ctx->d_af);
        if (dev_alloc(ctx, &ctx->d_af, (size_t)std::max<long>(M, 1) * K)) return 1;
        ctx->af_rows = M; ctx->af_cols = K;
        CU(cudaMemcpyAsync(ctx->d_af, af_inout, (size_t)M * K * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (buf_alloc(ctx, F, (size_t)M * ldf * sizeof(float)) || buf_alloc(ctx, dcols, K * sizeof(int))) return 1;
    std::vector<int> ident(K);
    for (int k = 0; k < K; ++k) ident[k] = k;
    CU(cudaMemcpyAsync(dcols.p, ident.data(), K * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    {   // columns [ldg, ldf): the caller's full-data AF, then pad columns that only idle lanes read (kept finite)
        DevBuf dpad;
        std::vector<float> half(ldf - ldg, 0.5f);
        if (buf_alloc(ctx, dpad, half.size() * sizeof(float))) return 1;
        CU(cudaMemcpyAsync(dpad.p, half.data(), half.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH("fill", bcast_row_kernel, grid_for(M * (ldf - ldg), 256, ctx->num_sm * 8), 256, 0, ctx->stream, F.as<float>() + ldg, ldf,
               ldf - ldg, M, dpad.as<float>());
        CU(cudaStreamSynchronize(ctx->stream));
    }
    std::vector<int> its_cols;
    CU(cudaStreamSynchronize(ctx->stream));
    tr.lap("alloc+h2d+init");
    const bool pipelined = fused && ctx->upload_pending;
    RefAfRun ref;
    if (fused && ref_af_begin(ctx, ref)) return 1;
    std::function<int(int)> on_pop = [&](int k) { return ref_af_em_one(ctx, ref, iter, tole, k); };
    if (run_em_loo(ctx, iter, tole, F.as<float>(), ldf, nullptr, nullptr, its_cols, nullptr,
                   pipelined ? &ctx->ev_pop : nullptr, pipelined ? &on_pop : nullptr)) return 1;
    tr.lap("em");
    if (fused) {
        upload_fence(ctx);                                       // every slab has been waited for on the device already
        if (!pipelined && run_em_pop(ctx, iter, tole, ref.FT.as<float>(), ref.its)) return 1;
        if (ref_af_finish(ctx, ref, fused->af_out, fused->iters_out)) return 1;
        tr.lap("ref_af");
    }
    RawBuf dA{ctx->d_af};
    LAUNCH("gather", gather_cols_kernel, grid_for(M * K, 256, ctx->num_sm * 8), 256, 0, ctx->stream, dA.as<float>(), K,
           dcols.as<int>(), K, F.as<float>(), ldf, ldg, M);
    for (int i = 0; i < N; ++i) iters_out[i] = its_cols[ctx->col_of_ind[i]];

    // clip to [1/(2n), 1-1/(2n)] with n = n_pop (glassy.py:80-85: n_pop-1 individuals were used)
    // The staged likelihood kernel (loo_like2) clamps while it builds its cells and the final gather clamps on the
    // way out, so the state matrix itself is not rewritten; the gather-through-L1 fallback still clamps it in place.
    std::vector<float> lo, hi;
    clip_bounds(ctx, 1, lo, hi);
    LooLike2Cfg c2{};
    const bool staged = !opt(ctx, "loolike_v1") && loo_like2_cfg(ctx, M, ldf, K, &c2);   // staged also means: clamp on the way, F not rewritten
    lo.resize(ldf, 0.0f); hi.resize(ldf, 1.0f);                  // full-data AF and pad columns: already inside [0, 1]
    DevBuf dlo, dhi;
    if (buf_alloc(ctx, dlo, ldf * sizeof(float)) || buf_alloc(ctx, dhi, ldf * sizeof(float))) return 1;
    CU(cudaMemcpyAsync(dlo.p, lo.data(), ldf * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dhi.p, hi.data(), ldf * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (!staged)
        LAUNCH("clip", clip_cols_kernel, grid_for(M * ldg, 256, ctx->num_sm * 8), 256, 0, ctx->stream, F.as<float>(), ldf, ldg, M,
               dlo.as<float>(), dhi.as<float>());

    // which column of F each (individual, population) pair reads (glassy.py:89 overwrite order)
    std::vector<int> rc((size_t)ldg * K, ldf - 1), last(K, -1);
    for (int i = 0; i < N; ++i) {
        int c = ctx->col_of_ind[i], p = ctx->pop_of_ind[i];
        for (int j = 0; j < K; ++j)
            rc[(size_t)c * K + j] = (j == p) ? c : (last[j] >= 0 ? ctx->col_of_ind[last[j]] : ldg + j);
        last[p] = i;
    }
    DevBuf drc;
    if (buf_alloc(ctx, drc, rc.size() * sizeof(int))) return 1;
    CU(cudaMemcpyAsync(drc.p, rc.data(), rc.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));

    int R = 1;
    float margin_all = 0.5f;
    {   // margin over everything the likelihood can read: clipped LOO columns and the full-data AF
        int Ra = 1;
        float ma = 0.f;
        if (af_R(ctx, dA.as<float>(), M * K, &Ra, &ma)) return 1;
        float mlo = 0.5f;
        for (int c = 0; c < ldg; ++c) if (ctx->ind_of_col[c] >= 0) mlo = std::min(mlo, lo[c]);
        R = std::min(Ra, pick_R(mlo));
        for (int k = 0; k < K; ++k) if (ctx->pops[k].n <= 1) R = 1;
        margin_all = std::min(ma, mlo);
    }
    // run-ordered panels: 2K shared frequencies per site + each individual's own state value (loo_like3)
    LooLike3Plan p3;
    const bool v3 = staged && plan_loo_like3(ctx, M, margin_all, R, &p3);
    DevBuf dA2, dXY2, dsc, dps, dlc, dper, dC;
    if (v3) {
        const int W2 = 4 * p3.KP;
        const long Mpad = (std::max<long>(M, 1) + kPL2TS - 1) / kPL2TS * kPL2TS;
        if (buf_alloc(ctx, dA2, (size_t)std::max<long>(M, 1) * W2 * sizeof(float)) || buf_alloc(ctx, dXY2, (size_t)Mpad * 2 * p3.KP * sizeof(ulonglong2)) ||
            buf_alloc(ctx, dsc, ldg * sizeof(int)) || buf_alloc(ctx, dps, K * sizeof(int)) || buf_alloc(ctx, dlc, K * sizeof(int)) ||
            buf_alloc(ctx, dper, (size_t)ctx->num_sm * 4 * W2 * sizeof(double)) || buf_alloc(ctx, dC, (size_t)W2 * sizeof(double))) return 1;
        CU(cudaMemcpyAsync(dsc.p, p3.slot_of_col.data(), ldg * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(dps.p, p3.pop_of_slot.data(), K * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(dlc.p, p3.lastcol_of_slot.data(), K * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH("loo_like_aux", loo_shared_af_kernel, grid_for(M * W2, 256, ctx->num_sm * 8), 256, 0, ctx->stream, F.as<float>(), ldf, ldg, M, K, p3.KP,
               dlc.as<int>(), dps.as<int>(), dlo.as<float>(), dhi.as<float>(), dA2.as<float>());
        LAUNCH("loo_like_aux", xy_precompute_kernel, grid_for(M * 2 * p3.KP, 256, ctx->num_sm * 8), 256, 0, ctx->stream, dA2.as<float>(), M, W2, 0,
               2 * p3.KP, dXY2.as<ulonglong2>());
    }
    LikeCfg c = like_cfg(ctx, M, 3);
    const int n_split = v3 ? p3.c2.gy : (staged ? c2.gy : c.gy);
    size_t np = (size_t)ldg * K;
    DevBuf partials, sums;
    if (buf_alloc(ctx, partials, (size_t)n_split * np * sizeof(double)) || buf_alloc(ctx, sums, np * sizeof(double))) return 1;
    const float2* Gsrc = use_ds ? ctx->G[1] : ctx->G[0];
    std::vector<double> tmp((size_t)N * K);
    for (int pass = 0; pass < (parts > 1 ? parts + 1 : 1); ++pass) {
        long pm = pass == 0 ? 1 : parts, pr = pass == 0 ? 0 : pass - 1;
        if (v3) {
            if (launch_loo_like3(ctx, p3, Gsrc, M, F.as<float>(), ldf, dlo.as<float>(), dhi.as<float>(), dXY2.as<ulonglong2>(), dsc.as<int>(),
                                 dps.as<int>(), pm, pr, partials.as<double>())) return 1;
            LAUNCH("reduce", reduce_partials_kernel, grid_for(np, 256, ctx->num_sm * 4), 256, 0, ctx->stream, partials.as<double>(),
                   n_split, (long)np, sums.as<double>());
            // + the population-only term log(2a(1-a)) of the ratio form, summed over this pass's sites per shared frequency
            const int W2 = 4 * p3.KP, block = W2 * (256 / W2), grid = ctx->num_sm * 4;
            LAUNCH("loo_like_aux", af_logsum_kernel, grid, block, block * sizeof(double), ctx->stream, dA2.as<float>(), M * W2, W2, dper.as<double>(),
                   pm, pr, ctx->site_offset);
            LAUNCH("loo_like_aux", af_logsum_reduce_kernel, (W2 + 127) / 128, 128, 0, ctx->stream, dper.as<double>(), grid, W2, dC.as<double>());
            LAUNCH("loo_like_aux", add_loo_const_kernel, grid_for(np, 256, ctx->num_sm * 4), 256, 0, ctx->stream, sums.as<double>(), ldg, K, p3.KP,
                   dsc.as<int>(), dps.as<int>(), dC.as<double>());
        }
        for (int k0 = 0; k0 < K && !v3;) {
            int KT = staged ? loo_like2_KT(c2, K - k0) : pick_KT(K - k0);
            int rc_ = 0;
            if (staged) rc_ = launch_loo_like2(ctx, KT, Gsrc, M, F.as<float>(), ldf, dlo.as<float>(), dhi.as<float>(), drc.as<int>(), K, k0, c2, pm, pr, R,
                                               partials.as<double>());
            else DISPATCH_KT(launch_loo_like_t, KT, R, ctx, Gsrc, M, F.as<float>(), ldf, drc.as<int>(), K, k0, c, pm, pr, partials.as<double>());
            if (rc_) return rc_;
            k0 += KT;
        }
        if (!v3)
            LAUNCH("reduce", reduce_partials_kernel, grid_for(np, 256, ctx->num_sm * 4), 256, 0, ctx->stream, partials.as<double>(),
                   n_split, (long)np, sums.as<double>());
        if (pass == 0) { if (cols_to_host(ctx, sums.as<double>(), K, ll)) return 1; }
        else {
            if (!ll_parts) return fail(ctx, "ll_parts is NULL but parts > 1");
            if (cols_to_host(ctx, sums.as<double>(), K, tmp.data())) return 1;
            for (int i = 0; i < N; ++i)
                memcpy(ll_parts + ((size_t)i * parts + pr) * K, tmp.data() + (size_t)i * K, K * sizeof(double));
        }
    }
    if (parts == 1 && ll_parts) memcpy(ll_parts, ll, (size_t)N * K * sizeof(double));
    tr.lap("like");

    // af as the reference leaves it: column j = LOO estimate of the LAST member of population j
    std::vector<int> lastcol(K);
    for (int j = 0; j < K; ++j) lastcol[j] = last[j] >= 0 ? ctx->col_of_ind[last[j]] : ldg + j;
    CU(cudaMemcpyAsync(dcols.p, lastcol.data(), K * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH("gather", gather_cols_kernel, grid_for(M * K, 256, ctx->num_sm * 8), 256, 0, ctx->stream, F.as<float>(), ldf,
           dcols.as<int>(), K, dA.as<float>(), K, 0, M, staged ? dlo.as<float>() : nullptr, staged ? dhi.as<float>() : nullptr);
    if (af_inout) CU(cudaMemcpyAsync(af_inout, dA.p, (size_t)M * K * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    tr.lap("af d2h");
    return 0;
}

int32_t wgs_loo_partial(wgs_ctx* ctx, float* af_inout, int32_t iter, double tole, int32_t use_ds, int32_t parts,
                        double* ll, double* ll_parts, int32_t* iters_out)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    return loo_impl(ctx, af_inout, iter, tole, use_ds, parts, ll, ll_parts, iters_out, nullptr);
}

int32_t wgs_ref_af_loo(wgs_ctx* ctx, int32_t iter, double tole, float* af_out, int32_t* af_iters_out, float* af_after_loo,
                       int32_t use_ds, int32_t parts, double* ll, double* ll_parts, int32_t* loo_iters_out)
{
    cudaSetDevice(ctx->device);                                  // no upload fence: the slabs are waited for one by one
    if (!af_iters_out || !ll || !loo_iters_out) return fail(ctx, "af_iters_out, ll and loo_iters_out must not be NULL");
    FusedRef fr{af_out, af_iters_out};
    int rc = loo_impl(ctx, af_after_loo, iter, tole, use_ds, parts, ll, ll_parts, loo_iters_out, &fr);
    upload_fence(ctx);                                           // error paths included: the host matrix is free again on return
    return rc;
}

int32_t wgs_fisher_partial(wgs_ctx* ctx, const float* af, float* f_obs, float* ne_obs, double* ne_ind_sum)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    if (!ctx->G[0]) return fail(ctx, "no GL matrix resident");
    if (!ctx->pops_set) return fail(ctx, "wgs_set_pops must be called before wgs_upload_gl for reference-panel operators");
    const long M = ctx->M();
    const int K = ctx->K, ldg = ctx->ldg;
    int nmax = 1;
    for (int k = 0; k < K; ++k) nmax = std::max(nmax, ctx->pops[k].n);
    int row16 = (nmax + 3) / 4 * 2;
    while (row16 % 8 != kEmT % 8) ++row16;
    const int accw_ld = (nmax + 3) / 4 * 4 + 4;
    int R = 128;
    while (R > 8 && 2 * (size_t)R * row16 * 16 > 100 * 1024) R /= 2;
    const int nw = R * kEmT / 32;
    size_t smem = 2 * (size_t)R * row16 * 16 + (size_t)nw * accw_ld * sizeof(float);
    if (smem > 220 * 1024) return fail(ctx, "population of %d individuals exceeds the Fisher shared-memory tile", nmax);
    CU(cudaFuncSetAttribute(fisher_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    int occ = 1;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fisher_kernel, R * kEmT, smem));
    long ntiles = (M + R - 1) / R;
    int nblocks = (int)std::max<long>(1, std::min<long>(ntiles, ((long)ctx->num_sm * std::max(occ, 1)) / K));
    // barrier-free register-tile kernel (fisher2) for populations of up to 512 individuals
    int tpr = 2;
    while (tpr < 32 && tpr * kFisherQ < (nmax + 1) / 2) tpr *= 2;
    const bool v2 = !opt(ctx, "fisher_v1") && tpr * kFisherQ >= (nmax + 1) / 2;
    const size_t smem2 = (size_t)(256 / tpr) * 2 * tpr * kFisherQ * sizeof(float);
    if (v2) {
        int occ2 = 1;
#define FISHER2_OCC(T) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, fisher2_kernel<T>, 256, smem2))
        if (tpr == 2) FISHER2_OCC(2); else if (tpr == 4) FISHER2_OCC(4); else if (tpr == 8) FISHER2_OCC(8);
        else if (tpr == 16) FISHER2_OCC(16); else FISHER2_OCC(32);
#undef FISHER2_OCC
        const long ntiles2 = (M + 256 / tpr - 1) / (256 / tpr);
        nblocks = (int)std::max<long>(1, std::min<long>(ntiles2, ((long)ctx->num_sm * std::max(occ2, 1)) / K));
    }
    DevBuf dA, dF, dNe, partials, sums;
    if (buf_alloc(ctx, dA, (size_t)M * K * sizeof(float)) || buf_alloc(ctx, dF, (size_t)M * K * sizeof(float)) ||
        buf_alloc(ctx, dNe, (size_t)M * K * sizeof(float)) || buf_alloc(ctx, partials, (size_t)nblocks * ldg * sizeof(double)) ||
        buf_alloc(ctx, sums, ldg * sizeof(double))) return 1;
    CU(cudaMemcpyAsync(dA.p, af, (size_t)M * K * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(partials.p, 0, (size_t)nblocks * ldg * sizeof(double), ctx->stream));   // pad columns are never written
#define FISHER2_LAUNCH(T)                                                                                               \
    LAUNCH("fisher", fisher2_kernel<T>, dim3(nblocks, K), 256, smem2, ctx->stream, ctx->G[0], ldg, M, ctx->d_pops, K, dA.as<float>(), \
           dF.as<float>(), dNe.as<float>(), partials.as<double>())
    if (v2) {
        if (tpr == 2) FISHER2_LAUNCH(2); else if (tpr == 4) FISHER2_LAUNCH(4); else if (tpr == 8) FISHER2_LAUNCH(8);
        else if (tpr == 16) FISHER2_LAUNCH(16); else FISHER2_LAUNCH(32);
    } else
        LAUNCH("fisher", fisher_kernel, dim3(nblocks, K), R * kEmT, smem, ctx->stream, ctx->G[0], ldg, M, ctx->d_pops, K, dA.as<float>(),
               dF.as<float>(), dNe.as<float>(), row16, accw_ld, partials.as<double>());
#undef FISHER2_LAUNCH
    add_work(ctx, "fisher", (double)M * ctx->N * 8.0 + (double)M * K * 12.0, (double)M * ctx->N);
    LAUNCH("reduce", reduce_partials_kernel, grid_for(ldg, 256, 64), 256, 0, ctx->stream, partials.as<double>(), nblocks, (long)ldg,
           sums.as<double>());
    CU(cudaMemcpyAsync(f_obs, dF.p, (size_t)M * K * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(ne_obs, dNe.p, (size_t)M * K * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (nccl_sums(ctx) && dev_all_sum(ctx, sums.p, ldg, WGS_F64)) return 1;
    std::vector<double> h(ldg);
    CU(cudaMemcpyAsync(h.data(), sums.p, ldg * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    for (int i = 0; i < ctx->N; ++i) ne_ind_sum[i] = h[ctx->col_of_ind[i]];
    return 0;
}

// exact binomial coefficient (depth <= 40 fits easily in 64 bits)
static double binom(int n, int k)
{
    if (k > n - k) k = n - k;
    unsigned long long r = 1;
    for (int i = 1; i <= k; ++i) r = r * (unsigned long long)(n - k + i) / (unsigned long long)i;
    return (double)r;
}

int32_t wgs_zscore(wgs_ctx* ctx, int32_t mode, const float* af, int32_t K, int32_t n_threshold, int32_t single_read,
                   int32_t ind_start, int32_t ind_end, int32_t iter, double tole, wgs_zrow* out)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    if (!ctx->G[0]) return fail(ctx, "no GL matrix resident");
    if (!ctx->AD || ctx->M_ad != ctx->M()) return fail(ctx, "no allele depths resident (wgs_upload_ad)");
    if (!ctx->pops_set && mode != 2) return fail(ctx, "wgs_set_pops must be called before wgs_upload_gl for the z-score operators");
    if (ind_start < 0 || ind_end > ctx->N || ind_start >= ind_end) return fail(ctx, "individual range [%d,%d) outside [0,%d)", ind_start, ind_end, ctx->N);
    if (mode < 0 || mode > 2) return fail(ctx, "mode must be 0 (assignment), 1 (reference) or 2 (preparation only)");
    if (mode == 0 && (!af || K != ctx->K)) return fail(ctx, "assignment mode needs af [M,%d]", ctx->K);
    const long M = ctx->M();
    const int ldg = ctx->ldg, N = ctx->N;
    const double e = 0.01;                                       // WGSassign.py:350, :430
    Trace tr(ctx, "zscore");

    std::vector<unsigned char> sel(ldg, 0);
    for (int i = ind_start; i < ind_end; ++i) sel[ctx->col_of_ind[i]] = 1;
    DevBuf dsel, dtable, ddeep;
    const size_t tab_n = (size_t)ldg * kZClasses;
    if (buf_alloc(ctx, dsel, ldg) || buf_alloc(ctx, dtable, tab_n * sizeof(ZTally)) || buf_alloc(ctx, ddeep, ldg * sizeof(unsigned long long))) return 1;
    CU(cudaMemcpyAsync(dsel.p, sel.data(), ldg, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(dtable.p, 0, tab_n * sizeof(ZTally), ctx->stream));
    CU(cudaMemsetAsync(ddeep.p, 0, ldg * sizeof(unsigned long long), ctx->stream));

    // launch geometry shared by the three streaming passes
    LikeCfg c = like_cfg(ctx, M, 4);
    {
        int wy_count = 8 / c.wx;
        long cap = 2047L * wy_count;                              // 32-bit register sums in ztally
        if (c.sites_per_block > cap) { c.sites_per_block = cap; c.gy = (int)((M + cap - 1) / cap); }
    }
    const double pairs = (double)M * (ind_end - ind_start);
    // The reference's class means are SEQUENTIAL float32 sums in site order (zscore.py:22); ztally_ord reproduces
    // them for any number of sites, and under site sharding the ranks run it one after the other in site order,
    // handing the table on (carry-in / carry-out).  Option z_exact_means = 1 selects the order-independent
    // fixed-point tally instead (one pass, no chain; NOT the reference's means).
    const bool exact_means = opt(ctx, "z_exact_means") != 0;
    const bool sharded = ctx->fn != nullptr;
    if (sharded && !exact_means && ctx->world <= 1)
        return fail(ctx, "site-sharded z-score: wgs_set_rank (or wgs_nccl_init) must give this context its rank first");
    DevBuf dmaxd;
    if (buf_alloc(ctx, dmaxd, sizeof(int))) return 1;
    CU(cudaMemsetAsync(dmaxd.p, 0, sizeof(int), ctx->stream));
    if (exact_means) {
        // order-independent fixed-point tally (shardable)
        const size_t zsmem = (size_t)kZHotX * 7 * 256 * sizeof(int);
        CU(cudaFuncSetAttribute(ztally_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zsmem));
        LAUNCH("ztally", ztally_kernel, dim3(c.gx, c.gy), 256, zsmem, ctx->stream, ctx->G[0], ctx->AD, ldg, M, dsel.as<unsigned char>(),
               c.wx, c.sites_per_block, dtable.as<ZTally>(), ddeep.as<unsigned long long>());
        LAUNCH("zaux", zmaxdepth_kernel<ZTally>, ctx->num_sm * 4, 256, 0, ctx->stream, dtable.as<ZTally>(), (long)tab_n, dmaxd.as<int>());
    } else {
        // Column groups: a launch covers the blocks of one group; under sharding the groups pipeline through the ranks
        // (rank r works on group g while rank r+1 works on group g-1), so the chain costs (W + G - 1) / G tallies
        // instead of W.
        const int nblk = (ldg + kZOrdWarps - 1) / kZOrdWarps;
        CU(cudaFuncSetAttribute(ztally_ord_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kZOrdSmem));
        const bool sharded_nccl = sharded && ctx->nccl_comm;
        const int ngroups = sharded_nccl ? std::max(1, std::min(opt(ctx, "ztally_groups", 1), nblk)) : 1;
        auto run_group = [&](int g) -> int {
            const int b0 = (int)((long)nblk * g / ngroups), b1 = (int)((long)nblk * (g + 1) / ngroups);
            if (b1 <= b0) return 0;
            LAUNCH("ztally", ztally_ord_kernel, b1 - b0, kZOrdWarps * 32, kZOrdSmem, ctx->stream, ctx->G[0], ctx->AD, ldg, M,
                   dsel.as<unsigned char>(), b0 * kZOrdWarps, std::min(ldg, b1 * kZOrdWarps), dtable.as<ZTallyF>(), ddeep.as<unsigned long long>());
            return 0;
        };
        const size_t tbytes = tab_n * sizeof(ZTallyF);
        if (!sharded) {
            if (run_group(0)) return 1;
        } else if (ctx->nccl_comm) {
            // rank r: receive a group's table rows from r-1, add this rank's sites, send them on; the last rank's table is
            // the whole file's and is broadcast back (everything on the compute stream, over NVLink)
            const int r = ctx->rank, W = ctx->world;
            int rc_ = 0;
            for (int g = 0; g < ngroups && !rc_; ++g) {
                const int b0 = (int)((long)nblk * g / ngroups), b1 = (int)((long)nblk * (g + 1) / ngroups);
                const size_t c_lo = (size_t)b0 * kZOrdWarps, c_hi = std::min<size_t>(ldg, (size_t)b1 * kZOrdWarps);
                if (c_hi <= c_lo) continue;
                char* slice = (char*)dtable.p + c_lo * kZClasses * sizeof(ZTallyF);
                const size_t sbytes = (c_hi - c_lo) * kZClasses * sizeof(ZTallyF);
                if (r > 0) rc_ = g_nccl.Recv(slice, sbytes, kNcclChar, r - 1, ctx->nccl_comm, ctx->stream);
                if (!rc_ && run_group(g)) return 1;
                if (!rc_ && r < W - 1) rc_ = g_nccl.Send(slice, sbytes, kNcclChar, r + 1, ctx->nccl_comm, ctx->stream);
            }
            if (!rc_) rc_ = g_nccl.Broadcast(dtable.p, dtable.p, tbytes, kNcclChar, W - 1, ctx->nccl_comm, ctx->stream);
            if (rc_) return fail(ctx, "NCCL table hand-over failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc_) : "?");
        } else {
            // no communicator: the same chain through the host callback - round q leaves rank q's table on every
            // rank (an integer sum in which only rank q contributes: exact bits)
            std::vector<long long> h(tbytes / 8);
            for (int q = 0; q < ctx->world; ++q) {
                if (q == ctx->rank) {
                    if (run_group(0)) return 1;
                    CU(cudaMemcpyAsync(h.data(), dtable.p, tbytes, cudaMemcpyDeviceToHost, ctx->stream));
                    CU(cudaStreamSynchronize(ctx->stream));
                } else std::fill(h.begin(), h.end(), 0LL);
                ctx->fn(h.data(), (int64_t)h.size(), WGS_I64, ctx->user);
                if (q + 1 == ctx->rank || q + 1 == ctx->world)
                    CU(cudaMemcpyAsync(dtable.p, h.data(), tbytes, cudaMemcpyHostToDevice, ctx->stream));
                CU(cudaStreamSynchronize(ctx->stream));
            }
        }
        LAUNCH("zaux", zmaxdepth_kernel<ZTallyF>, ctx->num_sm * 4, 256, 0, ctx->stream, dtable.as<ZTallyF>(), (long)tab_n, dmaxd.as<int>());
    }
    add_work(ctx, "ztally", pairs * 10.0, pairs);
    // Only the classes up to the deepest observed depth travel and are scanned: compact host tables [ldg][ncls].
    // The sequential tally leaves the same (whole-file) table on every rank; the fixed-point tally is summed on the
    // host below, so under sharding every rank must keep the full table shape there.
    int dmax = kZDepthCap;
    if (!sharded || !exact_means) {
        CU(cudaMemcpyAsync(&dmax, dmaxd.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        dmax = std::max(1, std::min(dmax, kZDepthCap));
    }
    const int ncls = (dmax + 1) * (dmax + 2) / 2;
    const size_t ctab_n = (size_t)ldg * ncls;
    // class counts and float32 class means, per (column, class)
    std::vector<long long> ccnt(ctab_n, 0);
    std::vector<float> cmean(ctab_n * 3, 0.f);
    if (exact_means) {
        std::vector<ZTally> table(ctab_n);
        CU(cudaMemcpy2DAsync(table.data(), (size_t)ncls * sizeof(ZTally), dtable.p, (size_t)kZClasses * sizeof(ZTally),
                             (size_t)ncls * sizeof(ZTally), (size_t)ldg, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->fn) ctx->fn(table.data(), (int64_t)ctab_n * 4, WGS_I64, ctx->user);
        for (size_t g = 0; g < ctab_n; ++g) {
            ccnt[g] = table[g].cnt;
            if (table[g].cnt > 0) {
                double n = (double)table[g].cnt;
                cmean[3 * g + 0] = (float)((double)table[g].s0 * kZUnit / n);
                cmean[3 * g + 1] = (float)((double)table[g].s1 * kZUnit / n);
                cmean[3 * g + 2] = (float)((double)table[g].s2 * kZUnit / n);
            }
        }
    } else {
        std::vector<ZTallyF> table(ctab_n);
        CU(cudaMemcpy2DAsync(table.data(), (size_t)ncls * sizeof(ZTallyF), dtable.p, (size_t)kZClasses * sizeof(ZTallyF),
                             (size_t)ncls * sizeof(ZTallyF), (size_t)ldg, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        for (size_t g = 0; g < ctab_n; ++g) {
            ccnt[g] = table[g].cnt;
            if (table[g].cnt > 0) {
                float n = (float)table[g].cnt;
                cmean[3 * g + 0] = table[g].s0 / n; cmean[3 * g + 1] = table[g].s1 / n; cmean[3 * g + 2] = table[g].s2 / n;
            }
        }
    }

    // sites deeper than the class table (or saturated on upload): whole-file counts per individual, for the reference's
    // class-count asserts below and for wgs_zscore_deep_sites
    std::vector<unsigned long long> deep(ldg);
    if (nccl_sums(ctx) && dev_all_sum(ctx, ddeep.p, ldg, WGS_I64)) return 1;
    CU(cudaMemcpyAsync(deep.data(), ddeep.p, ldg * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->fn && !nccl_sums(ctx)) ctx->fn(deep.data(), (int64_t)ldg, WGS_I64, ctx->user);
    ctx->z_deep_sites = 0;
    for (int col = 0; col < ldg; ++col) ctx->z_deep_sites += (long)deep[col];
    tr.lap("tally+d2h");
    // ---- class decisions on the host (zscore.py:23-39, :63-79): a few hundred rows per individual ----
    std::vector<signed char> kmax(ctab_n, -1);
    std::vector<float> kmean(ctab_n, 0.f);
    std::vector<float4> zlike_t(ctab_n, make_float4(0.f, 0.f, 0.f, 0.f));     // class-major: [ncls][ldg]
    std::vector<float4> fac_of_class(ncls);                     // binomial read probabilities: the same for every individual
    for (int d = 0; d <= dmax; ++d)
        for (int alt = 0; alt <= d; ++alt) {
            const int ref = d - alt;
            const double comb = binom(d, alt);
            fac_of_class[zclass_id(ref, alt)] = make_float4((float)(comb * std::pow(1.0 - e, ref) * std::pow(e, alt)), (float)(comb * std::pow(0.5, d)),
                                                            (float)(comb * std::pow(1.0 - e, alt) * std::pow(e, ref)), 0.f);
        }
    ctx->zclasses.assign(N, std::vector<int>());
    ctx->ztable.assign(N, std::vector<float>());
    std::vector<int> n_classes(N, 0);
    for (int i = ind_start; i < ind_end; ++i) {
        const int col = ctx->col_of_ind[i];
        const long long* t = ccnt.data() + (size_t)col * ncls;
        int n_pass = 0;                                           // classes passing the count filter
        int per_depth[kZDepthCap + 1] = {0};
        for (int d = 0; d <= dmax; ++d)
            for (int alt = 0; alt <= d; ++alt) {
                int id = zclass_id(d - alt, alt);
                if (t[id] <= 0) continue;
                bool pass = single_read ? (d == 1) : (t[id] > n_threshold && d != 0);
                if (pass) { ++n_pass; ++per_depth[d]; }
            }
        // In the reference every distinct (ref, alt) pair is a class, however deep: a site deeper than our table is
        // at least one more class that passes a zero count threshold (it can never be KEPT: its depth would need
        // all depth+1 splits) - it only matters for these two asserts (zscore.py:34-35).
        if (!single_read && n_threshold <= 0 && deep[col] > 0) ++n_pass;
        if (n_pass == 0) return fail(ctx, "No loci were kept! Too stringent filtering?");
        if (n_pass == 1) return fail(ctx, "Not enough loci were kept! Too stringent filtering?");
        auto& rows = ctx->zclasses[i];
        for (int d = 0; d <= dmax; ++d) {
            if (!(d < per_depth[d])) continue;                   // zscore.py:38: depth kept iff all d+1 splits present
            for (int alt = 0; alt <= d; ++alt) {
                int ref = d - alt, id = zclass_id(ref, alt);
                size_t g = (size_t)col * ncls + id;
                float m0 = cmean[3 * g], m1 = cmean[3 * g + 1], m2 = cmean[3 * g + 2];
                int mx = 0; float mv = m0;
                if (m1 > mv) { mx = 1; mv = m1; }
                if (m2 > mv) { mx = 2; mv = m2; }
                kmax[(size_t)id * ldg + col] = (signed char)mx; kmean[(size_t)id * ldg + col] = mv;
                zlike_t[(size_t)id * ldg + col] = make_float4(m0, m1, m2, 0.f);
                rows.push_back(ref); rows.push_back(alt); rows.push_back(d); rows.push_back((int)t[id]);
                ++n_classes[i];
            }
        }
        {   // every observed class, for the AD_summary drop-in (zscore.py:20-22)
            auto& tab = ctx->ztable[i];
            for (int d = 0; d <= dmax; ++d)
                for (int alt = 0; alt <= d; ++alt) {
                    int id = zclass_id(d - alt, alt);
                    if (t[id] <= 0) continue;
                    size_t g = (size_t)col * ncls + id;
                    float row[7] = {(float)(d - alt), (float)alt, (float)t[id], cmean[3 * g], cmean[3 * g + 1], cmean[3 * g + 2],
                                    kmax[(size_t)id * ldg + col] >= 0 ? 1.f : 0.f};
                    tab.insert(tab.end(), row, row + 7);
                }
        }
        if (n_classes[i] == 0 && mode != 2) return fail(ctx, "individual %d: no read depth has all of its allele-count splits (zscore.py:36-39 leaves AD_array empty)", i);
    }

    tr.lap("class decisions");
    DevBuf dkmax, dkmean, dlike, dfac, dkeep, dkept;
    if (buf_alloc(ctx, dkmax, ctab_n) || buf_alloc(ctx, dkmean, ctab_n * sizeof(float)) || buf_alloc(ctx, dlike, ctab_n * sizeof(float4)) ||
        buf_alloc(ctx, dfac, (size_t)ncls * sizeof(float4)) || buf_alloc(ctx, dkeep, (size_t)std::max<long>(M, 1) * ldg) ||
        buf_alloc(ctx, dkept, ldg * sizeof(unsigned long long))) return 1;
    // class-major host tables [ncls][ldg] go up as they are; classes deeper than dmax do not occur (the kernels treat them as "not kept")
    CU(cudaMemcpyAsync(dkmax.p, kmax.data(), ctab_n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dkmean.p, kmean.data(), ctab_n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dlike.p, zlike_t.data(), ctab_n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dfac.p, fac_of_class.data(), (size_t)ncls * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(dkept.p, 0, ldg * sizeof(unsigned long long), ctx->stream));
    LAUNCH("zkeep", zkeep_kernel, dim3(c.gx, c.gy), 256, 0, ctx->stream, ctx->G[0], ctx->AD, ldg, M, dsel.as<unsigned char>(),
           dkmax.as<signed char>(), dkmean.as<float>(), ncls, c.wx, c.sites_per_block, dkeep.as<unsigned char>(), dkept.as<unsigned long long>());
    add_work(ctx, "zkeep", pairs * 11.0, pairs);
    std::vector<long long> kept(ldg);
    if (nccl_sums(ctx) && dev_all_sum(ctx, dkept.p, ldg, WGS_I64)) return 1;
    CU(cudaMemcpyAsync(kept.data(), dkept.p, ldg * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->fn && !nccl_sums(ctx)) ctx->fn(kept.data(), ldg, WGS_I64, ctx->user);

    tr.lap("tables h2d+keep");
    if (mode == 2) {                                             // preparation only: tallies, class tables, kept-site counts
        for (int i = ind_start; i < ind_end; ++i) {
            wgs_zrow& r = out[i - ind_start];
            r.z = r.w_obs = r.z_mu = r.z_var = std::numeric_limits<float>::quiet_NaN();
            r.loci_kept = kept[ctx->col_of_ind[i]];
            r.n_classes = n_classes[i];
            r.em_iters = 0;
        }
        return 0;
    }

    // ---- allele frequencies seen by each individual ----
    DevBuf dAF, dafcol;
    std::vector<int> afcol(ldg, 0), em_iters(ldg, 0);
    int af_ld = 0;
    if (buf_alloc(ctx, dafcol, ldg * sizeof(int))) return 1;
    if (mode == 0) {
        af_ld = ctx->K;
        if (buf_alloc(ctx, dAF, (size_t)std::max<long>(M, 1) * af_ld * sizeof(float))) return 1;
        CU(cudaMemcpyAsync(dAF.p, af, (size_t)M * af_ld * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        for (int col = 0; col < ldg; ++col) afcol[col] = ctx->pop_of_col[col];
    } else {
        af_ld = ldg;
        if (buf_alloc(ctx, dAF, (size_t)std::max<long>(M, 1) * af_ld * sizeof(float))) return 1;
        std::vector<double> cnt(ldg);
        std::vector<unsigned char> sel_em(sel);
        for (int col = 0; col < ldg; ++col) { cnt[col] = (double)kept[col]; if (kept[col] == 0) sel_em[col] = 0; }
        DevBuf dcnt;
        if (buf_alloc(ctx, dcnt, ldg * sizeof(double))) return 1;
        CU(cudaMemcpyAsync(dcnt.p, cnt.data(), ldg * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if (run_em_loo(ctx, iter, tole, dAF.as<float>(), af_ld, dkeep.as<unsigned char>(), dcnt.as<double>(), em_iters, &sel_em)) return 1;
        std::vector<float> lo, hi;
        clip_bounds(ctx, 1, lo, hi);                              // WGSassign.py:360-364
        DevBuf dlo, dhi;
        if (buf_alloc(ctx, dlo, ldg * sizeof(float)) || buf_alloc(ctx, dhi, ldg * sizeof(float))) return 1;
        CU(cudaMemcpyAsync(dlo.p, lo.data(), ldg * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(dhi.p, hi.data(), ldg * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH("clip", clip_cols_kernel, grid_for(M * ldg, 256, ctx->num_sm * 8), 256, 0, ctx->stream, dAF.as<float>(), af_ld, ldg, M,
               dlo.as<float>(), dhi.as<float>());
        for (int col = 0; col < ldg; ++col) afcol[col] = col;
    }
    CU(cudaMemcpyAsync(dafcol.p, afcol.data(), ldg * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    tr.lap("af / loo em");

    DevBuf partials, sums;
    const size_t np3 = (size_t)ldg * 3;
    if (buf_alloc(ctx, partials, (size_t)c.gy * np3 * sizeof(double)) || buf_alloc(ctx, sums, np3 * sizeof(double))) return 1;
    LAUNCH("zmoments", zmoments_kernel, dim3(c.gx, c.gy), 256, (size_t)ncls * sizeof(float4), ctx->stream, ctx->G[0], ctx->AD, ldg, M,
           dkeep.as<unsigned char>(), dAF.as<float>(), af_ld, dafcol.as<int>(), dlike.as<float4>(), dfac.as<float4>(), ncls, c.wx,
           c.sites_per_block, partials.as<double>());
    add_work(ctx, "zmoments", pairs * 15.0, pairs);
    LAUNCH("reduce", reduce_partials_kernel, grid_for(np3, 256, 64), 256, 0, ctx->stream, partials.as<double>(), c.gy, (long)np3, sums.as<double>());
    if (nccl_sums(ctx) && dev_all_sum(ctx, sums.p, np3, WGS_F64)) return 1;
    std::vector<double> h(np3);
    CU(cudaMemcpyAsync(h.data(), sums.p, np3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    tr.lap("moments");
    if (ctx->fn && !nccl_sums(ctx)) ctx->fn(h.data(), (int64_t)np3, WGS_F64, ctx->user);

    for (int i = ind_start; i < ind_end; ++i) {
        const int col = ctx->col_of_ind[i];
        wgs_zrow& r = out[i - ind_start];
        r.w_obs = (float)h[(size_t)col * 3 + 0];                  // zscore.py:100 / WGSassign.py:369-370: float32 sums
        r.z_mu = (float)h[(size_t)col * 3 + 1];
        r.z_var = (float)h[(size_t)col * 3 + 2];
        r.z = (r.w_obs - r.z_mu) / sqrtf(r.z_var);                // WGSassign.py:371 in float32 (NaN when nothing was kept)
        r.loci_kept = kept[col];
        r.n_classes = n_classes[i];
        r.em_iters = mode == 1 ? em_iters[col] : 0;
    }
    return 0;
}

int32_t wgs_zscore_table(wgs_ctx* ctx, int32_t ind, int32_t max_rows, float* rows_out, int32_t* n_rows)
{
    if (ind < 0 || ind >= (int)ctx->ztable.size()) return fail(ctx, "no class table for individual %d", ind);
    const auto& v = ctx->ztable[ind];
    int n = (int)v.size() / 7;
    *n_rows = n;
    for (int r = 0; r < std::min(n, max_rows); ++r) memcpy(rows_out + 7 * r, v.data() + 7 * r, 7 * sizeof(float));
    return 0;
}

int32_t wgs_zkeep_one(wgs_ctx* ctx, int32_t ind, int32_t n_classes, const int32_t* ad_array, const float* class_means,
                      int32_t* keep_out, int64_t cap, int64_t* n_kept)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    if (!ctx->G[0] || !ctx->AD) return fail(ctx, "GL matrix and allele depths must be resident");
    if (ind < 0 || ind >= ctx->N) return fail(ctx, "individual %d outside [0,%d)", ind, ctx->N);
    const long M = ctx->M();
    const int ldg = ctx->ldg, col = ctx->col_of_ind[ind];
    const size_t tab_n = (size_t)ldg * kZClasses;
    std::vector<signed char> kmax(tab_n, -1);
    std::vector<float> kmean(tab_n, 0.f);
    for (int c = 0; c < n_classes; ++c) {
        int ref = ad_array[4 * c], alt = ad_array[4 * c + 1];
        if (ref < 0 || alt < 0 || ref + alt > kZDepthCap) continue;
        const float* m = class_means + 3 * c;
        int mx = 0; float mv = m[0];
        if (m[1] > mv) { mx = 1; mv = m[1]; }
        if (m[2] > mv) { mx = 2; mv = m[2]; }
        size_t g = (size_t)zclass_id(ref, alt) * ldg + col;     // class-major, like wgs_zscore's tables
        kmax[g] = (signed char)mx; kmean[g] = mv;
    }
    std::vector<unsigned char> sel(ldg, 0);
    sel[col] = 1;
    DevBuf dsel, dkmax, dkmean, dkeep, dkept;
    if (buf_alloc(ctx, dsel, ldg) || buf_alloc(ctx, dkmax, tab_n) || buf_alloc(ctx, dkmean, tab_n * sizeof(float)) ||
        buf_alloc(ctx, dkeep, (size_t)std::max<long>(M, 1) * ldg) || buf_alloc(ctx, dkept, ldg * sizeof(unsigned long long))) return 1;
    CU(cudaMemcpyAsync(dsel.p, sel.data(), ldg, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dkmax.p, kmax.data(), tab_n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dkmean.p, kmean.data(), tab_n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(dkept.p, 0, ldg * sizeof(unsigned long long), ctx->stream));
    LikeCfg c = like_cfg(ctx, M, 4);
    LAUNCH("zkeep", zkeep_kernel, dim3(c.gx, c.gy), 256, 0, ctx->stream, ctx->G[0], ctx->AD, ldg, M, dsel.as<unsigned char>(),
           dkmax.as<signed char>(), dkmean.as<float>(), kZClasses, c.wx, c.sites_per_block, dkeep.as<unsigned char>(), dkept.as<unsigned long long>());
    std::vector<unsigned char> colmask((size_t)std::max<long>(M, 1));
    CU(cudaMemcpy2DAsync(colmask.data(), 1, dkeep.as<unsigned char>() + col, ldg, 1, (size_t)M, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    int64_t nk = 0;
    for (long s = 0; s < M; ++s)
        if (colmask[s]) { if (nk < cap) keep_out[nk] = (int32_t)s; ++nk; }
    *n_kept = nk;
    return 0;
}

int32_t wgs_zmoments_list(wgs_ctx* ctx, int32_t ind, const int32_t* L_keep, int64_t mk, const float* A_vec, int32_t n_classes,
                          const float* AD_factorial, const float* AD_like, const int32_t* AD_index, int32_t idx_rows, int32_t idx_cols,
                          float* W_obs_out, float* W_l_out, float* W_var_out)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    if (!ctx->G[0] || !ctx->AD) return fail(ctx, "GL matrix and allele depths must be resident");
    if (ind < 0 || ind >= ctx->N) return fail(ctx, "individual %d outside [0,%d)", ind, ctx->N);
    for (int64_t e = 0; e < mk; ++e) if (L_keep[e] < 0 || L_keep[e] >= ctx->M()) return fail(ctx, "L_keep[%lld] outside the matrix", (long long)e);
    for (int e = 0; e < idx_rows * idx_cols; ++e) if (AD_index[e] < 0 || AD_index[e] >= n_classes) return fail(ctx, "AD_index entry outside the class tables");
    const size_t nk = (size_t)std::max<int64_t>(mk, 1);
    DevBuf dk, da, df, dl, di, d0, d1, d2;
    if (buf_alloc(ctx, dk, nk * 4) || buf_alloc(ctx, da, nk * 4) || buf_alloc(ctx, df, (size_t)n_classes * 12) || buf_alloc(ctx, dl, (size_t)n_classes * 12) ||
        buf_alloc(ctx, di, (size_t)idx_rows * idx_cols * 4) || buf_alloc(ctx, d0, nk * 4) || buf_alloc(ctx, d1, nk * 4) || buf_alloc(ctx, d2, nk * 4)) return 1;
    CU(cudaMemcpyAsync(dk.p, L_keep, mk * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(da.p, A_vec, mk * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(df.p, AD_factorial, (size_t)n_classes * 12, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dl.p, AD_like, (size_t)n_classes * 12, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(di.p, AD_index, (size_t)idx_rows * idx_cols * 4, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH("zmoments", zmoments_list_kernel, grid_for(mk, 128, ctx->num_sm * 16), 128, 0, ctx->stream, ctx->G[0], ctx->AD, ctx->ldg,
           ctx->col_of_ind[ind], dk.as<int>(), (long)mk, da.as<float>(), df.as<float>(), dl.as<float>(), di.as<int>(), idx_rows, idx_cols,
           d0.as<float>(), d1.as<float>(), d2.as<float>());
    if (W_obs_out) CU(cudaMemcpyAsync(W_obs_out, d0.p, mk * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (W_l_out) CU(cudaMemcpyAsync(W_l_out, d1.p, mk * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (W_var_out) CU(cudaMemcpyAsync(W_var_out, d2.p, mk * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return 0;
}

int64_t wgs_zscore_deep_sites(const wgs_ctx* ctx) { return ctx->z_deep_sites; }

int32_t wgs_zscore_classes(wgs_ctx* ctx, int32_t ind, int32_t max_rows, int32_t* rows_out, int32_t* n_rows)
{
    if (ind < 0 || ind >= (int)ctx->zclasses.size()) return fail(ctx, "no class table for individual %d", ind);
    const auto& v = ctx->zclasses[ind];
    int n = (int)v.size() / 4;
    *n_rows = n;
    for (int r = 0; r < std::min(n, max_rows); ++r) memcpy(rows_out + 4 * r, v.data() + 4 * r, 4 * sizeof(int));
    return 0;
}

int32_t wgs_debug_stream(wgs_ctx* ctx, int32_t mode, double* ms_out, double* bytes_out)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    if (!ctx->G[0]) return fail(ctx, "no GL matrix resident");
    const long M = ctx->M();
    const int K = std::max(ctx->K, 1);
    DevBuf sink;
    if (buf_alloc(ctx, sink, 16)) return 1;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    dim3 grid = mode == 0 ? dim3(ctx->num_sm * 8) : dim3(std::max(1, ctx->num_sm * 8 / K), K);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a, ctx->stream);
        LAUNCH("probe", stream_probe_kernel, grid, 256, 0, ctx->stream, ctx->G[0], ctx->ldg, M, ctx->d_pops, K, mode, sink.as<float>());
        cudaEventRecord(b, ctx->stream);
    }
    CU(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    *ms_out = ms;
    double bytes = 0;
    if (mode == 0) bytes = (double)M * ctx->ldg * 8.0;
    else for (int k = 0; k < K; ++k) bytes += (double)M * ((ctx->pops[k].n + 1) / 2) * 16.0;
    *bytes_out = bytes;
    CU(cudaGetLastError());
    return 0;
}

int32_t wgs_debug_seqsum(wgs_ctx* ctx, const float* x, int64_t n, float carry_in, float* out)
{
    cudaSetDevice(ctx->device);
    DevBuf dx, dc, dout;
    if (buf_alloc(ctx, dx, (size_t)std::max<int64_t>(n, 1) * sizeof(float)) || buf_alloc(ctx, dc, sizeof(float)) || buf_alloc(ctx, dout, sizeof(float))) return 1;
    CU(cudaMemcpyAsync(dx.p, x, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dc.p, &carry_in, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH("em_resolve", seqsum_vec_kernel, 1, kSeqWarps * 32, 0, ctx->stream, dx.as<float>(), (long)n, dc.as<float>(), dout.as<float>());
    CU(cudaMemcpyAsync(out, dout.p, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return 0;
}

int64_t wgs_launch_count(const wgs_ctx* ctx) { return ctx->launches; }

int32_t wgs_partials_combined(const wgs_ctx* ctx) { return nccl_sums(ctx) ? 1 : 0; }

int32_t wgs_timing_reset(wgs_ctx* ctx, int32_t enable)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    fold_timing(ctx);
    ctx->tdone.clear();
    ctx->timing = enable != 0;
    return 0;
}

int32_t wgs_timing_get(wgs_ctx* ctx, const char* name, double* ms, int64_t* launches)
{
    cudaSetDevice(ctx->device);
    upload_fence(ctx);
    fold_timing(ctx);
    auto it = ctx->tdone.find(name);
    *ms = it == ctx->tdone.end() ? 0.0 : it->second.ms;
    *launches = it == ctx->tdone.end() ? 0 : it->second.launches;
    return 0;
}

int32_t wgs_timing_work(wgs_ctx* ctx, const char* name, double* bytes, double* units)
{
    auto it = ctx->tdone.find(name);
    *bytes = it == ctx->tdone.end() ? 0.0 : it->second.bytes;
    *units = it == ctx->tdone.end() ? 0.0 : it->second.units;
    return 0;
}

}  // extern "C"
