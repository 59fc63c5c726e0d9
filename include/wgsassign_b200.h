/*
 * wgsassign_b200 - C ABI of the B200-native WGSassign genotype-likelihood hot path.
 *
 * This header is the drop-in boundary.  The reference has no FFI of
 * its own: its native layer is a set of Cython `cpdef`s that its Python drivers call once
 * per EM iteration / per (individual, population).  That granularity cannot be kept on a
 * GPU, so each entry point below replaces one reference *driver* (Python function + the
 * Cython kernels under it) and is what a ctypes binding in the reference's modules would
 * call.  Citations are file:line in the reference checkout (WGSassign/...).
 *
 * Conventions
 *   - plain C: pointers and sizes only; every call returns 0 on success, non-zero on
 *     error, and wgs_last_error() then describes it (no exceptions cross the ABI);
 *   - all array arguments are HOST pointers, C-contiguous, in the reference's layouts:
 *       L   float32 [M, 2N]  L[s,2i]=P(G=0), L[s,2i+1]=P(G=1)      (reader_cy.pyx:16-77)
 *       A   float32 [M, K]   columns in np.unique(pop names) order  (WGSassign.py:213-243)
 *       AD  int32   [M, 2N]  (ref/major, alt/minor) read counts      (WGSassign.py:320)
 *     the pointer is borrowed for the duration of the call only; pinned host memory
 *     (wgs_host_alloc) is copied at full PCIe rate, pageable memory works too;
 *   - outputs are caller-allocated; calls are synchronous (return when outputs are valid);
 *   - there is NO CPU fallback: without a CUDA device wgs_create fails.
 *
 * Site sharding (one process per GPU): each rank creates a context on its device, uploads
 * its contiguous site range, and calls wgs_set_shard() with the global site count and an
 * all-reduce callback; per-site outputs are then local slices and the small
 * per-individual/per-population sums are returned as float64 partials ("_partial") that
 * the caller combines in fixed rank order.
 */
#ifndef WGSASSIGN_B200_H
#define WGSASSIGN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wgs_ctx wgs_ctx;

#define WGS_ABI_VERSION 1

/* dtype codes for the all-reduce callback */
#define WGS_F64 0
#define WGS_I64 1

/* Sum `n` elements of `buf` (dtype WGS_F64 / WGS_I64) across all ranks, in place, host memory. */
typedef void (*wgs_allreduce_fn)(void *buf, int64_t n, int32_t dtype, void *user);

int32_t     wgs_abi_version(void);
int32_t     wgs_device_count(void);
/* PCI bus id of a device ("0000:1b:00.0", as sysfs names it): the host side uses it to allocate its pinned staging
 * buffers on the NUMA node the GPU hangs off (dist.bind_near_gpu).  len >= 16. */
int32_t     wgs_device_pci_bus_id(int32_t device, char *out, int32_t len);
const char *wgs_last_error(const wgs_ctx *ctx); /* ctx may be NULL: last create() error */

int32_t wgs_create(int32_t device, wgs_ctx **out);
void    wgs_destroy(wgs_ctx *ctx);

/* Run-time switches (fallback kernels, experiments, numerics variants such as "z_exact_means"): the library never
 * reads environment variables on a call path - a result cannot depend on the ambient environment.  Unknown names
 * are an error.  (With WGS_DEBUG set in the environment, wgs_create seeds the table once from WGS_<NAME>=<int>.)
 * The names and their meaning are listed in csrc/wgs_api.cu (kOptionNames). */
int32_t wgs_set_option(wgs_ctx *ctx, const char *name, int32_t value);
int32_t wgs_get_option(const wgs_ctx *ctx, const char *name, int32_t *value);

/* Pinned host buffers for full-rate H2D (optional). */
void *wgs_host_alloc(int64_t bytes);
void  wgs_host_free(void *p);

/* ---- data residency --------------------------------------------------------------------- */

/* Population structure of the N individuals (Beagle column order): pop_of_ind[i] in [0,K) is
 * the index of individual i's population in np.unique order (WGSassign.py:211-213,
 * glassy.py:63-67).  Must be called before wgs_upload_gl for the reference-panel ops
 * (ref_af / loo / fisher / reference z-score); not needed for pop_like.  K may be 0 to clear. */
int32_t wgs_set_pops(wgs_ctx *ctx, const int32_t *pop_of_ind, int32_t N, int32_t K);

/* Repack L [M,2N] into the device layout (population-sorted (g0,g1) pairs, 32-byte aligned
 * population slabs; g2 is never stored, it is recomputed as 1-g0-g1 like the reference
 * does).  which=0: the GL matrix; which=1: the optional down-sampled matrix that
 * glassy.loo scores against (glassy.py:96). */
int32_t wgs_upload_gl(wgs_ctx *ctx, const float *L, int64_t M, int32_t N, int32_t which);

/* Asynchronous form of wgs_upload_gl(which=0): queues one strided DMA per population slab (straight
 * into the device layout, no staging pass) on a copy stream and returns at once.  L must stay valid
 * and unchanged until wgs_upload_wait() or the next operator call returns; pinned memory
 * (wgs_host_alloc) is needed for the copy to overlap anything.  Every operator waits for the whole
 * matrix first, except wgs_ref_af_loo, which starts on population 0 while the others are in flight.
 * ID files that interleave populations column by column fall back to the synchronous upload. */
int32_t wgs_upload_gl_async(wgs_ctx *ctx, const float *L, int64_t M, int32_t N);
int32_t wgs_upload_wait(wgs_ctx *ctx);

/* The same upload for a matrix that is still being parsed (the streaming Beagle reader): begin allocates for up to
 * M_capacity rows, rows() queues the DMAs of a finished row block straight into the device layout and returns at once
 * (the block must stay valid and unchanged until wgs_upload_wait or the next operator returns), end() fixes the row
 * count.  Parsing block b+1 overlaps the transfer of block b. */
int32_t wgs_upload_gl_begin(wgs_ctx *ctx, int64_t M_capacity, int32_t N);
int32_t wgs_upload_gl_rows(wgs_ctx *ctx, const float *L_rows, int64_t row0, int64_t nrows);
int32_t wgs_upload_gl_end(wgs_ctx *ctx, int64_t M_final);

/* Allele depths [M,2N] int32, packed on device to 2 x uint8 per individual.  A count above 254 is stored as the
 * sentinel "deeper than any class table": the site joins the deep sites (wgs_zscore_deep_sites) and is never kept,
 * exactly what the reference does with a depth whose depth+1 splits were not all observed (zscore.py:36-39).
 * Negative counts are an error. */
int32_t wgs_upload_ad(wgs_ctx *ctx, const int32_t *AD, int64_t M, int32_t N);
/* The same from saturating uint8 pairs [M,2N] (255 = "255 reads or more"), the form wgs_ad_stream_next_u8 parses into:
 * a quarter of the host memory and PCIe bytes of the int32 matrix. */
int32_t wgs_upload_ad_u8(wgs_ctx *ctx, const uint8_t *AD, int64_t M, int32_t N);

/* Multi-GPU site sharding: global number of sites, global index of this shard's first site,
 * and the cross-rank sum used for the EM stop rule and the z-score class tallies. */
int32_t wgs_set_shard(wgs_ctx *ctx, int64_t M_total, int64_t site_offset,
                      wgs_allreduce_fn fn, void *user);

/* Position of this context in the site order of the ranks (rank r holds the r-th contiguous site range).  Needed
 * by the operators whose arithmetic is ORDER dependent in the reference - the sequential float32 class means of
 * the z-score (zscore.py:22) and the sequential float32 stop-rule sum (emMAF_cy.pyx:26-33) - which are chained
 * from rank to rank.  wgs_nccl_init sets it as well. */
int32_t wgs_set_rank(wgs_ctx *ctx, int32_t rank, int32_t world);

/* 1 when the operators below return sums that are ALREADY combined over the ranks (a communicator is attached:
 * all-gather over NVLink + rank-ordered sum on the device), 0 when they are local partial sums that the caller
 * must add in rank order (no communicator: single GPU, or the host callback only). */
int32_t wgs_partials_combined(const wgs_ctx *ctx);

/* Optional NCCL communicator for the sharded EM stop rule.  Without it the per-iteration sums of squared
 * changes pass through `fn` on the host (one round trip per iteration); with it they are all-gathered on the
 * compute stream (NVLink) and added in rank order on the device, so a sharded EM runs with the same
 * look-ahead launch schedule as a single GPU.  libnccl.so.2 is bound at run time (dlopen).  Rank 0 calls
 * wgs_nccl_unique_id (128 bytes out) and distributes the id; every rank then calls wgs_nccl_init (collective).
 * `fn` is still required (class tallies and final sums use it). */
int32_t wgs_nccl_unique_id(void *id_out128);
int32_t wgs_nccl_init(wgs_ctx *ctx, const void *id128, int32_t rank, int32_t world);

/* Device-side seeded synthetic data of the SURVEY 8d model (benchmarks): fills the resident
 * GL (and AD when with_ad) for M sites x N individuals; requires wgs_set_pops first. */
int32_t wgs_synth(wgs_ctx *ctx, int64_t M, int32_t N, uint64_t seed, float depth, int32_t with_ad);
/* Copy a slice of the resident data back in the reference layouts (any pointer may be NULL). */
int32_t wgs_download(wgs_ctx *ctx, int64_t site0, int64_t nsites, float *L_out, int32_t *AD_out);

/* ---- operators -------------------------------------------------------------------------- */

/* emMAF.emMAF(L_pop, iter, tole, t) (emMAF.py:15-27 over emMAF_cy.pyx:10-33): EM allele
 * frequency of ONE group given its own [M,2n] matrix; f_out [M]; *iters_out = iteration it
 * converged at (1-based) or 0 if it did not within `iter`.  No clipping. Self-contained. */
int32_t wgs_emMAF(wgs_ctx *ctx, const float *L_pop, int64_t M, int32_t n, int32_t iter, double tole,
                  float *f_out, int32_t *iters_out);

/* Per-population EM + clipping (WGSassign.py:225-242): af_out [M,K], iters_out [K].  The result
 * also stays resident on the device; af_out may be NULL when only wgs_loo_partial will use it. */
int32_t wgs_ref_af(wgs_ctx *ctx, int32_t iter, double tole, float *af_out, int32_t *iters_out);

/* glassy.assignLL (glassy.py:18-44 over glassy_cy.pyx:12-21): float64 sums over this
 * context's sites, out [N,K]; the caller rounds to float32 after combining shards. */
int32_t wgs_pop_like_partial(wgs_ctx *ctx, const float *af, int32_t K, double *out);

/* glassy.loo (glassy.py:47-112): leave-one-out EM for every individual, then the
 * log-likelihood of each individual under every population with the reference's
 * in-place column overwrite order (glassy.py:89).  af_inout [M,K] is read (full-data AF)
 * and left as the reference leaves it.  ll [N,K] float64; ll_parts [N*parts,K] float64
 * (may be NULL when parts==1); iters_out [N].  use_ds: score against the which=1 matrix.
 * af_inout may be NULL: the matrix left on the device by wgs_ref_af is used and updated in place. */
int32_t wgs_loo_partial(wgs_ctx *ctx, float *af_inout, int32_t iter, double tole, int32_t use_ds,
                        int32_t parts, double *ll, double *ll_parts, int32_t *iters_out);

/* `--get_reference_af --loo` in one call (WGSassign.py:225-242 followed by glassy.py:47-112): the
 * same results as wgs_ref_af(af_out, af_iters_out) followed by wgs_loo_partial(NULL, ...), bit for
 * bit.  After wgs_upload_gl_async the leave-one-out EM of population k runs while the slabs of the
 * later populations are still crossing PCIe (it does not need the full-data frequencies); the
 * per-population EM and the likelihood pass follow when the matrix is complete.  af_out [M,K] (the
 * matrix the CLI saves as .pop_af.npy) and af_after_loo [M,K] (the state glassy.loo leaves in its
 * `af` argument) may each be NULL. */
int32_t wgs_ref_af_loo(wgs_ctx *ctx, int32_t iter, double tole, float *af_out, int32_t *af_iters_out,
                       float *af_after_loo, int32_t use_ds, int32_t parts, double *ll, double *ll_parts,
                       int32_t *loo_iters_out);

/* fisher.fisher_obs + fisher.fisher_obs_ind in one pass (fisher.py:11-59 over
 * fisher_cy.pyx:12-65): f_obs, ne_obs [M,K] float32 (local sites); ne_ind_sum [N] float64 =
 * sum over local sites of the per-individual n-tilde (caller divides by the global M). */
int32_t wgs_fisher_partial(wgs_ctx *ctx, const float *af, float *f_obs, float *ne_obs,
                           double *ne_ind_sum);

/* z-score (WGSassign.py:311-446 over zscore.py:11-120 and zscore_cy.pyx:10-56). */
typedef struct {
    float   z;         /* (w_obs - z_mu) / sqrt(z_var), float32 like the reference */
    float   w_obs;     /* sum of observed per-site log-likelihoods */
    float   z_mu;      /* sum of expected per-site log-likelihoods */
    float   z_var;     /* sum of per-site variances */
    int64_t loci_kept; /* len(L_keep) - bit-exact */
    int32_t n_classes; /* rows of AD_array - bit-exact */
    int32_t em_iters;  /* reference mode: LOO EM iterations on the kept sites */
} wgs_zrow;

/* mode 0: --get_assignment_z_score (af [M,K] + pop_of_ind give each individual's column);
 * mode 1: --get_reference_z_score (LOO EM on the kept sites, af ignored);
 * mode 2: preparation only (class tallies + kept-site counts; z fields are NaN).
 * Individuals ind_start <= i < ind_end; out[ind_end-ind_start]. */
int32_t wgs_zscore(wgs_ctx *ctx, int32_t mode, const float *af, int32_t K, int32_t n_threshold,
                   int32_t single_read, int32_t ind_start, int32_t ind_end, int32_t iter,
                   double tole, wgs_zrow *out);

/* Per-individual allele-depth class table (the reference's AD_array, zscore.py:39) of the last
 * wgs_zscore call, for tally parity: rows of (ref, alt, depth, n_loci) ordered by (depth, alt).
 * The reference's row order is first occurrence in the file; no output depends on it. */
int32_t wgs_zscore_classes(wgs_ctx *ctx, int32_t ind, int32_t max_rows, int32_t *rows_out,
                           int32_t *n_rows);

/* Finer-grained z-score entry points behind the reference's per-individual functions:
 *  - wgs_zscore(mode 2, ...) runs the preparation only (tallies, class decisions, kept counts);
 *  - wgs_zscore_table: every observed class of individual `ind` from the last wgs_zscore call,
 *    rows of 7 floats (ref, alt, n_loci, mean GL0, GL1, GL2, kept flag): zscore.AD_summary's dict;
 *  - wgs_zkeep_one: zscore.get_L_keep (zscore.py:43-61) for the caller's AD_array [C,4] and class
 *    means [C,3]; writes the kept site indices (ascending) and their number;
 *  - wgs_zmoments_list: zscore_cy.expected_W_l + variance_W_l (zscore_cy.pyx:10-56) over the
 *    caller's kept-site list, AF vector and class tables; per-site float32 outputs [mk]. */
int32_t wgs_zscore_table(wgs_ctx *ctx, int32_t ind, int32_t max_rows, float *rows_out, int32_t *n_rows);
int32_t wgs_zkeep_one(wgs_ctx *ctx, int32_t ind, int32_t n_classes, const int32_t *ad_array,
                      const float *class_means, int32_t *keep_out, int64_t cap, int64_t *n_kept);
int32_t wgs_zmoments_list(wgs_ctx *ctx, int32_t ind, const int32_t *L_keep, int64_t mk, const float *A_vec,
                          int32_t n_classes, const float *AD_factorial, const float *AD_like,
                          const int32_t *AD_index, int32_t idx_rows, int32_t idx_cols,
                          float *W_obs_out, float *W_l_out, float *W_var_out);

/* Number of (site, individual) pairs of the last wgs_zscore call whose read depth exceeded
 * the dense class table (40 reads); such sites are treated as "class not kept", which is what
 * the reference does unless every one of the depth+1 splits of that depth was observed. */
int64_t wgs_zscore_deep_sites(const wgs_ctx *ctx);

/* ---- host-side readers of the text inputs ---------------------------------------------------------- */
/* Beagle genotype likelihoods (reader_cy.pyx:16-77): gzip text -> float32 [M, 2N], equal to the reference reader's
 * matrix bit for bit ((float)atof of the first two GLs of each triple), plus sample and site names.
 * Streaming form: a background thread inflates while a persistent pool of `threads` parser threads (<= 0: all cores)
 * converts the rows of the current block STRAIGHT INTO THE CALLER'S BUFFER (pinned memory in the CLI: no intermediate
 * copy), so a block can be uploaded while the next one is parsed.  wgs_beagle_stream_keep restricts conversion to a
 * row range (site-sharded ranks keep only their own rows; the rest is counted and named, never converted);
 * wgs_beagle_stream_next returns the number of rows written (0 at end of file, -1 on error). */
typedef struct wgs_beagle_stream wgs_beagle_stream;
int32_t     wgs_beagle_stream_open(const char *path, int32_t threads, wgs_beagle_stream **out);
int32_t     wgs_beagle_stream_inds(const wgs_beagle_stream *s);
const char *wgs_beagle_stream_sample(const wgs_beagle_stream *s, int32_t i);
int32_t     wgs_beagle_stream_keep(wgs_beagle_stream *s, int64_t row_lo, int64_t row_hi);  /* row_hi < 0: to the end */
int32_t     wgs_beagle_stream_names(wgs_beagle_stream *s, int32_t keep);   /* 0: do not collect site names (and stop at row_hi) */
int64_t     wgs_beagle_stream_next(wgs_beagle_stream *s, float *out, int64_t max_rows);
int64_t     wgs_beagle_stream_rows_seen(const wgs_beagle_stream *s);
const char *wgs_beagle_stream_site(const wgs_beagle_stream *s, int64_t row);               /* rows seen so far */
/* every site name seen so far, each followed by '\n', into out[0 .. cap); returns the bytes needed (out may be NULL) */
int64_t     wgs_beagle_stream_sites_joined(const wgs_beagle_stream *s, char *out, int64_t cap);
int64_t     wgs_beagle_stream_estimate_rows(const wgs_beagle_stream *s);                   /* total rows, from the bytes read so far */
/* 1 when a Beagle / allele-depth stream reads a BGZF file (bgzip, what ANGSD writes: independent members of <= 64 KB
 * of text, inflated by several threads at once), 0 for a plain gzip stream (one inflate thread). */
/* The part-th of `parts` equal byte ranges of a BGZF Beagle file: the rows that START inside it (every row of the file
 * belongs to exactly one part; the parts in order are the file).  Processes reading different parts inflate disjoint
 * shares of the file - what a site-sharded job wants instead of every rank inflating all of it.  Row numbers count from
 * the part's first row; the sample names are those of the file's header.  An error for a plain gzip file. */
int32_t     wgs_beagle_stream_open_part(const char *path, int32_t threads, int32_t part, int32_t parts, wgs_beagle_stream **out);
int32_t     wgs_stream_is_bgzf(const void *stream);
int32_t     wgs_beagle_stream_stats(const wgs_beagle_stream *s, double *inflate_s, double *parse_s,
                                    int64_t *compressed_bytes, int64_t *uncompressed_bytes);
void        wgs_beagle_stream_close(wgs_beagle_stream *s);

/* Allele depths of the z-score modes (WGSassign.py:320, :399: np.loadtxt of an [M, 2N] integer text matrix, plain or
 * gzipped): the same pipeline, converting to saturating uint8 pairs (255 = "255 reads or more", the device layout of
 * wgs_upload_ad_u8) or to the reference's int32. */
typedef struct wgs_ad_stream wgs_ad_stream;
int32_t wgs_ad_stream_open(const char *path, int32_t threads, wgs_ad_stream **out);
int32_t wgs_ad_stream_inds(const wgs_ad_stream *s);
int32_t wgs_ad_stream_keep(wgs_ad_stream *s, int64_t row_lo, int64_t row_hi);
int64_t wgs_ad_stream_next_u8(wgs_ad_stream *s, uint8_t *out, int64_t max_rows);
int64_t wgs_ad_stream_next_i32(wgs_ad_stream *s, int32_t *out, int64_t max_rows);
int64_t wgs_ad_stream_rows_seen(const wgs_ad_stream *s);
int64_t wgs_ad_stream_estimate_rows(const wgs_ad_stream *s);
int32_t wgs_ad_stream_stats(const wgs_ad_stream *s, double *inflate_s, double *parse_s, int64_t *compressed_bytes,
                            int64_t *uncompressed_bytes);
void    wgs_ad_stream_close(wgs_ad_stream *s);

/* Whole-file form of the Beagle reader (one call, like reader_cy.readBeagle). */
typedef struct wgs_beagle wgs_beagle;
int32_t     wgs_beagle_open(const char *path, int32_t threads, wgs_beagle **out);
const char *wgs_beagle_last_error(void);
int64_t     wgs_beagle_sites(const wgs_beagle *b);
int32_t     wgs_beagle_inds(const wgs_beagle *b);
const char *wgs_beagle_sample(const wgs_beagle *b, int32_t i);
const char *wgs_beagle_site(const wgs_beagle *b, int64_t s);
int32_t     wgs_beagle_copy(const wgs_beagle *b, float *L_out); /* [sites, 2*inds] */
void        wgs_beagle_close(wgs_beagle *b);

/* Diagnostic read stream over the resident GL matrix (no arithmetic): mode 0 = flat 128-bit
 * read of the whole matrix, mode 1 = one population slab at a time (the access pattern of the
 * per-population kernels).  Returns the device time of one pass and the bytes it read. */
int32_t wgs_debug_stream(wgs_ctx *ctx, int32_t mode, double *ms_out, double *bytes_out);

/* Diagnostic: the device's order-exact float32 summation (the primitive behind the EM stop rule, rmse1d of
 * emMAF_cy.pyx:26-33): *out = ((carry_in + x[0]) + x[1]) + ... in float32, computed with the blocked integer
 * scheme of block_seqsum32 (csrc/wgs_kernels.cuh).  Tests compare it bit for bit with a serial loop. */
int32_t wgs_debug_seqsum(wgs_ctx *ctx, const float *x, int64_t n, float carry_in, float *out);

/* ---- instrumentation --------------------------------------------------------------------- */
/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
int64_t wgs_launch_count(const wgs_ctx *ctx);
/* Device time (ms, CUDA events on the context's stream) and launch count of the named kernel
 * family accumulated since wgs_timing_reset: "pop_like", "em_pop", "loo_em", "loo_like",
 * "fisher", "repack", "ztally", "zkeep", "zmoments". */
int32_t wgs_timing_reset(wgs_ctx *ctx, int32_t enable);
int32_t wgs_timing_get(wgs_ctx *ctx, const char *name, double *ms, int64_t *launches);
/* Algorithmic work of the same kernel family since wgs_timing_reset: bytes that must cross
 * HBM (inputs read once + outputs written once per launch) and work units (likelihood or
 * posterior evaluations); DESIGN.md states the per-unit figures. */
int32_t wgs_timing_work(wgs_ctx *ctx, const char *name, double *bytes, double *units);

#ifdef __cplusplus
}
#endif
#endif /* WGSASSIGN_B200_H */
