/* brr_gpu.h -- C ABI of the B200-native BayesR Gibbs-sweep engine.
 *
 * This is the drop-in boundary for the hot path of ctggroup/BayesRRcmd (reference paths below are
 * relative to the reference repository):
 *
 *   - sweep level (primary):  AnalysisGraph::exec(Analysis*, numInds, numSnps, markerIndices)
 *                             src/analysisgraph.hpp:18-21, called once per MCMC iteration from
 *                             BayesRBase::runGibbs, src/BayesRBase.cpp:209          -> brr_gpu_sweep
 *   - marker level (parity):  BayesRKernel::computeNum / calculateEpsilonChange
 *                             src/bayesrkernel.h:11-15                              -> brr_gpu_compute_num,
 *                                                                                      brr_gpu_apply_delta
 *   - column supply:          MarkerCache::populate, src/markercache.cpp:19-45      -> brr_gpu_upload_*
 *   - per-sweep scalars:      updateMu (src/DenseBayesRRmz.cpp:36-42,
 *                             src/SparseBayesRRG.cpp:106-113), ||eps||^2 for sigmaE
 *                             (src/BayesRBase.cpp:238)                              -> brr_gpu_eps_*
 *
 * Conventions: every function returns 0 on success, a negative brr_gpu_status otherwise;
 * brr_gpu_last_error() describes the last failure of the calling thread.  All pointers are HOST
 * pointers unless the name says `_dev`; buffers are caller-owned; the context owns all device
 * memory; a context is not thread-safe; no C++ exceptions cross this boundary; no torch types.
 * There is NO CPU fallback: without a CUDA device brr_gpu_create fails with BRR_GPU_ENODEVICE.
 *
 * A reference-side binding (what a maintainer would add to src/) is shown in INTEGRATION.md.
 */
#ifndef BRR_GPU_H
#define BRR_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BRR_GPU_ABI_VERSION 1

typedef enum {
    BRR_GPU_OK = 0,
    BRR_GPU_EINVAL = -1,     /* bad argument */
    BRR_GPU_ENODEVICE = -2,  /* no usable CUDA device / wrong architecture */
    BRR_GPU_ECUDA = -3,      /* a CUDA runtime call failed (see brr_gpu_last_error) */
    BRR_GPU_ESTATE = -4,     /* call sequence error (e.g. sweep before upload) */
    BRR_GPU_ENOMEM = -5
} brr_gpu_status;

typedef enum {
    BRR_GPU_LAYOUT_DENSE2BIT = 0, /* 2 bits per genotype; semantics of DenseMarker / DenseRKernel
                                     (src/densemarkerbuilder.cpp:23-60, src/densebayesrkernel.cpp) */
    BRR_GPU_LAYOUT_RAGGED = 1,    /* ascending index lists of ones / twos / missing; semantics of
                                     RaggedSparseMarker / RaggedBayesRKernel + SparseBayesRRG's
                                     analytically tracked sum(eps) (src/raggedbayesrkernel.cpp,
                                     src/sparsebayesrkernel.cpp, src/SparseBayesRRG.cpp:76-113) */
    BRR_GPU_LAYOUT_EIGEN = 2      /* the same index lists with the semantics of EigenSparseMarker /
                                     EigenBayesRKernel: missing genotypes imputed by the mean in the
                                     statistics and in the dot (src/eigensparsemarkerbuilder.cpp:40-78,
                                     src/eigenbayesrkernel.cpp:10-20) */
} brr_gpu_layout;

typedef struct brr_gpu_ctx brr_gpu_ctx;

typedef struct {
    uint32_t abi_version;     /* BRR_GPU_ABI_VERSION */
    int32_t device;           /* CUDA device ordinal */
    int32_t layout;           /* brr_gpu_layout */
    uint32_t n_inds;          /* individuals held by THIS context (rows of its shard) */
    uint32_t n_inds_total;    /* individuals over all shards (N in N-1; = n_inds when world == 1) */
    uint32_t n_markers;       /* M */
    uint32_t n_components;    /* K = cols(S)+1, 2..16 */
    uint32_t n_groups;        /* >= 1 */
    uint32_t block_markers;   /* markers sampled per grid-wide exchange (0 = default: 64 dense, 32 sparse; 1..64 dense, 1..32 sparse).  Any
                                 value gives the SAME chain as the reference's one-by-one loop: later
                                 markers of a block are corrected with the block's Gram matrix. */
    uint32_t max_ctas;        /* 0 = one CTA per SM */
    int32_t rank;             /* shard index, 0..world-1 */
    int32_t world;            /* number of shards (GPUs); individuals are split across them.  Dense 2-bit layout only:
                                 the ragged / eigen layouts have no cross-shard exchange and brr_gpu_create rejects
                                 world > 1 (or n_inds_total != n_inds) for them with BRR_GPU_EINVAL */
} brr_gpu_config;

/* Outputs of one sweep: what runGibbs needs after exec() (src/BayesRBase.cpp:238-247). */
typedef struct {
    double eps_sum;           /* sum(eps) after the sweep (ragged: the analytically tracked m_epsilonSum) */
    double eps_sqnorm;        /* ||eps||^2 after the sweep (this shard) */
    uint32_t n_nonskipped;    /* markers with beta_old != 0 || beta_new != 0 (src/BayesRBase.cpp:372) */
    uint32_t n_blocks;        /* grid-wide exchanges performed */
    float kernel_ms;          /* device time of the sweep kernel(s), CUDA events on the ctx stream */
    uint32_t n_launches;      /* kernels launched by this call */
} brr_gpu_sweep_out;

const char *brr_gpu_last_error(void);
int brr_gpu_abi_version(void);

int brr_gpu_create(const brr_gpu_config *cfg, brr_gpu_ctx **out);
void brr_gpu_destroy(brr_gpu_ctx *ctx);

/* ---- column supply (the device-resident MarkerCache) ---------------------------------------- */

/* PLINK SNP-major payload (no 3-byte magic): marker j's column starts at bed + j*col_stride and
 * holds ceil(n_inds/4) bytes for THIS shard's individuals (bit pairs LSB first; 00->2, 10->1,
 * 11->0, 01->missing: src/preprocessgraph.cpp:44-53).  Uploads markers [first, first+count).
 * Per-marker genotype counts are taken on the device. */
int brr_gpu_upload_dense2bit(brr_gpu_ctx *ctx, const uint8_t *bed, size_t col_stride, uint32_t first,
                             uint32_t count);

/* Genotype counts of this shard, M x 4 (n0, n1, n2, n_missing). */
int brr_gpu_get_counts(brr_gpu_ctx *ctx, uint32_t *counts);
/* Finalise the standardisation from GLOBAL counts (sum over shards; NULL = this shard's own):
 * dense: the 4-entry value table of DenseMarkerBuilder::endColumn (src/densemarkerbuilder.cpp:43-60);
 * ragged: mean, sd, Zsum of RaggedSparseMarkerBuilder::endColumn (src/raggedsparsemarkerbuilder.cpp:43-55). */
int brr_gpu_finalize_markers(brr_gpu_ctx *ctx, const uint32_t *global_counts);
/* Override / read back the dense value table, M x 4 doubles indexed by class 0,1,2,missing. */
int brr_gpu_set_dense_lut(brr_gpu_ctx *ctx, const double *lut);
int brr_gpu_get_dense_lut(brr_gpu_ctx *ctx, double *lut);

/* Ragged layout: offsets has 3*M+1 entries into idx (ones_0, twos_0, missing_0, ones_1, ...);
 * idx holds shard-local individual ids, ascending inside each list.  stats (M x 4: mean, sd, sqrdZ,
 * Zsum as in src/sparsemarker.h:8-11) may be NULL: then they are computed from the list lengths
 * by brr_gpu_finalize_markers. */
int brr_gpu_upload_ragged(brr_gpu_ctx *ctx, const uint64_t *offsets, const int32_t *idx, const double *stats);
/* Build the ragged layout on the host side of the ABI from a PLINK payload (convenience). */
int brr_gpu_upload_ragged_from_bed(brr_gpu_ctx *ctx, const uint8_t *bed, size_t col_stride);
int brr_gpu_get_ragged_stats(brr_gpu_ctx *ctx, double *stats);

/* ---- model and chain state ------------------------------------------------------------------ */

/* group_of_marker: M ints in 0..n_groups-1 or NULL (all 0, src/data.cpp:130);
 * cva: n_groups x (K-1) row-major (Options::S, src/options.hpp:37). */
int brr_gpu_set_model(brr_gpu_ctx *ctx, const int32_t *group_of_marker, const double *cva);

/* Any pointer may be NULL (leave unchanged / do not fetch). eps: n_inds; beta, components, acum: M. */
int brr_gpu_set_state(brr_gpu_ctx *ctx, const double *eps, const double *beta, const double *components,
                      const double *acum);
int brr_gpu_get_state(brr_gpu_ctx *ctx, double *eps, double *beta, double *components, double *acum);

/* eps += c (the two halves of updateMu); returns this shard's sum(eps) AFTER the add in *sum_out. */
int brr_gpu_eps_add_scalar(brr_gpu_ctx *ctx, double c, double *sum_out);
/* this shard's sum(eps) and ||eps||^2 */
int brr_gpu_eps_reduce(brr_gpu_ctx *ctx, double *sum_out, double *sqnorm_out);

/* ---- the sweep = AnalysisGraph::exec -------------------------------------------------------- */

/* marker_order: n_order distinct marker ids (a permutation when n_order == M);
 * rand_uz: M x 2 (u, z) indexed by MARKER ID (src/BayesRBase.cpp:343-344);
 * pi: n_groups x K; sigmaG: n_groups; eps_sum_in: m_epsilonSum (ragged layout only);
 * async_tokens: 0 = synchronous ppbayes; T > 0 = bounded-staleness asyncppbayes with T markers in
 * flight (Options::analysisTokens, src/options.hpp:33), T <= 64 for the dense layout, T <= 32 for the
 * ragged / eigen layouts (BRR_GPU_EINVAL beyond).
 * v (n_groups x K) and betasqn (n_groups) are OVERWRITTEN with this sweep's counts / sums
 * (runGibbs zeroes them before exec, src/BayesRBase.cpp:200-202).
 * Errors raised INSIDE the kernel (BRR_GPU_ESTATE: the residual grew past the fixed-point range of the
 * sweep, a peer stopped taking part, a protocol time-out; brr_gpu_last_error names which) leave the
 * device state of the chain (eps, beta, components, acum) undefined: the sweep updates it in place.
 * A caller that wants to go on restores the last good state with brr_gpu_set_state. */
int brr_gpu_sweep(brr_gpu_ctx *ctx, const uint32_t *marker_order, uint32_t n_order, const double *rand_uz,
                  const double *pi, const double *sigmaG, double sigmaE, double eps_sum_in,
                  uint32_t async_tokens, double *v, double *betasqn, brr_gpu_sweep_out *out);

/* The same sweep split in three so that a caller can keep the inputs resident in HBM and time the
 * kernel alone: stage (H2D), run (kernel only, asynchronous on the context stream), fetch (D2H). */
int brr_gpu_sweep_stage(brr_gpu_ctx *ctx, const uint32_t *marker_order, uint32_t n_order,
                        const double *rand_uz, const double *pi, const double *sigmaG, double sigmaE,
                        double eps_sum_in, uint32_t async_tokens);
int brr_gpu_sweep_run(brr_gpu_ctx *ctx);
int brr_gpu_sweep_fetch(brr_gpu_ctx *ctx, double *v, double *betasqn, brr_gpu_sweep_out *out);

/* ---- marker-level parity hooks = BayesRKernel ------------------------------------------------ */

/* computeNum(eps, beta_old) for one marker against the resident eps (this shard's part of the dot;
 * the beta_old*(N-1) and mean*eps_sum/sd terms are added once, by rank 0's caller, when world>1:
 * with world == 1 the result is the reference's computeNum). */
int brr_gpu_compute_num(brr_gpu_ctx *ctx, uint32_t marker, double beta_old, double eps_sum, double *num_out);
/* eps += calculateEpsilonChange(beta_old, beta_new); *eps_sum_delta (may be NULL) receives the
 * sparse kernels' epsilonSum delta (src/sparsebayesrkernel.cpp:29-34), 0 for dense. */
int brr_gpu_apply_delta(brr_gpu_ctx *ctx, uint32_t marker, double beta_old, double beta_new,
                        double *eps_sum_delta);

/* ---- multi-GPU: individuals sharded over ranks, per-block exchange over NVLink ---------------- */

#define BRR_GPU_IPC_HANDLE_BYTES 64
/* Export this context's exchange mailbox; gather the handles of all ranks (any transport), then
 * attach.  After attach, brr_gpu_sweep on every rank (called collectively, same arguments except
 * the shard-local ones) exchanges the per-block integer dot / Gram counts inside the kernel by
 * peer stores + flags over NVLink; every rank then performs the identical scalar sampling, so beta,
 * components and acum are replicated and eps stays sharded.  The sums are integers: the chain is
 * bit-identical for any number of ranks.  A rank that stops participating makes the others fail
 * with BRR_GPU_ESTATE after a few seconds instead of hanging. */
int brr_gpu_ipc_export(brr_gpu_ctx *ctx, void *handle_out);
int brr_gpu_ipc_attach(brr_gpu_ctx *ctx, const void *handles_all_ranks);
/* The same wiring for several contexts living in ONE process (rank i = ctxs[i], world = n): raw peer
 * pointers instead of IPC handles.  With both contexts on one device this also serves as a single-GPU test
 * of the exchange protocol (the two kernels must then be co-resident: keep n_inds small). */
int brr_gpu_attach_local(brr_gpu_ctx **ctxs, int n);
uint32_t brr_gpu_num_blocks(brr_gpu_ctx *ctx);
/* The context's CUDA stream (cudaStream_t), so the external allreduce can be ordered on it. */
void *brr_gpu_stream(brr_gpu_ctx *ctx);

/* ---- synthetic workloads, generated on the device (bench / scale tests) ---------------------- */

/* Fill the dense store with iid Binomial(2, maf_j) genotypes, maf_j ~ U(maf_lo, maf_hi) (log-uniform
 * if log_uniform != 0), missing with probability p_missing; counter-based RNG keyed by
 * (seed, marker, GLOBAL individual id) so that shards of one dataset agree. */
int brr_gpu_synth_dense(brr_gpu_ctx *ctx, uint64_t seed, double maf_lo, double maf_hi, int log_uniform,
                        double p_missing, uint32_t ind_offset);
/* The same genotypes straight into the ragged / eigen store (index lists built on the device). */
int brr_gpu_synth_ragged(brr_gpu_ctx *ctx, uint64_t seed, double maf_lo, double maf_hi, int log_uniform,
                         double p_missing, uint32_t ind_offset);
/* The same genotypes as a host PLINK payload for markers [first, first+count) (for CPU baselines). */
int brr_gpu_synth_bed_host(uint64_t seed, double maf_lo, double maf_hi, int log_uniform, double p_missing,
                           uint32_t n_inds, uint32_t ind_offset, uint32_t first, uint32_t count, uint8_t *bed_out);

/* introspection */
int brr_gpu_info(brr_gpu_ctx *ctx, uint32_t *n_ctas, uint32_t *slice_inds, uint32_t *block_markers,
                 uint64_t *genotype_bytes);
uint64_t brr_gpu_launch_count(brr_gpu_ctx *ctx);
/* Version tag of the dense sweep kernel this library was built with (bench lines and ncu captures carry it). */
const char *brr_gpu_kernel_version(void);
/* Phase breakdown of the dense sweep kernel (SM cycles of CTA 0, accumulated over sweeps), available when
 * the environment has BRR_GPU_PROFILE=1 at brr_gpu_create: 40 counters in the order wait (grid barrier),
 * cross-GPU exchange, read-back, sampling, join (walkers waiting for the dot warps), publish (next block's
 * round), Gram-row cross counts, Gram-row products, Gram-row barrier, Gram-row read-back, triangular solve,
 * re-sampling, finalize, carry, axpy, quantize, tail (prefetch), prologue; then 4 counts (walk iterations,
 * Gram rounds, Gram rows, blocks that updated the residual); then the dot warps' wait for the code rows and
 * their dot pass (these two overlap the walker phases); then finer marks inside publish (reduce, missing sums,
 * arrive, preamble) and the Gram-row round (loop, reduction, three marks of the cross counts), and 4 spare.  reset != 0 clears them. */
int brr_gpu_profile_cycles(brr_gpu_ctx *ctx, uint64_t *cycles40, int reset);
/* The same for the version-5 dense kernel (sweep_dense5.cuh), 24 counters: walker warp 0 of CTA 0: wait for the plan, wait for
 * the block's counter, read-back (+ Gram rows), sampling, in-block corrections, filing; first dot warp: block-start gate,
 * residual update, wait for the rows, MMA loop, missing sums + cross counts, combine + REDs; then 4 counts (blocks walked with
 * Gram rows, -, restarts, residual updates); planner warp: plan, TMA issue. */
int brr_gpu_profile_cycles5(brr_gpu_ctx *ctx, uint64_t *cycles24, int reset);
/* The same for the version-6 dense kernel (sweep_dense6.cuh), 32 counters, one warp of CTA 0 per role: first dot warp (wait for the
 * decision, residual update, wait for the rows, MMA loop, accumulate), post warp (wait, work), publisher (wait, work), first
 * producer (wait for the plan, wait for a free stage, issue), control warp (result poll, plan, decide), walker warp 0 (wait for the
 * slot counter, read-back, rows of earlier blocks, sampling, in-block corrections + result line, validation); then counts:
 * -, blocks walked by CTA 0, residual updates, -, blocks walked with an in-model row. */
int brr_gpu_profile_cycles6(brr_gpu_ctx *ctx, uint64_t *cycles32, int reset);
/* Version-6 kernel, per CTA (the CTAs wait for each other through the block results, so the slowest one paces the sweep): 8 words per
 * CTA, the first dot warp's cycles in {wait for the decision, residual update + bookkeeping, wait for the rows, MMA loop, wait for the
 * accumulator set, store + arrive}, the blocks it went through, the SM it ran on.  words: [8 * max_ctas]; *n_ctas = CTAs written. */
int brr_gpu_profile_ctas6(brr_gpu_ctx *ctx, uint64_t *words, uint32_t max_ctas, uint32_t *n_ctas, int reset);
/* Which dense kernel the last sweep ran (4 = on-demand Gram rows, 5 = steady-state pipeline, 6 = steady-state pipeline with the
 * walk distributed over the CTAs; 0 = not dense) and, for versions 5 / 6, how often the pipeline restarted behind a marker that
 * entered the model.  The library chooses per sweep: version 6 (one GPU or several) when the sweep is synchronous and the previous
 * sweep of the chain had <= 3% non-skipped markers, version 4 otherwise (chain start, asynchronous mode).  Environment:
 * BRR_GPU_DENSE_KERNEL=v4|v5|v6 forces a kernel, BRR_GPU_V6_DEPTH=1..7 the blocks version 6 keeps in flight (default 4 on one GPU,
 * 7 across GPUs), BRR_GPU_PROFILE=1 enables the phase counters, BRR_GPU_V6_DBG=1 prints where the warps waited after a protocol
 * time-out (builds with -DBRR_V6_DEBUG). */
int brr_gpu_last_sweep_info(brr_gpu_ctx *ctx, uint32_t *kernel_version, uint32_t *restarts);

#ifdef __cplusplus
}
#endif
#endif /* BRR_GPU_H */
