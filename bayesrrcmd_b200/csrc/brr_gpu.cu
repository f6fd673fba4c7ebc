// C-ABI implementation (include/brr_gpu.h): context, device-resident marker store, sweep launcher.
// There is no CPU fallback anywhere in this file: every entry point either runs CUDA kernels on
// the context's device or fails.
#include "../../include/brr_gpu.h"

#include "common.cuh"
#include "sampler.cuh"
#include "store_kernels.cuh"
#include "sweep_dense.cuh"
#include "sweep_dense5.cuh"
#include "sweep_dense6.cuh"
#include "sweep_ragged.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

using namespace brr;

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(BRR_GPU_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct brr_gpu_ctx {
    brr_gpu_config cfg{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int numSMs = 0;
    uint32_t nCTA = 0, SL = 0, PW = 0, RS = 0, B = 0;
    size_t smem_bytes = 0;

    // dense store: 2-bit codes, [M][nCTA][PW] words
    uint32_t *codes = nullptr;
    size_t codes_bytes = 0;
    double *lut = nullptr;
    uint32_t *counts_local = nullptr, *counts_global = nullptr;
    // ragged store
    RaggedStore rag{};
    // model
    int32_t *group = nullptr;
    double *cva = nullptr;
    // chain state
    double *eps = nullptr, *beta = nullptr, *comp = nullptr, *acum = nullptr;
    // sweep inputs
    uint32_t *order = nullptr;
    double *rand_uz = nullptr, *pi = nullptr, *sigmaG = nullptr;
    double sigmaE = 0, eps_sum_in = 0;
    uint32_t n_order = 0, async_tokens = 0;
    // scratch + outputs
    double *part_dot = nullptr, *cta_red = nullptr, *red_partial = nullptr;
    int *slots = nullptr;                   // dense sweep: rotating reduction slots + absmax + barrier, one allocation
    unsigned long long *absmax = nullptr;
    unsigned *bar = nullptr;
    double *v = nullptr, *betasq = nullptr, *out_scalars = nullptr;
    unsigned *out_counts = nullptr;
    unsigned long long *prof = nullptr;     // phase cycle counters (BRR_GPU_PROFILE=1)
    // version-5 dense kernel (single GPU, synchronous): ring of per-block slots, per-CTA partial outputs
    int *slots5 = nullptr;
    double *part_bsq = nullptr;
    int *part_v = nullptr;
    unsigned long long *prof5 = nullptr;
    size_t smem5_bytes = 0;
    bool v5_ok = false;
    // version-6 dense kernel (walk distributed over the CTAs): result lines, per-CTA constant tables, in-model mask of the sweep's positions
    int *results6 = nullptr;
    GroupConsts *gc6 = nullptr;
    uint32_t *inmask6 = nullptr;
    unsigned long long *prof6 = nullptr;
    unsigned *dbg6 = nullptr;
    uint32_t v6_launches = 0;               // version-6 launches so far: the tag of the cross-GPU cells
    size_t smem6_bytes = 0;
    bool v6_ok = false;
    int dense_mode = 0;                     // 0 = auto (per sweep), 4 / 5 / 6 = forced (BRR_GPU_DENSE_KERNEL)
    bool have_last = false;                 // the previous sweep's non-skipped count is known and the chain state was not replaced since
    double last_nonskipped_frac = 1.0;
    uint32_t last_kernel = 0, last_restarts = 0;
    // multi-GPU exchange (world > 1)
    int *xchg = nullptr;                    // this rank's exchange buffer (mailboxes + flags), cudaMalloc'd, IPC-exported
    int **peer_ptrs_dev = nullptr;          // device array [world] of every rank's exchange buffer
    std::vector<void *> peer_opened;        // IPC mappings to close
    unsigned *bar2 = nullptr;
    unsigned long long *round_ctr = nullptr; // device: exchange rounds completed so far (identical on all ranks)
    bool attached = false;
    uint32_t xchg_ctas = 1;
    // pinned host mirrors for the small outputs
    double *h_small = nullptr;   // [kMaxGK + kMaxGK + 2 + 2*1024]
    unsigned *h_counts = nullptr;

    std::vector<uint32_t> order_stamp;      // brr_gpu_sweep_stage: marker j was seen in call number order_stamp[j]
    uint32_t order_gen = 0;
    bool uploaded = false, finalized = false, model_set = false, staged = false, ran = false;
    uint64_t launches = 0;
    uint32_t last_launches = 0;
    float last_ms = 0;
};

static int ragged_compute_num(brr_gpu_ctx *, uint32_t, double, double, double *);
static int ragged_apply_delta(brr_gpu_ctx *, uint32_t, double, double, double *);

static const int kRedBlocks = 592;   // 4 x 148: blocks of the plain eps reductions

extern "C" const char *brr_gpu_last_error(void) { return g_err.c_str(); }
extern "C" int brr_gpu_abi_version(void) { return BRR_GPU_ABI_VERSION; }

template <class T>
static int dev_alloc(T **p, size_t n)
{
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(p), std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return fail(BRR_GPU_ENOMEM, "cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
    return 0;
}

extern "C" void brr_gpu_destroy(brr_gpu_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    void *ptrs[] = {c->codes, c->lut, c->counts_local, c->counts_global, c->group, c->cva, c->eps, c->beta, c->comp,
                    c->acum, c->order, c->rand_uz, c->pi, c->sigmaG, c->part_dot, c->cta_red,
                    c->red_partial, c->slots, c->v, c->betasq, c->out_scalars, c->out_counts, c->prof, c->round_ctr,
                    c->slots5, c->part_bsq, c->part_v, c->prof5, c->results6, c->gc6, c->inmask6, c->prof6, c->dbg6,
                    c->rag.idx16, c->rag.cta_base, c->rag.stats, c->rag.seg};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (void *q : c->peer_opened) cudaIpcCloseMemHandle(q);
    if (c->xchg) cudaFree(c->xchg);
    if (c->peer_ptrs_dev) cudaFree(c->peer_ptrs_dev);
    if (c->h_small) cudaFreeHost(c->h_small);
    if (c->h_counts) cudaFreeHost(c->h_counts);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// the instantiation of the dense sweep kernel for (several GPUs, phase profiling)
static const void *dense_sweep_fn(bool multi, bool prof)
{
    if (multi) return prof ? (const void *)sweep_dense_kernel<true, true> : (const void *)sweep_dense_kernel<true, false>;
    return prof ? (const void *)sweep_dense_kernel<false, true> : (const void *)sweep_dense_kernel<false, false>;
}

static int choose_geometry(brr_gpu_ctx *c)
{
    const brr_gpu_config &cfg = c->cfg;
    uint32_t maxCTA = cfg.max_ctas ? std::min<uint32_t>(cfg.max_ctas, c->numSMs) : c->numSMs;
    const bool dense = cfg.layout == BRR_GPU_LAYOUT_DENSE2BIT;
    c->B = cfg.block_markers ? cfg.block_markers : (dense ? 64 : 32);    // sparse: 32-marker windows measured 22% faster than 16 at N=500k (profiles/)
    // slice = multiple of 64 individuals (one 64-individual MMA k-step pair = 4 code words = 16 bytes, the TMA granule)
    uint32_t per = (cfg.n_inds + maxCTA - 1) / maxCTA;
    uint32_t SL = ((per + 63) / 64) * 64;
    const uint32_t minSL = 256;
    if (SL < minSL) SL = minSL;
    c->SL = SL;
    c->nCTA = (cfg.n_inds + SL - 1) / SL;
    c->PW = SL / 16;
    c->RS = c->PW + (dense ? kTailWords : 0u);
    while (c->RS % 8 != 4) c->RS += 4;      // shared-memory row stride = 4 (mod 8) words: conflict-free fragment loads
    const uint32_t maxSL = dense ? (uint32_t)kMaxSL : 8192u;
    if (SL > maxSL) return fail(BRR_GPU_EINVAL, "n_inds=%u needs %u individuals per CTA; at most %u (=%u per GPU) are supported: shard over more GPUs",
                                cfg.n_inds, SL, maxSL, maxSL * maxCTA);
    return 0;
}

extern "C" int brr_gpu_create(const brr_gpu_config *cfg, brr_gpu_ctx **out)
{
    if (!cfg || !out) return fail(BRR_GPU_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->abi_version != BRR_GPU_ABI_VERSION) return fail(BRR_GPU_EINVAL, "ABI version %u, library is %d", cfg->abi_version, BRR_GPU_ABI_VERSION);
    if (cfg->n_inds < 2 || cfg->n_markers < 1) return fail(BRR_GPU_EINVAL, "need n_inds >= 2 and n_markers >= 1");
    if (cfg->n_components < 2 || cfg->n_components > (uint32_t)kMaxK) return fail(BRR_GPU_EINVAL, "n_components must be 2..%d", kMaxK);
    if (cfg->n_groups < 1 || cfg->n_groups * cfg->n_components > (uint32_t)kMaxGK) return fail(BRR_GPU_EINVAL, "n_groups*n_components must be 1..%d", kMaxGK);
    {
        const uint32_t maxb = cfg->layout == BRR_GPU_LAYOUT_DENSE2BIT ? kMaxB : kMaxW;
        if (cfg->block_markers > maxb) return fail(BRR_GPU_EINVAL, "block_markers must be <= %u for this layout", maxb);
    }
    if (cfg->layout < BRR_GPU_LAYOUT_DENSE2BIT || cfg->layout > BRR_GPU_LAYOUT_EIGEN) return fail(BRR_GPU_EINVAL, "unknown layout %d", cfg->layout);
    if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) return fail(BRR_GPU_EINVAL, "bad rank/world");
    if (cfg->n_inds_total < cfg->n_inds) return fail(BRR_GPU_EINVAL, "n_inds_total < n_inds");
    if (cfg->layout != BRR_GPU_LAYOUT_DENSE2BIT && (cfg->world > 1 || cfg->n_inds_total != cfg->n_inds))
        return fail(BRR_GPU_EINVAL, "the ragged / eigen layouts have no cross-shard exchange: world must be 1 and n_inds_total == n_inds "
                                    "(got world=%d, n_inds=%u, n_inds_total=%u); shard the dense 2-bit layout instead", cfg->world, cfg->n_inds, cfg->n_inds_total);
    if (cfg->world == 1 && cfg->n_inds_total != cfg->n_inds) return fail(BRR_GPU_EINVAL, "world=1 but n_inds_total (%u) != n_inds (%u)", cfg->n_inds_total, cfg->n_inds);
    // the cross-GPU accumulators carry an arrival count above bit 56: |sum of 3 * limb-combined q| must stay below 2^55
    if (cfg->world > 1 && cfg->n_inds_total > 5500000u) return fail(BRR_GPU_EINVAL, "n_inds_total=%u exceeds the 5.5M bound of the cross-GPU accumulators", cfg->n_inds_total);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(BRR_GPU_ENODEVICE, "no CUDA device: this library has no CPU fallback");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(BRR_GPU_ENODEVICE, "device %d of %d", cfg->device, ndev);
    CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) return fail(BRR_GPU_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
    if (!prop.cooperativeLaunch) return fail(BRR_GPU_ENODEVICE, "device lacks cooperative launch");

    brr_gpu_ctx *c = new (std::nothrow) brr_gpu_ctx;
    if (!c) return fail(BRR_GPU_ENOMEM, "out of host memory");
    c->cfg = *cfg;
    c->numSMs = prop.multiProcessorCount;
    int rc = choose_geometry(c);
    if (rc) { delete c; return rc; }

#define TRY_OR_DESTROY(expr) do { int rc_ = (expr); if (rc_) { brr_gpu_destroy(c); return rc_; } } while (0)
#define CUDA_OR_DESTROY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { int rc_ = fail(BRR_GPU_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); brr_gpu_destroy(c); return rc_; } } while (0)
    CUDA_OR_DESTROY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_OR_DESTROY(cudaEventCreate(&c->ev0));
    CUDA_OR_DESTROY(cudaEventCreate(&c->ev1));

    const uint32_t M = cfg->n_markers, N = cfg->n_inds, K = cfg->n_components, nG = cfg->n_groups;
    if (cfg->layout == BRR_GPU_LAYOUT_DENSE2BIT) {
        c->codes_bytes = size_t(M) * c->nCTA * (c->PW + kTailWords) * 4;
        TRY_OR_DESTROY(dev_alloc(&c->codes, c->codes_bytes / 4));
        TRY_OR_DESTROY(dev_alloc(&c->lut, size_t(M) * 4));
        c->smem_bytes = sweep_smem_bytes(c->RS);
        if (c->smem_bytes > size_t(prop.sharedMemPerBlockOptin)) {
            int rc_ = fail(BRR_GPU_EINVAL, "sweep kernel needs %zu bytes of shared memory per CTA (%u individuals per CTA); the device offers %zu: shard over more GPUs",
                           c->smem_bytes, c->SL, size_t(prop.sharedMemPerBlockOptin));
            brr_gpu_destroy(c); return rc_;
        }
        int occ = 8;
        for (const void *fn : {dense_sweep_fn(false, false), dense_sweep_fn(false, true), dense_sweep_fn(true, false), dense_sweep_fn(true, true)}) {
            CUDA_OR_DESTROY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
            int o = 0;
            CUDA_OR_DESTROY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, fn, kThreads, c->smem_bytes));
            occ = std::min(occ, o);
        }
        if (occ < 1 || uint32_t(occ) * c->numSMs < c->nCTA) {
            int rc_ = fail(BRR_GPU_EINVAL, "sweep kernel does not fit: %zu bytes of shared memory per CTA, %u CTAs", c->smem_bytes, c->nCTA);
            brr_gpu_destroy(c); return rc_;
        }
        const size_t slot_ints = size_t(kSlots) * kSlotInts + 8;     // + absmax (2 ints) + barrier counter + pad
        TRY_OR_DESTROY(dev_alloc(&c->slots, slot_ints));
        c->absmax = reinterpret_cast<unsigned long long *>(c->slots + size_t(kSlots) * kSlotInts);
        c->bar = reinterpret_cast<unsigned *>(c->slots + size_t(kSlots) * kSlotInts + 2);
        c->bar2 = c->bar + 1;
        if (const char *e = getenv("BRR_GPU_DENSE_KERNEL")) {
            if (!strcmp(e, "v4")) c->dense_mode = 4; else if (!strcmp(e, "v5")) c->dense_mode = 5; else if (!strcmp(e, "v6")) c->dense_mode = 6;
        }
        c->smem5_bytes = v5::sweep5_smem_bytes(c->RS);
        if (cfg->world == 1 && c->dense_mode != 4 && c->smem5_bytes <= size_t(prop.sharedMemPerBlockOptin)) {
            bool fits = true;
            for (const void *fn : {(const void *)v5::sweep_dense5_kernel<false>, (const void *)v5::sweep_dense5_kernel<true>}) {
                CUDA_OR_DESTROY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem5_bytes));
                int o = 0;
                CUDA_OR_DESTROY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, fn, kThreads, c->smem5_bytes));
                fits = fits && o >= 1 && uint32_t(o) * c->numSMs >= c->nCTA;
            }
            if (fits) {
                TRY_OR_DESTROY(dev_alloc(&c->slots5, size_t(v5::kSlotRing) * v5::kSlotInts5));
                TRY_OR_DESTROY(dev_alloc(&c->part_bsq, size_t(c->nCTA) * 2 * cfg->n_groups));
                TRY_OR_DESTROY(dev_alloc(&c->part_v, size_t(c->nCTA) * cfg->n_groups * cfg->n_components));
                c->v5_ok = true;
            }
        }
        c->smem6_bytes = v6::sweep6_smem_bytes(c->RS, c->SL);
        if (c->dense_mode != 4 && c->dense_mode != 5 && c->smem6_bytes <= size_t(prop.sharedMemPerBlockOptin)) {
            bool fits = true;
            for (const void *fn : {(const void *)v6::sweep_dense6_kernel<false, false>, (const void *)v6::sweep_dense6_kernel<false, true>,
                                   (const void *)v6::sweep_dense6_kernel<true, false>, (const void *)v6::sweep_dense6_kernel<true, true>}) {
                CUDA_OR_DESTROY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem6_bytes));
                int o = 0;
                CUDA_OR_DESTROY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, fn, v6::kThreads6, c->smem6_bytes));
                fits = fits && o >= 1 && uint32_t(o) * c->numSMs >= c->nCTA;
            }
            if (fits) {
                if (!c->slots5) TRY_OR_DESTROY(dev_alloc(&c->slots5, size_t(v5::kSlotRing) * v5::kSlotInts5));
                if (!c->part_bsq) TRY_OR_DESTROY(dev_alloc(&c->part_bsq, size_t(c->nCTA) * 2 * cfg->n_groups));
                if (!c->part_v) TRY_OR_DESTROY(dev_alloc(&c->part_v, size_t(c->nCTA) * cfg->n_groups * cfg->n_components));
                TRY_OR_DESTROY(dev_alloc(&c->results6, size_t(v6::kResRing) * 32));
                TRY_OR_DESTROY(dev_alloc(&c->gc6, size_t(c->nCTA)));
                TRY_OR_DESTROY(dev_alloc(&c->inmask6, size_t(M) / 32 + 4));
                TRY_OR_DESTROY(dev_alloc(&c->dbg6, size_t(c->nCTA) * 32));
                c->v6_ok = true;
            }
        }
        if (c->dense_mode == 6 && !c->v6_ok) {
            int rc_ = fail(BRR_GPU_EINVAL, "BRR_GPU_DENSE_KERNEL=v6 but the version-6 kernel is not available for this context (world=%d, %zu bytes of shared memory per CTA)", cfg->world, c->smem6_bytes);
            brr_gpu_destroy(c); return rc_;
        }
        if (c->dense_mode == 5 && !c->v5_ok) {
            int rc_ = fail(BRR_GPU_EINVAL, "BRR_GPU_DENSE_KERNEL=v5 but the version-5 kernel is not available for this context (world=%d, %zu bytes of shared memory per CTA)", cfg->world, c->smem5_bytes);
            brr_gpu_destroy(c); return rc_;
        }
        if (cfg->world > 1) {
            // (the version-6 import area follows the version-4 exchange buffer in the same, IPC-exported, allocation)
            const size_t xints = xchg_total_ints(cfg->world) + 2 * v6::imp_u64(cfg->world);
            TRY_OR_DESTROY(dev_alloc(&c->xchg, xints));
            CUDA_OR_DESTROY(cudaMemset(c->xchg, 0, sizeof(int) * xints));
            const int hdr = (int)c->nCTA;       // peers read this at attach time to agree on the number of exchange CTAs
            CUDA_OR_DESTROY(cudaMemcpy(c->xchg + xchg_total_ints(cfg->world) - 8, &hdr, sizeof(int), cudaMemcpyHostToDevice));
            TRY_OR_DESTROY(dev_alloc(&c->peer_ptrs_dev, (size_t)cfg->world));
            TRY_OR_DESTROY(dev_alloc(&c->round_ctr, 1));
            CUDA_OR_DESTROY(cudaMemset(c->round_ctr, 0, sizeof(unsigned long long)));
        }
    } else {
        TRY_OR_DESTROY(dev_alloc(&c->bar, 4));
        c->slots = reinterpret_cast<int *>(c->bar);      // single owner pointer for destroy
        c->smem_bytes = ragged_smem_bytes(c->SL);
        CUDA_OR_DESTROY(cudaFuncSetAttribute(sweep_ragged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
        int occ = 0;
        CUDA_OR_DESTROY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_ragged_kernel, kThreads, c->smem_bytes));
        if (occ < 1 || uint32_t(occ) * c->numSMs < c->nCTA) {
            int rc_ = fail(BRR_GPU_EINVAL, "ragged sweep kernel does not fit: %zu bytes of shared memory per CTA, %u CTAs", c->smem_bytes, c->nCTA);
            brr_gpu_destroy(c); return rc_;
        }
    }
    TRY_OR_DESTROY(dev_alloc(&c->counts_local, size_t(M) * 4));
    TRY_OR_DESTROY(dev_alloc(&c->counts_global, size_t(M) * 4));
    TRY_OR_DESTROY(dev_alloc(&c->group, M));
    TRY_OR_DESTROY(dev_alloc(&c->cva, size_t(nG) * (K - 1)));
    TRY_OR_DESTROY(dev_alloc(&c->eps, N));
    TRY_OR_DESTROY(dev_alloc(&c->beta, M));
    TRY_OR_DESTROY(dev_alloc(&c->comp, M));
    TRY_OR_DESTROY(dev_alloc(&c->acum, M));
    TRY_OR_DESTROY(dev_alloc(&c->order, M));
    TRY_OR_DESTROY(dev_alloc(&c->rand_uz, size_t(M) * 2));
    TRY_OR_DESTROY(dev_alloc(&c->pi, size_t(nG) * K));
    TRY_OR_DESTROY(dev_alloc(&c->sigmaG, nG));
    TRY_OR_DESTROY(dev_alloc(&c->part_dot, size_t(2) * c->nCTA * kMaxB));
    TRY_OR_DESTROY(dev_alloc(&c->cta_red, size_t(2) * std::max<uint32_t>(c->nCTA, kRedBlocks)));
    TRY_OR_DESTROY(dev_alloc(&c->red_partial, size_t(2) * kRedBlocks));
    TRY_OR_DESTROY(dev_alloc(&c->v, size_t(nG) * K));
    TRY_OR_DESTROY(dev_alloc(&c->betasq, nG));
    TRY_OR_DESTROY(dev_alloc(&c->out_scalars, 4));
    TRY_OR_DESTROY(dev_alloc(&c->out_counts, 4));
    if (const char *e = getenv("BRR_GPU_PROFILE")) {
        if (e[0] == '1') {
            TRY_OR_DESTROY(dev_alloc(&c->prof, kProfPhases));
            CUDA_OR_DESTROY(cudaMemset(c->prof, 0, sizeof(unsigned long long) * kProfPhases));
            TRY_OR_DESTROY(dev_alloc(&c->prof5, v5::kPfN));
            CUDA_OR_DESTROY(cudaMemset(c->prof5, 0, sizeof(unsigned long long) * v5::kPfN));
            TRY_OR_DESTROY(dev_alloc(&c->prof6, v6::kPfN + size_t(8) * c->nCTA));
            CUDA_OR_DESTROY(cudaMemset(c->prof6, 0, sizeof(unsigned long long) * (v6::kPfN + size_t(8) * c->nCTA)));
        }
    }
    CUDA_OR_DESTROY(cudaMallocHost(reinterpret_cast<void **>(&c->h_small), sizeof(double) * (2 * kMaxGK + 8 + 2 * kRedBlocks)));
    CUDA_OR_DESTROY(cudaMallocHost(reinterpret_cast<void **>(&c->h_counts), sizeof(unsigned) * 8));
    CUDA_OR_DESTROY(cudaMemsetAsync(c->eps, 0, sizeof(double) * N, c->stream));
    CUDA_OR_DESTROY(cudaMemsetAsync(c->beta, 0, sizeof(double) * M, c->stream));
    CUDA_OR_DESTROY(cudaMemsetAsync(c->comp, 0, sizeof(double) * M, c->stream));
    CUDA_OR_DESTROY(cudaMemsetAsync(c->acum, 0, sizeof(double) * M, c->stream));
    CUDA_OR_DESTROY(cudaMemsetAsync(c->group, 0, sizeof(int32_t) * M, c->stream));
    CUDA_OR_DESTROY(cudaMemsetAsync(c->counts_local, 0, sizeof(uint32_t) * 4 * M, c->stream));
    CUDA_OR_DESTROY(cudaStreamSynchronize(c->stream));
    *out = c;
    return 0;
}

#define CHECK_CTX(c)                                                  \
    do {                                                              \
        if (!(c)) return fail(BRR_GPU_EINVAL, "null context");        \
        CUDA_TRY(cudaSetDevice((c)->cfg.device));                     \
    } while (0)

// ---- column supply -----------------------------------------------------------------------------

extern "C" int brr_gpu_upload_dense2bit(brr_gpu_ctx *c, const uint8_t *bed, size_t col_stride, uint32_t first, uint32_t count)
{
    CHECK_CTX(c);
    if (c->cfg.layout != BRR_GPU_LAYOUT_DENSE2BIT) return fail(BRR_GPU_ESTATE, "context layout is not dense2bit");
    if (!bed || uint64_t(first) + uint64_t(count) > uint64_t(c->cfg.n_markers)) return fail(BRR_GPU_EINVAL, "bad marker range");
    const size_t col_bytes = (size_t(c->cfg.n_inds) + 3) / 4;
    if (col_stride < col_bytes) return fail(BRR_GPU_EINVAL, "col_stride %zu < %zu", col_stride, col_bytes);
    const size_t chunk_markers = std::max<size_t>(1, std::min<size_t>(count, (size_t(256) << 20) / col_bytes));
    uint8_t *stage = nullptr;
    int rc = dev_alloc(&stage, chunk_markers * col_bytes);
    if (rc) return rc;
    for (uint32_t done = 0; done < count;) {
        const uint32_t n = (uint32_t)std::min<size_t>(chunk_markers, count - done);
        cudaError_t e = cudaMemcpy2DAsync(stage, col_bytes, bed + size_t(done) * col_stride, col_stride, col_bytes, n,
                                          cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) {
            for (uint32_t y0 = 0; y0 < n; y0 += 65535) {
                const uint32_t ny = std::min<uint32_t>(65535, n - y0);
                bed_to_codes_kernel<<<dim3(c->nCTA, ny), 128, 0, c->stream>>>(stage + size_t(y0) * col_bytes, col_bytes,
                                                                           first + done + y0, c->cfg.n_inds, c->nCTA,
                                                                           c->SL, c->PW, c->codes);
                ++c->launches;
            }
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { cudaFree(stage); return fail(BRR_GPU_ECUDA, "upload failed: %s", cudaGetErrorString(e)); }
        done += n;
    }
    cudaFree(stage);
    const uint32_t wpb = 8;
    {
        const size_t n_rows = size_t(count) * c->nCTA;
        build_row_tails_kernel<<<(unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, 0, c->stream>>>(c->codes, size_t(first) * c->nCTA, n_rows, c->PW);
    }
    count_classes_kernel<<<(count + wpb - 1) / wpb, wpb * 32, 0, c->stream>>>(c->codes, first, count, c->cfg.n_inds, c->nCTA, c->PW, c->counts_local);
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->uploaded = true;
    c->finalized = false;
    return 0;
}

extern "C" int brr_gpu_get_counts(brr_gpu_ctx *c, uint32_t *counts)
{
    CHECK_CTX(c);
    if (!counts) return fail(BRR_GPU_EINVAL, "null counts");
    if (!c->uploaded) return fail(BRR_GPU_ESTATE, "no markers uploaded");
    CUDA_TRY(cudaMemcpy(counts, c->counts_local, sizeof(uint32_t) * 4 * c->cfg.n_markers, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int brr_gpu_finalize_markers(brr_gpu_ctx *c, const uint32_t *global_counts)
{
    CHECK_CTX(c);
    if (!c->uploaded) return fail(BRR_GPU_ESTATE, "no markers uploaded");
    const uint32_t M = c->cfg.n_markers;
    if (global_counts)
        CUDA_TRY(cudaMemcpyAsync(c->counts_global, global_counts, sizeof(uint32_t) * 4 * M, cudaMemcpyHostToDevice, c->stream));
    else
        CUDA_TRY(cudaMemcpyAsync(c->counts_global, c->counts_local, sizeof(uint32_t) * 4 * M, cudaMemcpyDeviceToDevice, c->stream));
    if (c->cfg.layout == BRR_GPU_LAYOUT_DENSE2BIT) {
        dense_lut_kernel<<<(M + 255) / 256, 256, 0, c->stream>>>(c->counts_global, M, c->cfg.n_inds_total, c->lut);
    } else {
        if (!c->rag.stats_given)
            ragged_stats_kernel<<<(M + 255) / 256, 256, 0, c->stream>>>(c->counts_global, M, c->cfg.n_inds_total,
                                                                        c->cfg.layout == BRR_GPU_LAYOUT_EIGEN, c->rag.stats);
    }
    ++c->launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->finalized = true;
    return 0;
}

extern "C" int brr_gpu_set_dense_lut(brr_gpu_ctx *c, const double *lut)
{
    CHECK_CTX(c);
    if (c->cfg.layout != BRR_GPU_LAYOUT_DENSE2BIT || !lut) return fail(BRR_GPU_EINVAL, "dense layout and non-null lut required");
    if (!c->finalized) return fail(BRR_GPU_ESTATE, "call brr_gpu_finalize_markers first (the class counts are needed too)");
    CUDA_TRY(cudaMemcpy(c->lut, lut, sizeof(double) * 4 * c->cfg.n_markers, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int brr_gpu_get_dense_lut(brr_gpu_ctx *c, double *lut)
{
    CHECK_CTX(c);
    if (c->cfg.layout != BRR_GPU_LAYOUT_DENSE2BIT || !lut) return fail(BRR_GPU_EINVAL, "dense layout and non-null lut required");
    if (!c->finalized) return fail(BRR_GPU_ESTATE, "markers not finalized");
    CUDA_TRY(cudaMemcpy(lut, c->lut, sizeof(double) * 4 * c->cfg.n_markers, cudaMemcpyDeviceToHost));
    return 0;
}

// ---- model and state ---------------------------------------------------------------------------

extern "C" int brr_gpu_set_model(brr_gpu_ctx *c, const int32_t *group_of_marker, const double *cva)
{
    CHECK_CTX(c);
    if (!cva) return fail(BRR_GPU_EINVAL, "null cva");
    const uint32_t M = c->cfg.n_markers, nG = c->cfg.n_groups, K = c->cfg.n_components;
    if (group_of_marker) {
        for (uint32_t j = 0; j < M; ++j)
            if (group_of_marker[j] < 0 || uint32_t(group_of_marker[j]) >= nG)
                return fail(BRR_GPU_EINVAL, "group_of_marker[%u]=%d outside 0..%u", j, group_of_marker[j], nG - 1);
        CUDA_TRY(cudaMemcpy(c->group, group_of_marker, sizeof(int32_t) * M, cudaMemcpyHostToDevice));
    } else {
        CUDA_TRY(cudaMemset(c->group, 0, sizeof(int32_t) * M));
    }
    for (uint32_t i = 0; i < nG * (K - 1); ++i)
        if (!(cva[i] > 0.0)) return fail(BRR_GPU_EINVAL, "cva[%u] must be positive", i);
    CUDA_TRY(cudaMemcpy(c->cva, cva, sizeof(double) * nG * (K - 1), cudaMemcpyHostToDevice));
    c->model_set = true;
    return 0;
}

extern "C" int brr_gpu_set_state(brr_gpu_ctx *c, const double *eps, const double *beta, const double *components, const double *acum)
{
    CHECK_CTX(c);
    const uint32_t M = c->cfg.n_markers, N = c->cfg.n_inds;
    if (eps) CUDA_TRY(cudaMemcpyAsync(c->eps, eps, sizeof(double) * N, cudaMemcpyHostToDevice, c->stream));
    if (beta) CUDA_TRY(cudaMemcpyAsync(c->beta, beta, sizeof(double) * M, cudaMemcpyHostToDevice, c->stream));
    if (components) CUDA_TRY(cudaMemcpyAsync(c->comp, components, sizeof(double) * M, cudaMemcpyHostToDevice, c->stream));
    if (acum) CUDA_TRY(cudaMemcpyAsync(c->acum, acum, sizeof(double) * M, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (beta) c->have_last = false;          // the kernel choice of the next sweep must not rely on the previous chain
    return 0;
}

extern "C" int brr_gpu_get_state(brr_gpu_ctx *c, double *eps, double *beta, double *components, double *acum)
{
    CHECK_CTX(c);
    const uint32_t M = c->cfg.n_markers, N = c->cfg.n_inds;
    if (eps) CUDA_TRY(cudaMemcpyAsync(eps, c->eps, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
    if (beta) CUDA_TRY(cudaMemcpyAsync(beta, c->beta, sizeof(double) * M, cudaMemcpyDeviceToHost, c->stream));
    if (components) CUDA_TRY(cudaMemcpyAsync(components, c->comp, sizeof(double) * M, cudaMemcpyDeviceToHost, c->stream));
    if (acum) CUDA_TRY(cudaMemcpyAsync(acum, c->acum, sizeof(double) * M, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

static int eps_reduce_impl(brr_gpu_ctx *c, double add, int do_add, double *sum_out, double *sq_out)
{
    eps_add_reduce_kernel<<<kRedBlocks, 256, 0, c->stream>>>(c->eps, c->cfg.n_inds, add, do_add, c->red_partial);
    ++c->launches;
    CUDA_TRY(cudaGetLastError());
    double *h = c->h_small + 2 * kMaxGK + 8;
    CUDA_TRY(cudaMemcpyAsync(h, c->red_partial, sizeof(double) * 2 * kRedBlocks, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    double s = 0.0, q = 0.0;
    for (int i = 0; i < kRedBlocks; ++i) { s += h[2 * i]; q += h[2 * i + 1]; }
    if (sum_out) *sum_out = s;
    if (sq_out) *sq_out = q;
    return 0;
}

extern "C" int brr_gpu_eps_add_scalar(brr_gpu_ctx *c, double v, double *sum_out)
{
    CHECK_CTX(c);
    return eps_reduce_impl(c, v, 1, sum_out, nullptr);
}

extern "C" int brr_gpu_eps_reduce(brr_gpu_ctx *c, double *sum_out, double *sqnorm_out)
{
    CHECK_CTX(c);
    return eps_reduce_impl(c, 0.0, 0, sum_out, sqnorm_out);
}

// ---- the sweep ---------------------------------------------------------------------------------

extern "C" int brr_gpu_sweep_stage(brr_gpu_ctx *c, const uint32_t *marker_order, uint32_t n_order, const double *rand_uz,
                                   const double *pi, const double *sigmaG, double sigmaE, double eps_sum_in,
                                   uint32_t async_tokens)
{
    CHECK_CTX(c);
    if (!c->uploaded || !c->finalized) return fail(BRR_GPU_ESTATE, "upload and finalize the markers before sweeping");
    if (!c->model_set) return fail(BRR_GPU_ESTATE, "brr_gpu_set_model not called");
    if (!marker_order || !rand_uz || !pi || !sigmaG) return fail(BRR_GPU_EINVAL, "null sweep input");
    const uint32_t M = c->cfg.n_markers, nG = c->cfg.n_groups, K = c->cfg.n_components;
    if (n_order > M) return fail(BRR_GPU_EINVAL, "n_order %u > n_markers %u", n_order, M);
    {
        const uint32_t maxb = c->cfg.layout == BRR_GPU_LAYOUT_DENSE2BIT ? kMaxB : kMaxW;
        if (async_tokens > maxb) return fail(BRR_GPU_EINVAL, "async_tokens must be <= %u for this layout", maxb);
    }
    if (!(sigmaE > 0.0)) return fail(BRR_GPU_EINVAL, "sigmaE must be positive");
    {   // distinct, in-range marker ids (stamps of the current call instead of a cleared table: this runs once per sweep on the caller's thread)
        if (c->order_stamp.size() != M) { c->order_stamp.assign(M, 0u); c->order_gen = 0; }
        if (++c->order_gen == 0u) { std::fill(c->order_stamp.begin(), c->order_stamp.end(), 0u); c->order_gen = 1; }
        const uint32_t gen = c->order_gen;
        uint32_t *stamp = c->order_stamp.data();
        for (uint32_t i = 0; i < n_order; ++i) {
            const uint32_t j = marker_order[i];
            if (j >= M) return fail(BRR_GPU_EINVAL, "marker_order[%u]=%u out of range", i, j);
            if (stamp[j] == gen) return fail(BRR_GPU_EINVAL, "marker %u appears twice in marker_order", j);
            stamp[j] = gen;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(c->order, marker_order, sizeof(uint32_t) * n_order, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->rand_uz, rand_uz, sizeof(double) * 2 * M, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->pi, pi, sizeof(double) * nG * K, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->sigmaG, sigmaG, sizeof(double) * nG, cudaMemcpyHostToDevice, c->stream));
    c->n_order = n_order; c->sigmaE = sigmaE; c->eps_sum_in = eps_sum_in; c->async_tokens = async_tokens;
    c->staged = true;
    return 0;
}

static int launch_dense(brr_gpu_ctx *c)
{
    DenseSweepParams p{};
    p.n_local = c->cfg.n_inds; p.n_total = c->cfg.n_inds_total; p.M = c->cfg.n_markers; p.K = c->cfg.n_components;
    p.nG = c->cfg.n_groups; p.B = c->async_tokens ? c->async_tokens : c->B; p.nCTA = c->nCTA; p.SL = c->SL; p.PW = c->PW; p.RS = c->RS;
    p.codes = c->codes; p.lut = c->lut; p.counts = c->counts_global; p.group = c->group;
    p.order = c->order; p.n_order = c->n_order; p.rand_uz = c->rand_uz; p.pi = c->pi; p.sigmaG = c->sigmaG; p.cva = c->cva;
    p.sigmaE = c->sigmaE; p.async_mode = c->async_tokens ? 1u : 0u;
    p.eps = c->eps; p.beta = c->beta; p.comp = c->comp; p.acum = c->acum;
    p.slots = c->slots; p.absmax = c->absmax; p.bar = c->bar; p.cta_red = c->cta_red;
    p.v = c->v; p.betasq = c->betasq; p.out_scalars = c->out_scalars; p.out_counts = c->out_counts;
    p.prof = c->prof;
    p.world = c->cfg.world; p.rank = c->cfg.rank; p.peer_xchg = c->peer_ptrs_dev; p.bar2 = c->bar2; p.round_ctr = c->round_ctr; p.xchg_ctas = c->xchg_ctas;
    if (c->cfg.world > 1 && !c->attached)
        return fail(BRR_GPU_ESTATE, "world=%d but brr_gpu_ipc_attach has not been called: the shards cannot exchange their partial sums", c->cfg.world);
    CUDA_TRY(cudaMemsetAsync(c->slots, 0, sizeof(int) * (size_t(kSlots) * kSlotInts + 8), c->stream));
    CUDA_TRY(cudaMemsetAsync(c->out_counts, 0, sizeof(unsigned) * 4, c->stream));
    void *args[] = {&p};
    // per-function state, shared by every context of the process: re-apply this context's size before each launch
    CUDA_TRY(cudaFuncSetAttribute(dense_sweep_fn(c->cfg.world > 1, p.prof != nullptr), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    CUDA_TRY(cudaLaunchCooperativeKernel(dense_sweep_fn(c->cfg.world > 1, p.prof != nullptr), dim3(c->nCTA), dim3(kThreads), args,
                                         c->smem_bytes, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    c->launches += 1; c->last_launches = 1;
    return 0;
}

// Version 5 (sweep_dense5.cuh): the steady-state kernel.  Chosen per sweep: it restarts its pipeline at every marker that ENTERS
// the model, so it is used when the previous sweep of this chain had few non-skipped markers (the chain has left its start), and
// version 4 (on-demand Gram rows) otherwise -- both produce the same chain.  BRR_GPU_DENSE_KERNEL=v4|v5 forces one of them.
static const double kV5MaxNonskipped = 0.03;
static bool use_v5(const brr_gpu_ctx *c)
{
    if (!c->v5_ok || c->async_tokens || c->cfg.world > 1) return false;
    if (c->dense_mode == 5) return true;
    if (c->dense_mode == 4 || c->dense_mode == 6) return false;
    return !c->v6_ok && c->have_last && c->last_nonskipped_frac <= kV5MaxNonskipped;
}
// Version 6 (sweep_dense6.cuh): as version 5 with the walk distributed over the CTAs; preferred over version 5 when it fits.
static bool use_v6(const brr_gpu_ctx *c)
{
    if (!c->v6_ok || c->async_tokens || (c->cfg.world > 1 && !c->attached) || c->cfg.world > v6::kMaxWorld) return false;
    if (c->dense_mode == 6) return true;
    if (c->dense_mode == 4 || c->dense_mode == 5) return false;
    return c->have_last && c->last_nonskipped_frac <= kV5MaxNonskipped;
}

static int launch_dense6(brr_gpu_ctx *c)
{
    v6::Sweep6Params p{};
    p.n_local = c->cfg.n_inds; p.n_total = c->cfg.n_inds_total; p.M = c->cfg.n_markers; p.K = c->cfg.n_components;
    p.nG = c->cfg.n_groups; p.nCTA = c->nCTA; p.SL = c->SL; p.PW = c->PW; p.RS = c->RS;
    p.codes = c->codes; p.lut = c->lut; p.counts = c->counts_global; p.group = c->group;
    p.order = c->order; p.n_order = c->n_order; p.inmask = c->inmask6; p.rand_uz = c->rand_uz; p.pi = c->pi; p.sigmaG = c->sigmaG; p.cva = c->cva;
    p.sigmaE = c->sigmaE;
    p.eps = c->eps; p.beta = c->beta; p.comp = c->comp; p.acum = c->acum;
    p.slots = c->slots5; p.results = c->results6; p.gc = c->gc6; p.absmax = c->absmax; p.bar = c->bar; p.cta_red = c->cta_red;
    p.part_bsq = c->part_bsq; p.part_v = c->part_v;
    p.v = c->v; p.betasq = c->betasq; p.out_scalars = c->out_scalars; p.out_counts = c->out_counts;
    p.prof = c->prof6; p.dbg = c->dbg6;
    p.world = c->cfg.world; p.rank = c->cfg.rank; p.peer_xchg = c->peer_ptrs_dev; p.imp_off = xchg_total_ints(c->cfg.world);
    // Pipeline depth: 4 on one GPU (0.075 s per sweep; 5: 0.078, 7: 0.088), 7 across GPUs, where a walk waits for every peer's sums (8 GPUs: 0.072 s
    // against 0.086 at depth 4; 2 GPUs: 4, 6, 7 time the same).  Which updates reach a dot through the residual and which through a Gram correction
    // depends on the depth, so chains of different depths agree to rounding (1e-16 in a residual), not bit for bit; all ranks of one run are identical.
    p.depth = c->cfg.world > 1 ? 7u : 4u;
    if (const char *e = getenv("BRR_GPU_V6_DEPTH")) { const int d = atoi(e); if (d >= 1 && d <= v6::kD) p.depth = uint32_t(d); }
    p.launch_tag = ++c->v6_launches;                      // (every rank runs the same sequence of sweeps and makes the same kernel choice)
    CUDA_TRY(cudaMemsetAsync(c->dbg6, 0, sizeof(unsigned) * size_t(c->nCTA) * 32, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->slots5, 0, sizeof(int) * size_t(v5::kSlotRing) * v5::kSlotInts5, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->results6, 0, sizeof(int) * size_t(v6::kResRing) * 32, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->slots + size_t(kSlots) * kSlotInts, 0, sizeof(int) * 8, c->stream));     // absmax + barrier counters
    CUDA_TRY(cudaMemsetAsync(c->out_counts, 0, sizeof(unsigned) * 4, c->stream));
    const bool mg = c->cfg.world > 1;
    const void *fn = mg ? (p.prof ? (const void *)v6::sweep_dense6_kernel<true, true> : (const void *)v6::sweep_dense6_kernel<true, false>)
                        : (p.prof ? (const void *)v6::sweep_dense6_kernel<false, true> : (const void *)v6::sweep_dense6_kernel<false, false>);
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem6_bytes));
    void *args[] = {&p};
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    const uint32_t n_words = (c->n_order + 31) / 32;
    if (n_words) v6::inmodel_mask_kernel<<<(n_words * 32 + 255) / 256, 256, 0, c->stream>>>(c->order, c->n_order, c->beta, c->inmask6, n_words);
    CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(c->nCTA), dim3(v6::kThreads6), args, c->smem6_bytes, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    c->launches += 1 + (n_words ? 1 : 0); c->last_launches = 1 + (n_words ? 1 : 0);
    return 0;
}

static int launch_dense5(brr_gpu_ctx *c)
{
    v5::Sweep5Params p{};
    p.n_local = c->cfg.n_inds; p.n_total = c->cfg.n_inds_total; p.M = c->cfg.n_markers; p.K = c->cfg.n_components;
    p.nG = c->cfg.n_groups; p.nCTA = c->nCTA; p.SL = c->SL; p.PW = c->PW; p.RS = c->RS;
    p.codes = c->codes; p.lut = c->lut; p.counts = c->counts_global; p.group = c->group;
    p.order = c->order; p.n_order = c->n_order; p.rand_uz = c->rand_uz; p.pi = c->pi; p.sigmaG = c->sigmaG; p.cva = c->cva;
    p.sigmaE = c->sigmaE;
    p.eps = c->eps; p.beta = c->beta; p.comp = c->comp; p.acum = c->acum;
    p.slots = c->slots5; p.absmax = c->absmax; p.bar = c->bar; p.cta_red = c->cta_red;
    p.part_bsq = c->part_bsq; p.part_v = c->part_v;
    p.v = c->v; p.betasq = c->betasq; p.out_scalars = c->out_scalars; p.out_counts = c->out_counts;
    p.prof = c->prof5;
    CUDA_TRY(cudaMemsetAsync(c->slots5, 0, sizeof(int) * size_t(v5::kSlotRing) * v5::kSlotInts5, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->slots + size_t(kSlots) * kSlotInts, 0, sizeof(int) * 8, c->stream));     // absmax + barrier counters
    CUDA_TRY(cudaMemsetAsync(c->out_counts, 0, sizeof(unsigned) * 4, c->stream));
    const void *fn = p.prof ? (const void *)v5::sweep_dense5_kernel<true> : (const void *)v5::sweep_dense5_kernel<false>;
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem5_bytes));
    void *args[] = {&p};
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(c->nCTA), dim3(kThreads), args, c->smem5_bytes, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    c->launches += 1; c->last_launches = 1;
    return 0;
}

static int launch_ragged(brr_gpu_ctx *c);

extern "C" int brr_gpu_sweep_run(brr_gpu_ctx *c)
{
    CHECK_CTX(c);
    if (!c->staged) return fail(BRR_GPU_ESTATE, "brr_gpu_sweep_stage not called");
    int rc;
    if (c->cfg.layout == BRR_GPU_LAYOUT_DENSE2BIT) {
        const bool six = use_v6(c), five = !six && use_v5(c);
        rc = six ? launch_dense6(c) : five ? launch_dense5(c) : launch_dense(c);
        c->last_kernel = six ? 6u : five ? 5u : 4u;
    } else {
        rc = launch_ragged(c);
        c->last_kernel = 0;
    }
    if (rc) return rc;
    c->ran = true;
    return 0;
}

extern "C" int brr_gpu_sweep_fetch(brr_gpu_ctx *c, double *v, double *betasqn, brr_gpu_sweep_out *out)
{
    CHECK_CTX(c);
    if (!c->ran) return fail(BRR_GPU_ESTATE, "no sweep has run");
    const uint32_t nG = c->cfg.n_groups, K = c->cfg.n_components;
    double *h = c->h_small;
    CUDA_TRY(cudaMemcpyAsync(h, c->v, sizeof(double) * nG * K, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(h + kMaxGK, c->betasq, sizeof(double) * nG, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(h + 2 * kMaxGK, c->out_scalars, sizeof(double) * 2, cudaMemcpyDeviceToHost, c->stream));   // [0] = sum eps, [1] = ||eps||^2
    CUDA_TRY(cudaMemcpyAsync(c->h_counts, c->out_counts, sizeof(unsigned) * 4, cudaMemcpyDeviceToHost, c->stream));
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return fail(BRR_GPU_ECUDA, "sweep failed: %s", cudaGetErrorString(e));
    if ((c->h_counts[2] & 4u) && c->last_kernel == 6 && c->dbg6 && getenv("BRR_GPU_V6_DBG")) {     // where every warp of the reported CTA (and of CTA 0) was waiting
        std::vector<unsigned> h(size_t(c->nCTA) * 32);
        cudaMemcpy(h.data(), c->dbg6, sizeof(unsigned) * h.size(), cudaMemcpyDeviceToHost);
        for (uint32_t cta = 1; cta < c->nCTA; ++cta) {          // CTAs whose warps 2..19 wait somewhere else than CTA 0's
            bool same = true;
            for (int w = 2; w < 20; ++w) same = same && h[size_t(cta) * 32 + w] == h[w];
            if (same) continue;
            fprintf(stderr, "[brr_gpu v6] differs: CTA %u:", cta);
            for (int w = 0; w < 20; ++w) { const unsigned x = h[size_t(cta) * 32 + w]; fprintf(stderr, " w%d:%u@%u", w, x & 0xfffu, (x >> 12) & 0x7ffffu); }
            fprintf(stderr, "\n");
        }
        for (uint32_t cta : {0u, (c->h_counts[3] >> 16) & 0x7fffu, ((h[12] >> 12) & 1023u) % c->nCTA, (((h[12] >> 12) & 1023u) + 1) % c->nCTA}) {
            if (cta >= c->nCTA) continue;
            fprintf(stderr, "[brr_gpu v6] CTA %u:", cta);
            for (int w = 0; w < 20; ++w) { const unsigned x = h[size_t(cta) * 32 + w]; fprintf(stderr, " w%d:%u@%u", w, x & 0xfffu, (x >> 12) & 0x7ffffu); }
            fprintf(stderr, "\n");
        }
    }
    if (c->h_counts[2] != 0)
        return fail(BRR_GPU_ESTATE, "sweep failed (flags 0x%x):%s%s%s; the chain state on the device is no longer valid -- restore it with brr_gpu_set_state", c->h_counts[2],
                    (c->h_counts[2] & 1u) ? " the residual grew past the fixed-point range of the dot products inside one sweep" : "",
                    (c->h_counts[2] & 2u) ? " a peer GPU did not deliver its sums in time (is every rank calling brr_gpu_sweep?)" : "",
                    (c->h_counts[2] & 4u) ? (std::string(" a wait inside the sweep pipeline timed out (internal protocol error; site ") + std::to_string(c->h_counts[3] & 0xffu) + ", warp " + std::to_string((c->h_counts[3] >> 8) & 0xffu) + ", CTA " + std::to_string((c->h_counts[3] >> 16) & 0x7fffu) + ")").c_str() : "");
    if (v) memcpy(v, h, sizeof(double) * nG * K);
    if (betasqn) memcpy(betasqn, h + kMaxGK, sizeof(double) * nG);
    c->have_last = c->n_order > 0 && c->n_order == c->cfg.n_markers;
    c->last_nonskipped_frac = c->n_order ? double(c->h_counts[0]) / double(c->n_order) : 1.0;
    c->last_restarts = c->last_kernel >= 5 ? c->h_counts[3] : 0u;
    cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1);
    if (out) {
        out->eps_sum = h[2 * kMaxGK]; out->eps_sqnorm = h[2 * kMaxGK + 1];
        out->n_nonskipped = c->h_counts[0]; out->n_blocks = c->h_counts[1];
        out->kernel_ms = c->last_ms; out->n_launches = c->last_launches;
    }
    return 0;
}

extern "C" int brr_gpu_sweep(brr_gpu_ctx *c, const uint32_t *marker_order, uint32_t n_order, const double *rand_uz,
                             const double *pi, const double *sigmaG, double sigmaE, double eps_sum_in, uint32_t async_tokens,
                             double *v, double *betasqn, brr_gpu_sweep_out *out)
{
    int rc = brr_gpu_sweep_stage(c, marker_order, n_order, rand_uz, pi, sigmaG, sigmaE, eps_sum_in, async_tokens);
    if (rc) return rc;
    rc = brr_gpu_sweep_run(c);
    if (rc) return rc;
    return brr_gpu_sweep_fetch(c, v, betasqn, out);
}

// ---- marker-level parity hooks -----------------------------------------------------------------

extern "C" int brr_gpu_compute_num(brr_gpu_ctx *c, uint32_t marker, double beta_old, double eps_sum, double *num_out)
{
    CHECK_CTX(c);
    if (!c->uploaded || !c->finalized) return fail(BRR_GPU_ESTATE, "upload and finalize the markers first");
    if (marker >= c->cfg.n_markers || !num_out) return fail(BRR_GPU_EINVAL, "bad marker / null output");
    const double NM1 = double(c->cfg.n_inds_total) - 1.0;
    if (c->cfg.layout == BRR_GPU_LAYOUT_DENSE2BIT) {
        dense_dot_kernel<<<c->nCTA, 256, 0, c->stream>>>(c->codes, c->lut, marker, c->eps, c->cfg.n_inds, c->nCTA, c->SL, c->PW, c->cta_red);
        ++c->launches;
        CUDA_TRY(cudaGetLastError());
        std::vector<double> h(c->nCTA);
        CUDA_TRY(cudaMemcpyAsync(h.data(), c->cta_red, sizeof(double) * c->nCTA, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        double s = 0.0;
        for (double x : h) s += x;
        *num_out = s + (c->cfg.rank == 0 ? beta_old * NM1 : 0.0);      // src/densebayesrkernel.cpp:14
        return 0;
    }
    return ragged_compute_num(c, marker, beta_old, eps_sum, num_out);
}

extern "C" int brr_gpu_apply_delta(brr_gpu_ctx *c, uint32_t marker, double beta_old, double beta_new, double *eps_sum_delta)
{
    CHECK_CTX(c);
    if (!c->uploaded || !c->finalized) return fail(BRR_GPU_ESTATE, "upload and finalize the markers first");
    if (marker >= c->cfg.n_markers) return fail(BRR_GPU_EINVAL, "bad marker");
    if (c->cfg.layout == BRR_GPU_LAYOUT_DENSE2BIT) {
        dense_axpy_kernel<<<(c->cfg.n_inds + 255) / 256, 256, 0, c->stream>>>(c->codes, c->lut, marker, beta_old - beta_new, c->eps,
                                                                          c->cfg.n_inds, c->nCTA, c->SL, c->PW);
        ++c->launches;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (eps_sum_delta) *eps_sum_delta = 0.0;
        return 0;
    }
    return ragged_apply_delta(c, marker, beta_old, beta_new, eps_sum_delta);
}

// ---- synthetic data ----------------------------------------------------------------------------

static void synth_thresholds(uint64_t seed, double maf_lo, double maf_hi, int log_uniform, double p_missing, uint32_t first,
                             uint32_t count, uint32_t *thr)
{
    for (uint32_t i = 0; i < count; ++i) {
        const uint32_t mk = first + i;
        const double u = double(mix64(seed * 0x2545F4914F6CDD1DULL + uint64_t(mk) * 0x9E3779B97F4A7C15ULL + 0x1234567ULL) >> 11) * (1.0 / 9007199254740992.0);
        const double maf = log_uniform ? maf_lo * std::pow(maf_hi / maf_lo, u) : maf_lo + (maf_hi - maf_lo) * u;
        const double p0 = (1.0 - maf) * (1.0 - maf), p1 = 2.0 * maf * (1.0 - maf);
        const double two32 = 4294967296.0;
        thr[3 * i + 0] = (uint32_t)std::min(4294967295.0, std::floor(p0 * two32));
        thr[3 * i + 1] = (uint32_t)std::min(4294967295.0, std::floor((p0 + p1) * two32));
        thr[3 * i + 2] = (uint32_t)std::min(4294967295.0, std::floor(p_missing * two32));
    }
}

extern "C" int brr_gpu_synth_dense(brr_gpu_ctx *c, uint64_t seed, double maf_lo, double maf_hi, int log_uniform, double p_missing,
                                   uint32_t ind_offset)
{
    CHECK_CTX(c);
    if (c->cfg.layout != BRR_GPU_LAYOUT_DENSE2BIT) return fail(BRR_GPU_ESTATE, "context layout is not dense2bit");
    if (!(maf_lo > 0.0 && maf_hi <= 0.5 && maf_lo <= maf_hi) || p_missing < 0.0 || p_missing >= 1.0) return fail(BRR_GPU_EINVAL, "bad maf / missing parameters");
    const uint32_t M = c->cfg.n_markers;
    std::vector<uint32_t> thr(size_t(M) * 3);
    synth_thresholds(seed, maf_lo, maf_hi, log_uniform, p_missing, 0, M, thr.data());
    uint32_t *dthr = nullptr;
    int rc = dev_alloc(&dthr, size_t(M) * 3);
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(dthr, thr.data(), sizeof(uint32_t) * 3 * M, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        synth_codes_kernel<<<dim3(c->nCTA, std::min<uint32_t>(M, 4096)), 128, 0, c->stream>>>(seed, dthr, M, c->cfg.n_inds, ind_offset, c->nCTA, c->SL, c->PW, c->codes);
        const uint32_t wpb = 8;
        const size_t n_rows = size_t(M) * c->nCTA;
        build_row_tails_kernel<<<(unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, 0, c->stream>>>(c->codes, 0, n_rows, c->PW);
        count_classes_kernel<<<(M + wpb - 1) / wpb, wpb * 32, 0, c->stream>>>(c->codes, 0, M, c->cfg.n_inds, c->nCTA, c->PW, c->counts_local);
        c->launches += 3;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(dthr);
    if (e != cudaSuccess) return fail(BRR_GPU_ECUDA, "synthetic generation failed: %s", cudaGetErrorString(e));
    c->uploaded = true; c->finalized = false;
    return 0;
}

extern "C" int brr_gpu_synth_bed_host(uint64_t seed, double maf_lo, double maf_hi, int log_uniform, double p_missing, uint32_t n_inds,
                                      uint32_t ind_offset, uint32_t first, uint32_t count, uint8_t *bed_out)
{
    if (!bed_out) return fail(BRR_GPU_EINVAL, "null output");
    const size_t cb = (size_t(n_inds) + 3) / 4;
    std::vector<uint32_t> thr(size_t(count) * 3);
    synth_thresholds(seed, maf_lo, maf_hi, log_uniform, p_missing, first, count, thr.data());
    static const uint8_t plink[4] = {3, 2, 0, 1};   // class 0,1,2,missing -> PLINK 11,10,00,01
    for (uint32_t i = 0; i < count; ++i) {
        uint8_t *col = bed_out + cb * i;
        memset(col, 0, cb);
        for (uint32_t n = 0; n < n_inds; ++n) {
            const uint32_t cls = synth_class(seed, first + i, ind_offset + n, &thr[3 * i]);
            col[n >> 2] |= uint8_t(plink[cls] << (2 * (n & 3)));
        }
    }
    return 0;
}

// ---- introspection -----------------------------------------------------------------------------

extern "C" int brr_gpu_info(brr_gpu_ctx *c, uint32_t *n_ctas, uint32_t *slice_inds, uint32_t *block_markers, uint64_t *genotype_bytes)
{
    if (!c) return fail(BRR_GPU_EINVAL, "null context");
    if (n_ctas) *n_ctas = c->nCTA;
    if (slice_inds) *slice_inds = c->SL;
    if (block_markers) *block_markers = c->B;
    if (genotype_bytes) *genotype_bytes = c->cfg.layout == BRR_GPU_LAYOUT_DENSE2BIT ? c->codes_bytes : c->rag.idx_bytes;
    return 0;
}

extern "C" uint64_t brr_gpu_launch_count(brr_gpu_ctx *c) { return c ? c->launches : 0; }
extern "C" const char *brr_gpu_kernel_version(void) { return kDenseKernelVersion; }

extern "C" int brr_gpu_profile_cycles(brr_gpu_ctx *c, uint64_t *cycles40, int reset)
{
    CHECK_CTX(c);
    if (!cycles40) return fail(BRR_GPU_EINVAL, "null output");
    if (!c->prof) return fail(BRR_GPU_ESTATE, "phase profiling is off: set BRR_GPU_PROFILE=1 before brr_gpu_create");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy(cycles40, c->prof, sizeof(uint64_t) * kProfPhases, cudaMemcpyDeviceToHost));
    if (reset) CUDA_TRY(cudaMemset(c->prof, 0, sizeof(uint64_t) * kProfPhases));
    return 0;
}

extern "C" int brr_gpu_profile_cycles5(brr_gpu_ctx *c, uint64_t *cycles24, int reset)
{
    CHECK_CTX(c);
    if (!cycles24) return fail(BRR_GPU_EINVAL, "null output");
    if (!c->prof5) return fail(BRR_GPU_ESTATE, "phase profiling is off: set BRR_GPU_PROFILE=1 before brr_gpu_create");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy(cycles24, c->prof5, sizeof(uint64_t) * v5::kPfN, cudaMemcpyDeviceToHost));
    if (reset) CUDA_TRY(cudaMemset(c->prof5, 0, sizeof(uint64_t) * v5::kPfN));
    return 0;
}

extern "C" int brr_gpu_profile_cycles6(brr_gpu_ctx *c, uint64_t *cycles32, int reset)
{
    CHECK_CTX(c);
    if (!cycles32) return fail(BRR_GPU_EINVAL, "null output");
    if (!c->prof6) return fail(BRR_GPU_ESTATE, "phase profiling is off: set BRR_GPU_PROFILE=1 before brr_gpu_create");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy(cycles32, c->prof6, sizeof(uint64_t) * v6::kPfN, cudaMemcpyDeviceToHost));
    if (reset) CUDA_TRY(cudaMemset(c->prof6, 0, sizeof(uint64_t) * v6::kPfN));
    return 0;
}

extern "C" int brr_gpu_profile_ctas6(brr_gpu_ctx *c, uint64_t *words, uint32_t max_ctas, uint32_t *n_ctas, int reset)
{
    CHECK_CTX(c);
    if (!words || !n_ctas) return fail(BRR_GPU_EINVAL, "null output");
    if (!c->prof6) return fail(BRR_GPU_ESTATE, "phase profiling is off: set BRR_GPU_PROFILE=1 before brr_gpu_create");
    const uint32_t n = std::min<uint32_t>(max_ctas, c->nCTA);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy(words, c->prof6 + v6::kPfN, sizeof(uint64_t) * 8 * n, cudaMemcpyDeviceToHost));
    if (reset) CUDA_TRY(cudaMemset(c->prof6 + v6::kPfN, 0, sizeof(uint64_t) * 8 * size_t(c->nCTA)));
    *n_ctas = n;
    return 0;
}

extern "C" int brr_gpu_last_sweep_info(brr_gpu_ctx *c, uint32_t *kernel_version, uint32_t *restarts)
{
    if (!c) return fail(BRR_GPU_EINVAL, "null context");
    if (kernel_version) *kernel_version = c->last_kernel;
    if (restarts) *restarts = c->last_restarts;
    return 0;
}

extern "C" void *brr_gpu_stream(brr_gpu_ctx *c) { return c ? (void *)c->stream : nullptr; }

#include "ragged_host.inl"
#include "multigpu_host.inl"
