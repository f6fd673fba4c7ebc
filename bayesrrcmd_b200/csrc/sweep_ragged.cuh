// Ragged / sparse (index-list) layout: device store and persistent sweep kernel (sm_100a).
//
// Semantics of RaggedSparseMarker + RaggedBayesRKernel + SparseBayesRRG (reference
// src/raggedbayesrkernel.cpp:10-35, src/sparsebayesrkernel.cpp:10-34, src/SparseBayesRRG.cpp:76-113),
// and -- with `eigen` set -- of EigenSparseMarker + EigenBayesRKernel (src/eigenbayesrkernel.cpp:10-20):
//     num    = beta_old (N-1) - mean * epsilonSum / sd + (sum_ones eps + 2 sum_twos eps [+ mean sum_missing eps]) / sd
//     eps   += -dB mean/sd everywhere, + dB/sd on ones, + 2 dB/sd on twos, net 0 on missing      (dB = beta_old - beta)
//     epsilonSum += dB Zsum/sd - dB mean N/sd          (tracked analytically, as the reference does)
//
// Device design
//   * As in the dense kernel, CTA c owns individuals [c*SL,(c+1)*SL) and keeps eps in shared memory.
//   * The "-dB mean/sd everywhere" term is a lazily applied scalar offset `off` (eps_true = eps_s + off),
//     so an update touches only the listed individuals (the reference builds a dense N-vector,
//     src/raggedbayesrkernel.cpp:17).  Missing individuals get the compensating +dB mean/sd.
//   * Index lists are stored per (CTA slice, marker) as 16-bit slice-local indices, so every CTA streams
//     only its own contiguous segment: 2 bytes per non-zero genotype.
//   * No Gram correction here (sparse x sparse intersections cost more than they save): a pass computes the
//     dots of a WINDOW of W markers against the current residual, walks them in order and commits up to and
//     including the first marker that changes eps (not skipUpdate, src/BayesRBase.cpp:372); the markers
//     behind it are re-done by the next pass.  The chain is identical to the one-by-one loop; W adapts to
//     the observed update rate.  Asynchronous mode commits the whole window (bounded staleness = W tokens).
#pragma once
#include "common.cuh"
#include "sampler.cuh"

namespace brr {

struct RaggedStore {
    uint16_t *idx16 = nullptr;     // all segments, CTA-major
    uint32_t *seg = nullptr;       // [nCTA][M][4]: start(ones), start(twos), start(missing), end; relative to cta_base
    uint64_t *cta_base = nullptr;  // [nCTA] element offset of the CTA's region in idx16
    double *stats = nullptr;       // [M][4] mean, sd, sqrdZ, Zsum
    size_t idx_bytes = 0;
    size_t n_entries = 0;
    bool stats_given = false;
    // unused legacy fields kept for the destroy list
    uint64_t *offsets = nullptr;
    int32_t *idx = nullptr;
};

struct RaggedSweepParams {
    uint32_t n_local, n_total, M, K, nG, W, nCTA, SL;
    uint32_t eigen;                // 1: EigenBayesRKernel semantics (missing enter the dot with the mean)
    const uint16_t *idx16;
    const uint32_t *seg;
    const uint64_t *cta_base;
    const double *stats;
    const int32_t *group;
    const uint32_t *order;
    uint32_t n_order;
    const double *rand_uz, *pi, *sigmaG, *cva;
    double sigmaE, eps_sum_in;
    uint32_t async_mode;
    double *eps, *beta, *comp, *acum;
    double *part_dot;              // [2][nCTA][kMaxW]
    unsigned *bar;
    double *cta_red;               // [nCTA][2]
    double *v, *betasq, *out_scalars;
    unsigned *out_counts;
};

struct RaggedMeta {
    uint32_t id[kMaxW];
    int32_t grp[kMaxW];
    double beta_old[kMaxW], u[kMaxW], z[kMaxW];
    double mean[kMaxW], sd[kMaxW], zsum[kMaxW];
    uint32_t s0[kMaxW], s1[kMaxW], s2[kMaxW], s3[kMaxW];   // this CTA's segment bounds
};

struct RaggedShared {
    GroupConsts gc;
    RaggedMeta meta;
    double dot_part[kWarps][2];
    double red_part[kWarps][kMaxW];
    double dsum[kMaxW];
    double dbeta[kMaxW];
    double warp_tmp[kWarps];
    double betasq_s[kMaxGK];
    int v_s[kMaxGK];
    double off;          // lazy offset: eps_true = eps_s + off
    double eps_sum;      // tracked m_epsilonSum
    uint32_t n_acc;      // markers committed by this pass
    uint32_t n_nonskipped_total, n_passes;
};

__host__ __device__ inline size_t ragged_smem_bytes(uint32_t SL)
{
    return ((sizeof(RaggedShared) + 15) & ~size_t(15)) + size_t(SL) * sizeof(double) + 16;
}

// RaggedSparseMarkerBuilder::endColumn (reference src/raggedsparsemarkerbuilder.cpp:43-55) on class counts;
// eigen != 0: EigenSparseMarkerBuilder::endColumn (src/eigensparsemarkerbuilder.cpp:40-57).
__global__ void ragged_stats_kernel(const uint32_t *__restrict__ counts, uint32_t M, uint32_t n_total, int eigen,
                                    double *__restrict__ stats)
{
    const uint32_t mk = blockIdx.x * blockDim.x + threadIdx.x;
    if (mk >= M) return;
    const double n1 = counts[4 * size_t(mk) + 1], n2 = counts[4 * size_t(mk) + 2], nm = counts[4 * size_t(mk) + 3];
    const double N = double(n_total);
    double Zsum = n1 + 2.0 * n2, sqrdZ = n1 + 4.0 * n2;
    const double mean = Zsum / N;                                                          // :49 / :46
    if (eigen) { sqrdZ += (mean * mean) * nm; Zsum += mean * nm; }                         // eigen :50,:53
    const double sd = sqrt((sqrdZ - 2.0 * mean * Zsum + N * mean * mean) / (N - 1.0));    // :53-54 / :56-57
    stats[4 * size_t(mk) + 0] = mean; stats[4 * size_t(mk) + 1] = sd;
    stats[4 * size_t(mk) + 2] = sqrdZ; stats[4 * size_t(mk) + 3] = Zsum;
}

// sum over a segment of eps_s[idx]; lanes stride, deterministic warp reduction
__device__ __forceinline__ double seg_gather_sum(const uint16_t *__restrict__ idx, uint32_t a, uint32_t b,
                                                 const double *eps_s, int lane)
{
    // eight index loads in flight per lane (the lists stream from HBM: memory-level parallelism is what bounds this loop),
    // four partial sums in a fixed order
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    uint32_t i = a + lane;
    for (; i + 224 < b; i += 256) {
        const uint16_t j0 = idx[i], j1 = idx[i + 32], j2 = idx[i + 64], j3 = idx[i + 96], j4 = idx[i + 128], j5 = idx[i + 160],
                       j6 = idx[i + 192], j7 = idx[i + 224];
        s0 += eps_s[j0]; s1 += eps_s[j1]; s2 += eps_s[j2]; s3 += eps_s[j3];
        s0 += eps_s[j4]; s1 += eps_s[j5]; s2 += eps_s[j6]; s3 += eps_s[j7];
    }
    for (; i < b; i += 32) s0 += eps_s[idx[i]];
    return warp_sum((s0 + s1) + (s2 + s3));
}

__global__ void __launch_bounds__(kThreads, 1) sweep_ragged_kernel(const RaggedSweepParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RaggedShared &sh = *reinterpret_cast<RaggedShared *>(smem_raw);
    double *eps_s = reinterpret_cast<double *>(smem_raw + ((sizeof(RaggedShared) + 15) & ~size_t(15)));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t cta = blockIdx.x, SL = p.SL, K = p.K;
    const uint32_t slice0 = cta * SL;
    const double NM1 = double(p.n_total - 1), Nd = double(p.n_total);
    const uint16_t *idx = p.idx16 + p.cta_base[cta];
    const uint32_t *segt = p.seg + size_t(cta) * p.M * 4;

    for (uint32_t i = tid; i < SL; i += kThreads) {
        const uint32_t gi = slice0 + i;
        eps_s[i] = (gi < p.n_local) ? p.eps[gi] : 0.0;
    }
    fill_group_consts(sh.gc, p.nG, K, NM1, p.sigmaE, p.pi, p.sigmaG, p.cva);
    for (int i = tid; i < kMaxGK; i += kThreads) { sh.v_s[i] = 0; sh.betasq_s[i] = 0.0; }
    if (tid == 0) { sh.off = 0.0; sh.eps_sum = p.eps_sum_in; sh.n_nonskipped_total = 0; sh.n_passes = 0; sh.n_acc = 0; }
    __syncthreads();

    uint32_t pos = 0, pass = 0;
    uint32_t W = p.W;                       // current window (adapts in synchronous mode)
    while (pos < p.n_order) {
        const uint32_t nb = min(W, p.n_order - pos);
        // ---- metadata of the window ------------------------------------------------------------
        if (uint32_t(tid) < nb) {
            const uint32_t id = p.order[pos + tid];
            RaggedMeta &m = sh.meta;
            m.id[tid] = id; m.grp[tid] = p.group ? p.group[id] : 0;
            m.beta_old[tid] = p.beta[id];
            m.u[tid] = p.rand_uz[2 * size_t(id)]; m.z[tid] = p.rand_uz[2 * size_t(id) + 1];
            const double4 st = *reinterpret_cast<const double4 *>(p.stats + 4 * size_t(id));
            m.mean[tid] = st.x; m.sd[tid] = st.y; m.zsum[tid] = st.w;
            const uint4 s = *reinterpret_cast<const uint4 *>(segt + 4 * size_t(id));
            m.s0[tid] = s.x; m.s1[tid] = s.y; m.s2[tid] = s.z; m.s3[tid] = s.w;
        }
        __syncthreads();
        // ---- partial dots: two warps per marker (ones | twos [+ missing]) -------------------------
        for (uint32_t task = warp; task < 2 * nb; task += kWarps) {
            const uint32_t m = task >> 1, half = task & 1;
            double r;
            if (half == 0) {
                r = seg_gather_sum(idx, sh.meta.s0[m], sh.meta.s1[m], eps_s, lane);
            } else {
                r = 2.0 * seg_gather_sum(idx, sh.meta.s1[m], sh.meta.s2[m], eps_s, lane);
                if (p.eigen) r += sh.meta.mean[m] * seg_gather_sum(idx, sh.meta.s2[m], sh.meta.s3[m], eps_s, lane);
            }
            if (lane == 0) sh.red_part[half][m] = r;      // reuse red_part rows 0/1 as scratch
        }
        __syncthreads();
        if (uint32_t(tid) < nb)
            p.part_dot[(size_t(pass & 1) * p.nCTA + cta) * kMaxW + tid] = sh.red_part[0][tid] + sh.red_part[1][tid];
        grid_arrive(p.bar);
        grid_wait(p.bar, (pass + 1) * p.nCTA);
        {
            const double *pd = p.part_dot + size_t(pass & 1) * p.nCTA * kMaxW;
            double acc = 0.0;
            if (uint32_t(lane) < nb)
                for (uint32_t c = warp; c < p.nCTA; c += kWarps) acc += ld_cg(pd + size_t(c) * kMaxW + lane);
            __syncthreads();                  // red_part scratch rows were read above
            sh.red_part[warp][lane] = acc;
        }
        __syncthreads();
        if (uint32_t(tid) < kMaxW) {
            double s = 0.0;
            for (int w = 0; w < kWarps; ++w) s += sh.red_part[w][tid];
            sh.dsum[tid] = s;
        }
        __syncthreads();
        // ---- walk ---------------------------------------------------------------------------------
        if (warp == 0) {
            const RaggedMeta &m = sh.meta;
            const int k = lane;
            const bool act = uint32_t(k) < nb;
            const double bo = act ? m.beta_old[k] : 0.0;
            const int g = act ? m.grp[k] : 0;
            Draw dr; dr.beta_new = bo; dr.acum0 = 0.0; dr.comp = -1;
            if (act) {
                // missing individuals (eigen) carry the lazy offset too: count them via the imputed Zsum
                const double listed = sh.dsum[k] + sh.off * m.zsum[k];
                const double num = bo * NM1 - m.mean[k] * sh.eps_sum / m.sd[k] + listed / m.sd[k];   // sparsebayesrkernel.cpp:26
                dr = sample_marker_k(sh.gc, g, K, num, p.sigmaE, bo, m.u[k], m.z[k]);
            }
            const bool nonskip = act && !(bo == 0.0 && dr.beta_new == 0.0);
            const unsigned mask = __ballot_sync(0xffffffffu, nonskip);
            const uint32_t n_acc = (p.async_mode || mask == 0) ? nb : uint32_t(__ffs(mask));   // incl. the first updater
            const bool commit = act && uint32_t(k) < n_acc;
            sh.dbeta[lane] = commit ? (bo - dr.beta_new) : 0.0;
            const unsigned cm = __ballot_sync(0xffffffffu, commit && nonskip);
            if (lane == 0) { sh.n_acc = n_acc; sh.n_nonskipped_total += __popc(cm); sh.n_passes += 1; }
            if (cta == 0 && commit) {
                const uint32_t id = m.id[k];
                if (!p.async_mode) p.acum[id] = dr.acum0;
                if (dr.comp >= 0) {
                    p.beta[id] = dr.beta_new;
                    p.comp[id] = double(dr.comp);
                    atomicAdd(&sh.v_s[g * K + dr.comp], 1);
                }
            }
            if (cta == 0) {
                for (uint32_t gi = 0; gi < p.nG; ++gi) {
                    const double b2 = (commit && g == int(gi) && dr.comp > 0) ? dr.beta_new * dr.beta_new : 0.0;
                    const double s = warp_sum(b2);
                    if (lane == 0) sh.betasq_s[gi] += s;
                }
            }
        }
        __syncthreads();
        // ---- residual updates of the committed markers (sync: at most one, the last committed) -------
        const uint32_t n_acc = sh.n_acc;
        for (uint32_t k = 0; k < n_acc; ++k) {
            const double dB = sh.dbeta[k];
            if (dB == 0.0) continue;
            const RaggedMeta &m = sh.meta;
            const double meanAdj = dB * m.mean[k] / m.sd[k];      // raggedbayesrkernel.cpp:15
            const double oneAdj = dB / m.sd[k];                   // :20
            for (uint32_t i = m.s0[k] + tid; i < m.s1[k]; i += kThreads) eps_s[idx[i]] += oneAdj;          // :21
            for (uint32_t i = m.s1[k] + tid; i < m.s2[k]; i += kThreads) eps_s[idx[i]] += 2 * oneAdj;      // :24
            for (uint32_t i = m.s2[k] + tid; i < m.s3[k]; i += kThreads) eps_s[idx[i]] += meanAdj;         // :27
            __syncthreads();
            if (tid == 0) {
                sh.off -= meanAdj;                                                                          // :17
                sh.eps_sum += dB * m.zsum[k] / m.sd[k] - dB * m.mean[k] * Nd / m.sd[k];                     // sparsebayesrkernel.cpp:33
            }
        }
        __syncthreads();
        pos += n_acc;
        ++pass;
        if (!p.async_mode) {                 // adapt the window to the update rate (uniform across CTAs)
            if (n_acc == nb && W < p.W) W = min(p.W, W * 2);
            else if (n_acc * 2 < nb && W > 1) W = max(1u, W / 2);
        }
    }

    // ---- epilogue ----------------------------------------------------------------------------------
    const double off = sh.off;
    double sq = 0.0;
    for (uint32_t i = tid; i < SL; i += kThreads) {
        const uint32_t gi = slice0 + i;
        if (gi < p.n_local) { const double e = eps_s[i] + off; p.eps[gi] = e; sq += e * e; }
    }
    sq = warp_sum(sq);
    if (lane == 0) sh.warp_tmp[warp] = sq;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += sh.warp_tmp[w];
        p.cta_red[2 * cta + 1] = s;
    }
    grid_arrive(p.bar);
    if (cta == 0) {
        grid_wait(p.bar, (pass + 1) * p.nCTA);
        if (warp == 0) {
            double q = 0.0;
            for (uint32_t c = lane; c < p.nCTA; c += 32) q += ld_cg(p.cta_red + 2 * c + 1);
            q = warp_sum(q);
            if (lane == 0) {
                p.out_scalars[0] = sh.eps_sum; p.out_scalars[1] = q;
                p.out_counts[0] = sh.n_nonskipped_total; p.out_counts[1] = sh.n_passes;
            }
        }
        for (uint32_t i = tid; i < p.nG * K; i += kThreads) p.v[i] = double(sh.v_s[i]);
        for (uint32_t i = tid; i < p.nG; i += kThreads) p.betasq[i] = sh.betasq_s[i];
    }
}

// ---- marker-level parity hooks (one block per slice, eps in global memory) -------------------------
__global__ void ragged_dot_kernel(const uint16_t *__restrict__ idx16, const uint32_t *__restrict__ seg,
                                  const uint64_t *__restrict__ cta_base, uint32_t M, uint32_t marker, int eigen, double mean,
                                  const double *__restrict__ eps, uint32_t SL, double *__restrict__ partial)
{
    __shared__ double ws[32];
    const uint32_t cta = blockIdx.x;
    const uint16_t *idx = idx16 + cta_base[cta];
    const uint32_t *s = seg + (size_t(cta) * M + marker) * 4;
    const double *e = eps + size_t(cta) * SL;
    double acc = 0.0;
    for (uint32_t i = s[0] + threadIdx.x; i < s[1]; i += blockDim.x) acc += e[idx[i]];
    for (uint32_t i = s[1] + threadIdx.x; i < s[2]; i += blockDim.x) acc += 2.0 * e[idx[i]];
    if (eigen) for (uint32_t i = s[2] + threadIdx.x; i < s[3]; i += blockDim.x) acc += mean * e[idx[i]];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (uint32_t w = 0; w < blockDim.x / 32; ++w) t += ws[w];
        partial[cta] = t;
    }
}

__global__ void ragged_update_kernel(const uint16_t *__restrict__ idx16, const uint32_t *__restrict__ seg,
                                     const uint64_t *__restrict__ cta_base, uint32_t M, uint32_t marker, double meanAdj,
                                     double oneAdj, double *__restrict__ eps, uint32_t n_local, uint32_t SL)
{
    const uint32_t cta = blockIdx.x;
    const uint16_t *idx = idx16 + cta_base[cta];
    const uint32_t *s = seg + (size_t(cta) * M + marker) * 4;
    double *e = eps + size_t(cta) * SL;
    const uint32_t n = min(SL, n_local - cta * SL);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) e[i] -= meanAdj;
    __syncthreads();
    for (uint32_t i = s[0] + threadIdx.x; i < s[1]; i += blockDim.x) e[idx[i]] += oneAdj;
    for (uint32_t i = s[1] + threadIdx.x; i < s[2]; i += blockDim.x) e[idx[i]] += 2 * oneAdj;
    for (uint32_t i = s[2] + threadIdx.x; i < s[3]; i += blockDim.x) e[idx[i]] += meanAdj;
}

} // namespace brr
