// Persistent cooperative Gibbs-sweep kernel, dense 2-bit layout (sm_100a).
//
// One launch = one AnalysisGraph::exec (reference src/sequential.cpp:10-27 driving
// src/BayesRBase.cpp:287-392 over all markers of the sweep, in the given order).
//
// Design
//   * One CTA per SM, all co-resident (cooperative launch).  CTA c owns the individuals
//     [c*SL, (c+1)*SL) and keeps their residuals eps in SHARED MEMORY for the whole sweep.
//   * Genotypes live in HBM as two bit planes per (marker, CTA slice): P1 = {g==1 or missing},
//     P2 = {g==2 or missing}.  A block of B markers (the next B of the permutation) is fetched with
//     TMA 1-D bulk copies (cp.async.bulk + mbarrier), double buffered, one block ahead.
//   * Per block every CTA computes, from its slice only:
//       - B partial dots x_j . eps   (masked fp64 sums S1,S2,Smiss -> value-table combination),
//       - the B x B integer Gram counts sum_i g_ij g_ik (+ the missing cross terms) by AND+POPC,
//     publishes them (plain stores / integer REDs), and joins ONE grid-wide barrier.
//   * After the barrier every CTA redundantly reduces the partials in a fixed order and one warp
//     walks the block: lane k samples marker k with
//         num_k = x_k.eps_blockstart - sum_{j<k updated} (beta_j,new - beta_j,old) x_j.x_k + beta_old,k (N-1),
//     which is exactly the dot against the residual the one-by-one loop would have seen.  Lanes
//     speculate that no earlier marker of the block changes eps; each marker that does (not
//     skipUpdate, src/BayesRBase.cpp:372) triggers one more SIMT round for the lanes behind it.
//   * Finally every CTA applies eps += (beta_old - beta_new) x_j for the block's updated markers to
//     its shared-memory slice (planes are still resident), and moves on.
//   Asynchronous mode (bounded staleness, src/BayesRBase.cpp:394-523 under src/parallelgraph.cpp):
//   same pipeline with B = tokens and no Gram correction.
#pragma once
#include "common.cuh"
#include "sampler.cuh"

namespace brr {

struct DenseSweepParams {
    // geometry
    uint32_t n_local, n_total, M, K, nG, B, nCTA, SL, PB;
    // store
    const uint8_t *planes;   // [M][nCTA][2][PB]
    const double *lut;       // [M][4] values of class 0,1,2,missing
    const uint32_t *counts;  // [M][4] GLOBAL class counts
    const int32_t *group;    // [M] or nullptr
    // sweep inputs (device)
    const uint32_t *order;
    uint32_t n_order;
    const double *rand_uz;   // [M][2]
    const double *pi;        // [nG][K]
    const double *sigmaG;    // [nG]
    const double *cva;       // [nG][K-1]
    double sigmaE;
    uint32_t async_mode;     // 0 = synchronous
    // chain state
    double *eps, *beta, *comp, *acum;
    // scratch
    double *part_dot;        // [2][nCTA][kMaxB]
    double *part_eps;        // [2][nCTA]
    int *gram;               // [kGramBuffers][kMaxB*kMaxB*4]
    unsigned *bar;           // grid barrier arrival counter (zero at launch)
    double *cta_red;         // [nCTA][2]
    // outputs
    double *v;               // [nG][K]
    double *betasq;          // [nG]
    double *out_scalars;     // [0]=sum eps, [1]=||eps||^2
    unsigned *out_counts;    // [0]=non-skipped markers, [1]=blocks
};

struct BlockMeta {
    uint32_t id[kMaxB];
    int32_t grp[kMaxB];
    double beta_old[kMaxB], u[kMaxB], z[kMaxB];
    double lut[kMaxB][4];
    double sj[kMaxB], cj[kMaxB], ej[kMaxB];   // x = sj * (g - cj + ej*[missing])
    double sumg[kMaxB], nmiss[kMaxB];         // global n1 + 2 n2, global missing count
};

struct SweepShared {   // fixed-size part; eps slice and plane buffers follow (dynamic)
    GroupConsts gc;
    BlockMeta meta[2];
    double G[kMaxB * kMaxB];
    double dot_part[kWarps][4][3];
    double red_part[kWarps][kMaxB + 1];
    double dsum[kMaxB];
    double dbeta[kMaxB];
    double eps_sum_total;
    double eps_sum_cta;
    double warp_tmp[kWarps];
    double betasq_s[kMaxGK];
    int v_s[kMaxGK];
    int gram_s[kMaxB * kMaxB * 4];
    uint32_t ns_list[kMaxB];
    uint32_t n_ns;
    uint32_t n_nonskipped_total;
    alignas(8) uint64_t mbar[2];
};

__host__ __device__ inline size_t sweep_smem_bytes(uint32_t B, uint32_t SL, uint32_t PB)
{
    size_t s = sizeof(SweepShared);
    s = (s + 15) & ~size_t(15);
    s += size_t(SL) * sizeof(double);           // eps slice, transposed [8][SL/8]
    s += size_t(2) * B * 2 * PB;                // two plane buffers
    return s + 16;
}

__device__ __forceinline__ void load_meta_regs(const DenseSweepParams &p, uint32_t pos, uint32_t &id, int &grp,
                                               double &bo, double &u, double &z, double lut[4], uint32_t cnt[4])
{
    id = p.order[pos];
    grp = p.group ? p.group[id] : 0;
    bo = p.beta[id];
    u = p.rand_uz[2 * size_t(id)];
    z = p.rand_uz[2 * size_t(id) + 1];
    const double4 *l4 = reinterpret_cast<const double4 *>(p.lut + 4 * size_t(id));
    const double4 l = *l4;
    lut[0] = l.x; lut[1] = l.y; lut[2] = l.z; lut[3] = l.w;
    const uint4 c = *reinterpret_cast<const uint4 *>(p.counts + 4 * size_t(id));
    cnt[0] = c.x; cnt[1] = c.y; cnt[2] = c.z; cnt[3] = c.w;
}

__device__ __forceinline__ void store_meta(BlockMeta &m, int k, uint32_t id, int grp, double bo, double u, double z,
                                           const double lut[4], const uint32_t cnt[4])
{
    m.id[k] = id; m.grp[k] = grp; m.beta_old[k] = bo; m.u[k] = u; m.z[k] = z;
    m.lut[k][0] = lut[0]; m.lut[k][1] = lut[1]; m.lut[k][2] = lut[2]; m.lut[k][3] = lut[3];
    const double s = 0.5 * (lut[2] - lut[0]);       // value table is affine in the allele count
    const double c = -lut[0] / s;
    m.sj[k] = s; m.cj[k] = c; m.ej[k] = lut[3] / s + c;
    m.sumg[k] = double(cnt[1]) + 2.0 * double(cnt[2]);
    m.nmiss[k] = double(cnt[3]);
}

__global__ void __launch_bounds__(kThreads, 1) sweep_dense_kernel(const DenseSweepParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SweepShared &sh = *reinterpret_cast<SweepShared *>(smem_raw);
    const size_t off_eps = (sizeof(SweepShared) + 15) & ~size_t(15);
    double *eps_s = reinterpret_cast<double *>(smem_raw + off_eps);
    unsigned char *planes_s = smem_raw + off_eps + size_t(p.SL) * sizeof(double);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t cta = blockIdx.x;
    const uint32_t B = p.B, K = p.K, PB = p.PB, SL = p.SL;
    const uint32_t UP8 = SL / 8;                     // 8-individual units (bytes of a plane) per slice
    const uint32_t PW = SL / 32;                     // 32-bit words per plane per slice
    const uint32_t chunk = 2 * PB;                   // bytes per (marker, CTA)
    const double NM1 = double(p.n_total - 1);
    const uint32_t nBlocks = (p.n_order + B - 1) / B;
    const uint32_t slice0 = cta * SL;                // first individual of this slice

    // ---- prologue -----------------------------------------------------------------------------
    // eps slice -> shared, transposed: individual i = 8u + t lives at eps_s[t*UP8 + u]
    double my_sum = 0.0;
    for (uint32_t i = tid; i < SL; i += kThreads) {
        const uint32_t gi = slice0 + i;
        const double e = (gi < p.n_local) ? p.eps[gi] : 0.0;
        eps_s[(i & 7u) * UP8 + (i >> 3)] = e;
        my_sum += e;
    }
    fill_group_consts(sh.gc, p.nG, K, NM1, p.sigmaE, p.pi, p.sigmaG, p.cva);
    for (int i = tid; i < kMaxGK; i += kThreads) { sh.v_s[i] = 0; sh.betasq_s[i] = 0.0; }
    for (int i = tid; i < int(kMaxB * kMaxB * 4); i += kThreads) sh.gram_s[i] = 0;
    if (tid == 0) {
        mbar_init(&sh.mbar[0], 1); mbar_init(&sh.mbar[1], 1);
        mbar_fence_init();
        sh.n_nonskipped_total = 0;
    }
    my_sum = warp_sum(my_sum);
    if (lane == 0) sh.warp_tmp[warp] = my_sum;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += sh.warp_tmp[w];
        sh.eps_sum_cta = s;
    }

    // TMA prefetch of one block of plane chunks (issued by warp 1)
    auto prefetch = [&](uint32_t b) {
        const uint32_t nb = min(B, p.n_order - b * B);
        uint64_t *bar = &sh.mbar[b & 1];
        unsigned char *dst = planes_s + size_t(b & 1) * B * chunk;
        if (lane == 0) mbar_arrive_expect_tx(bar, nb * chunk);
        __syncwarp();
        if (uint32_t(lane) < nb) {
            const uint32_t id = p.order[b * B + lane];
            const uint8_t *src = p.planes + (size_t(id) * p.nCTA + cta) * chunk;
            bulk_g2s(dst + size_t(lane) * chunk, src, chunk, bar);
        }
    };

    if (nBlocks > 0) {
        if (warp == 1) prefetch(0);
        if (warp == 2) {   // metadata of block 0
            const uint32_t nb0 = min(B, p.n_order);
            if (uint32_t(lane) < nb0) {
                uint32_t id, cnt[4]; int grp; double bo, u, z, lut[4];
                load_meta_regs(p, lane, id, grp, bo, u, z, lut, cnt);
                store_meta(sh.meta[0], lane, id, grp, bo, u, z, lut, cnt);
            }
        }
    }
    __syncthreads();

    // thread -> (t,u) positions of the shared eps slice it owns during the residual update
    constexpr int kMaxR = 16;    // SL <= 16*512 = 8192 individuals per CTA
    const int R = int((SL + kThreads - 1) / kThreads);

    const uint32_t NG4 = (B + 3) / 4;                 // groups of 4 markers
    const uint32_t NSEG = max(1u, uint32_t(kWarps) / NG4);
    const uint32_t UPS = (UP8 + NSEG - 1) / NSEG;     // units per segment
    const uint32_t NTP = NG4 * (NG4 + 1) / 2;         // 4x4 tile pairs of the Gram matrix
    const uint32_t NSG = max(1u, (2u * kWarps + NTP - 1) / NTP);   // word segments per tile pair
    const uint32_t WPS = (PW + NSG - 1) / NSG;

    for (uint32_t b = 0; b < nBlocks; ++b) {
        const uint32_t nb = min(B, p.n_order - b * B);
        const uint32_t buf = b & 1;
        BlockMeta &meta = sh.meta[buf];
        const unsigned char *pl = planes_s + size_t(buf) * B * chunk;

        // ---- 1. prefetch block b+1 (planes by TMA, metadata into registers) ---------------------
        uint32_t n_id = 0, n_cnt[4] = {0, 0, 0, 0}; int n_grp = 0; double n_bo = 0, n_u = 0, n_z = 0, n_lut[4] = {0, 0, 0, 0};
        const bool has_next = (b + 1 < nBlocks);
        const uint32_t nb_next = has_next ? min(B, p.n_order - (b + 1) * B) : 0;
        if (has_next) {
            if (warp == 1) prefetch(b + 1);
            if (warp == 2 && uint32_t(lane) < nb_next)
                load_meta_regs(p, (b + 1) * B + lane, n_id, n_grp, n_bo, n_u, n_z, n_lut, n_cnt);
        }

        // ---- 2. wait for this block's planes ---------------------------------------------------
        mbar_wait(&sh.mbar[buf], (b >> 1) & 1);

        // ---- 3. partial dots: warp = (marker group mg, unit segment sg) -------------------------
        {
            const uint32_t mg = warp % NG4, sg = warp / NG4;
            if (sg < NSEG) {
                double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0}, sm[4] = {0, 0, 0, 0};
                const uint32_t u0 = sg * UPS, u1 = min(UP8, u0 + UPS);
                for (uint32_t u = u0 + lane; u < u1; u += 32) {
                    uint32_t q1[4], q2[4], qm[4], anym = 0;
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const uint32_t m = 4 * mg + a;
                        uint32_t h1 = 0, h2 = 0;
                        if (m < nb) { h1 = pl[(2 * m) * PB + u]; h2 = pl[(2 * m + 1) * PB + u]; }
                        q1[a] = h1 & ~h2; q2[a] = h2 & ~h1; qm[a] = h1 & h2; anym |= qm[a];
                    }
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const double e = eps_s[t * UP8 + u];
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            add_if_bit(s1[a], e, q1[a], 1u << t);
                            add_if_bit(s2[a], e, q2[a], 1u << t);
                        }
                    }
                    if (anym) {
#pragma unroll 1
                        for (int a = 0; a < 4; ++a) {
                            uint32_t q = qm[a];
                            while (q) {
                                const int t = __ffs(q) - 1;
                                q &= q - 1;
                                sm[a] += eps_s[t * UP8 + u];
                            }
                        }
                    }
                }
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const double r1 = warp_sum(s1[a]), r2 = warp_sum(s2[a]), r3 = warp_sum(sm[a]);
                    if (lane == 0) { sh.dot_part[warp][a][0] = r1; sh.dot_part[warp][a][1] = r2; sh.dot_part[warp][a][2] = r3; }
                }
            }
        }

        // ---- 4. Gram counts by AND+POPC (synchronous mode only) --------------------------------
        if (!p.async_mode && nb > 1) {
            for (uint32_t task = warp; task < NTP * NSG; task += kWarps) {
                const uint32_t tp = task % NTP, seg = task / NTP;
                uint32_t ta = 0, rem = tp;                 // tp -> (ta <= tb)
                while (rem >= NG4 - ta) { rem -= NG4 - ta; ++ta; }
                const uint32_t tb = ta + rem;
                if (4 * ta >= nb || 4 * tb >= nb) continue;
                int gg[16], gm[16], mgc[16], mm[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { gg[i] = 0; gm[i] = 0; mgc[i] = 0; mm[i] = 0; }
                const uint32_t w0 = seg * WPS, w1 = min(PW, w0 + WPS);
                for (uint32_t w = w0 + lane; w < w1; w += 32) {
                    uint32_t a1[4], a2[4], am[4], b1[4], b2[4], bm[4], anym = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t mj = 4 * ta + i, mk = 4 * tb + i;
                        uint32_t x1 = 0, x2 = 0, y1 = 0, y2 = 0;
                        if (mj < nb) {
                            x1 = reinterpret_cast<const uint32_t *>(pl + (2 * mj) * PB)[w];
                            x2 = reinterpret_cast<const uint32_t *>(pl + (2 * mj + 1) * PB)[w];
                        }
                        if (mk < nb) {
                            y1 = reinterpret_cast<const uint32_t *>(pl + (2 * mk) * PB)[w];
                            y2 = reinterpret_cast<const uint32_t *>(pl + (2 * mk + 1) * PB)[w];
                        }
                        a1[i] = x1 & ~x2; a2[i] = x2 & ~x1; am[i] = x1 & x2;
                        b1[i] = y1 & ~y2; b2[i] = y2 & ~y1; bm[i] = y1 & y2;
                        anym |= am[i] | bm[i];
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            gg[i * 4 + j] += __popc(a1[i] & b1[j]) + 2 * __popc((a1[i] & b2[j]) | (a2[i] & b1[j])) +
                                             4 * __popc(a2[i] & b2[j]);
                    if (anym) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                gm[i * 4 + j] += __popc(a1[i] & bm[j]) + 2 * __popc(a2[i] & bm[j]);
                                mgc[i * 4 + j] += __popc(am[i] & b1[j]) + 2 * __popc(am[i] & b2[j]);
                                mm[i * 4 + j] += __popc(am[i] & bm[j]);
                            }
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t mj = 4 * ta + i, mk = 4 * tb + j;
                        const int r0 = __reduce_add_sync(0xffffffffu, gg[i * 4 + j]);
                        const int r1 = __reduce_add_sync(0xffffffffu, gm[i * 4 + j]);
                        const int r2 = __reduce_add_sync(0xffffffffu, mgc[i * 4 + j]);
                        const int r3 = __reduce_add_sync(0xffffffffu, mm[i * 4 + j]);
                        if (lane == 0 && mj < mk && mk < nb) {
                            int *g = &sh.gram_s[(mj * kMaxB + mk) * 4];
                            if (r0) atomicAdd(g + 0, r0);
                            if (r1) atomicAdd(g + 1, r1);
                            if (r2) atomicAdd(g + 2, r2);
                            if (r3) atomicAdd(g + 3, r3);
                        }
                    }
            }
        }
        __syncthreads();

        // ---- 5. publish this CTA's partials ----------------------------------------------------
        {
            const uint32_t par = b & 1;
            if (uint32_t(tid) < nb) {
                const uint32_t m = tid, mg = m / 4, a = m % 4;
                double S1 = 0, S2 = 0, Sm = 0;
                for (uint32_t sg = 0; sg < NSEG; ++sg) {
                    const uint32_t w = sg * NG4 + mg;
                    S1 += sh.dot_part[w][a][0]; S2 += sh.dot_part[w][a][1]; Sm += sh.dot_part[w][a][2];
                }
                const double *L = meta.lut[m];
                const double y = (L[1] - L[0]) * S1 + (L[2] - L[0]) * S2 + (L[3] - L[0]) * Sm;
                p.part_dot[(size_t(par) * p.nCTA + cta) * kMaxB + m] = y;
            }
            if (tid == 32) p.part_eps[size_t(par) * p.nCTA + cta] = sh.eps_sum_cta;
            if (!p.async_mode && nb > 1) {
                int *gg = p.gram + size_t(b % kGramBuffers) * (kMaxB * kMaxB * 4);
                for (uint32_t i = tid; i < nb * kMaxB * 4; i += kThreads) {
                    const int vv = sh.gram_s[i];
                    if (vv) { atomicAdd(gg + i, vv); sh.gram_s[i] = 0; }
                }
            }
        }

        // ---- 6. the one grid-wide exchange per block --------------------------------------------
        grid_arrive(p.bar);
        grid_wait(p.bar, (b + 1) * p.nCTA);

        // ---- 7. reduce partials in a fixed order (identical in every CTA) ------------------------
        {
            const uint32_t par = b & 1;
            const double *pd = p.part_dot + size_t(par) * p.nCTA * kMaxB;
            const double *pe = p.part_eps + size_t(par) * p.nCTA;
            // all loads first (independent, L2 latency paid once), then a fixed-order sum
            constexpr int kMaxPerWarp = 12;          // nCTA <= 16*12 = 192
            double vals[kMaxPerWarp];
#pragma unroll
            for (int i = 0; i < kMaxPerWarp; ++i) {
                const uint32_t c = warp + i * kWarps;
                vals[i] = (c < p.nCTA && uint32_t(lane) < nb) ? ld_cg(pd + size_t(c) * kMaxB + lane) : 0.0;
            }
            const double ev = (uint32_t(tid) < p.nCTA) ? ld_cg(pe + tid) : 0.0;
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < kMaxPerWarp; ++i) acc += vals[i];
            sh.red_part[warp][lane] = acc;
            const double es = warp_sum(ev);
            if (lane == 0) sh.red_part[warp][kMaxB] = es;
            if (cta == 0 && !p.async_mode) {   // recycle the Gram buffer block b+3 will accumulate into
                int *gz = p.gram + size_t((b + 3) % kGramBuffers) * (kMaxB * kMaxB * 4);
                for (uint32_t i = tid; i < kMaxB * kMaxB * 4; i += kThreads) gz[i] = 0;
            }
        }
        __syncthreads();
        if (uint32_t(tid) <= kMaxB) {
            double s = 0.0;
            for (int w = 0; w < kWarps; ++w) s += sh.red_part[w][tid];
            if (tid == kMaxB) sh.eps_sum_total = s; else sh.dsum[tid] = s;
        }
        if (!p.async_mode && nb > 1) {
            const int *gg = p.gram + size_t(b % kGramBuffers) * (kMaxB * kMaxB * 4);
            const double Nt = double(p.n_total);
            for (uint32_t pr = tid; pr < nb * nb; pr += kThreads) {
                const uint32_t j = pr / nb, k = pr % nb;
                if (j < k) {
                    const int4 c = __ldcg(reinterpret_cast<const int4 *>(gg + (j * kMaxB + k) * 4));
                    const double cj = meta.cj[j], ck = meta.cj[k], ej = meta.ej[j], ek = meta.ej[k];
                    double t = double(c.x) - ck * meta.sumg[j] - cj * meta.sumg[k] + Nt * cj * ck;
                    t += ek * double(c.y) + ej * double(c.z) - cj * ek * meta.nmiss[k] - ck * ej * meta.nmiss[j] +
                         ej * ek * double(c.w);
                    const double g = meta.sj[j] * meta.sj[k] * t;
                    sh.G[j * kMaxB + k] = g;
                    sh.G[k * kMaxB + j] = g;
                }
            }
        }
        __syncthreads();

        // ---- 8. the walk: one warp, lane = marker of the block ----------------------------------
        if (warp == 0) {
            const int k = lane;
            const bool act = uint32_t(k) < nb;
            const double bo = act ? meta.beta_old[k] : 0.0;
            const double uu = act ? meta.u[k] : 2.0, zz = act ? meta.z[k] : 0.0;
            const int g = act ? meta.grp[k] : 0;
            const double d0 = act ? (meta.lut[k][0] * sh.eps_sum_total + sh.dsum[k]) : 0.0;
            double corr = 0.0;
            Draw dr; dr.beta_new = bo; dr.acum0 = 0.0; dr.comp = -1;
            uint32_t done = 0;
            while (true) {
                const bool live = act && uint32_t(k) >= done;
                if (live) dr = sample_marker_k(sh.gc, g, K, d0 + corr + bo * NM1, p.sigmaE, bo, uu, zz);
                const bool ns = live && !(bo == 0.0 && dr.beta_new == 0.0);
                const unsigned mask = __ballot_sync(0xffffffffu, ns);
                if (p.async_mode || mask == 0) break;
                const int js = __ffs(mask) - 1;
                const double dB = __shfl_sync(0xffffffffu, bo - dr.beta_new, js);   // eps += dB * x_js
                if (k > js && act) corr += dB * sh.G[js * kMaxB + k];
                done = js + 1;
                if (done >= nb) break;
            }
            const double dB = act ? (bo - dr.beta_new) : 0.0;
            const bool upd = act && dB != 0.0;
            const bool nonskip = act && !(bo == 0.0 && dr.beta_new == 0.0);
            if (act) sh.dbeta[k] = dB;
            const unsigned um = __ballot_sync(0xffffffffu, upd);
            if (upd) sh.ns_list[__popc(um & ((1u << k) - 1u))] = k;
            const unsigned nsm = __ballot_sync(0xffffffffu, nonskip);
            if (lane == 0) { sh.n_ns = __popc(um); sh.n_nonskipped_total += __popc(nsm); }
            if (cta == 0 && act) {
                const uint32_t id = meta.id[k];
                if (!p.async_mode) p.acum[id] = dr.acum0;                  // src/BayesRBase.cpp:341
                if (dr.comp >= 0) {
                    p.beta[id] = dr.beta_new;                              // :350-352 / :492
                    p.comp[id] = double(dr.comp);                          // :356 / :493
                    atomicAdd(&sh.v_s[g * K + dr.comp], 1);                // :355 / :511
                }
            }
            if (cta == 0) {
                for (uint32_t gi = 0; gi < p.nG; ++gi) {                   // :353 / :522
                    const double b2 = (act && g == int(gi) && dr.comp > 0) ? dr.beta_new * dr.beta_new : 0.0;
                    const double s = warp_sum(b2);
                    if (lane == 0) sh.betasq_s[gi] += s;
                }
            }
        }
        __syncthreads();

        // ---- 9. residual update on the shared-memory slice -------------------------------------
        {
            const uint32_t n_ns = sh.n_ns;
            double lsum = 0.0;
            for (int r = 0; r < R && r < kMaxR; ++r) {
                const uint32_t pos = tid + r * kThreads;          // position in eps_s
                if (pos >= SL) break;
                const uint32_t t = pos / UP8, u = pos - t * UP8;
                const uint32_t gi = slice0 + 8 * u + t;
                double e = eps_s[pos];
                if (gi < p.n_local) {
                    for (uint32_t q = 0; q < n_ns; ++q) {
                        const uint32_t m = sh.ns_list[q];
                        const uint32_t h1 = pl[(2 * m) * PB + u], h2 = pl[(2 * m + 1) * PB + u];
                        const uint32_t code = ((h1 >> t) & 1u) | (((h2 >> t) & 1u) << 1);
                        e += sh.dbeta[m] * meta.lut[m][code];             // eps += (beta_old-beta) * x
                    }
                    if (n_ns) eps_s[pos] = e;
                    lsum += e;
                }
            }
            lsum = warp_sum(lsum);
            if (lane == 0) sh.warp_tmp[warp] = lsum;
        }
        // metadata of the next block: registers -> shared (loads were issued at the top)
        if (has_next && warp == 2 && uint32_t(lane) < nb_next)
            store_meta(sh.meta[buf ^ 1], lane, n_id, n_grp, n_bo, n_u, n_z, n_lut, n_cnt);
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < kWarps; ++w) s += sh.warp_tmp[w];
            sh.eps_sum_cta = s;
        }
        __syncthreads();
    }

    // ---- epilogue: eps back to HBM, sweep scalars ------------------------------------------------
    double sq = 0.0;
    for (uint32_t i = tid; i < SL; i += kThreads) {
        const uint32_t gi = slice0 + i;
        if (gi < p.n_local) {
            const double e = eps_s[(i & 7u) * UP8 + (i >> 3)];
            p.eps[gi] = e;
            sq += e * e;
        }
    }
    sq = warp_sum(sq);
    if (lane == 0) sh.warp_tmp[warp] = sq;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += sh.warp_tmp[w];
        p.cta_red[2 * cta + 0] = sh.eps_sum_cta;
        p.cta_red[2 * cta + 1] = s;
    }
    grid_arrive(p.bar);
    if (cta == 0) {
        grid_wait(p.bar, (nBlocks + 1) * p.nCTA);
        if (warp == 0) {
            double s = 0.0, q = 0.0;
            for (uint32_t c = lane; c < p.nCTA; c += 32) { s += ld_cg(p.cta_red + 2 * c); q += ld_cg(p.cta_red + 2 * c + 1); }
            s = warp_sum(s); q = warp_sum(q);
            if (lane == 0) {
                p.out_scalars[0] = s; p.out_scalars[1] = q;
                p.out_counts[0] = sh.n_nonskipped_total; p.out_counts[1] = nBlocks;
            }
        }
        for (uint32_t i = tid; i < p.nG * K; i += kThreads) p.v[i] = double(sh.v_s[i]);
        for (uint32_t i = tid; i < p.nG; i += kThreads) p.betasq[i] = sh.betasq_s[i];
    }
}

} // namespace brr
