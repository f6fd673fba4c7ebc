// Persistent cooperative Gibbs-sweep kernel, dense 2-bit layout (sm_100a), version 5: the steady-state specialist.
//
// One launch = one AnalysisGraph::exec (reference src/sequential.cpp:10-27 driving src/BayesRBase.cpp:287-392 over all
// markers of the sweep, in the given order), single GPU, synchronous mode.  Same arithmetic as version 4 (sweep_dense.cuh:
// exact integer dots of the 2-bit codes against 8 int8 limbs of the fixed-point residual on mma.sync.m16n8k32, integer
// reductions over CTAs, fp64 sampling), re-organised so that NOTHING in a CTA waits for anything else in the common case:
//
//   * warps 14,15 (producers): plan blocks of <= 64 consecutive markers of the permutation (ids, beta_old, (u, z), group
//     into a ring of block descriptors) and stream their code rows with TMA bulk copies into a ring of 8 stages of 16 rows;
//     a stage is re-armed the moment the three dot warps of its tile have released it, so ~2 blocks of rows are in flight.
//   * warps 2..13 (dot warps, 3 per 16-marker tile): own the residual (fp64 registers + int8 limbs in shared memory),
//     take the integer dots of a tile, its missing-genotype sums, add them to the block's global slot (one slot per block
//     out of a ring of 512: nothing is recycled on the critical path) and arrive at the slot's counter.  They run AHEAD of
//     the sampling by up to kD blocks.
//   * warps 0,1 (walkers, one thread per marker): wait for a slot's counter, read the reduced sums back, sample.
//
//   Run-ahead without ever repeating a dot pass: a marker that changes the residual is, in the steady state of a chain,
//   almost always a marker that is ALREADY in the model (beta_old != 0) -- and those are known when a block is planned.
//   Their code rows are copied to a small stash, and every dot pass also takes the integer Gram counts of the stashed rows of
//   the last kD blocks and of its own block against its 64 columns (one more MMA N-tile fed by the same A fragments).  The walk
//   of block c then corrects its dots by  sum_j (beta_old_j - beta_j) x_j.x_k  for the rows j of blocks c-kD..c-1 (the updates the
//   dot pass had not seen) and, inside the block, in marker order -- exactly the dot the one-by-one loop would have seen.  The
//   dot warps apply the updates of block c-kD-1 to the residual before the dot pass of block c: a fixed lag, so that every CTA
//   (and every run) agrees on which rows a slot's Gram lines mean.
//   A marker that ENTERS the model (beta_old == 0, beta_new != 0) has no stashed row.  The walk then ends its block right
//   after that marker, and the pipeline restarts behind it (the planned / dotted blocks after it are dropped; their slots are
//   never read).  Rare in the steady state (a few hundred per sweep at N=500k x M=1M); a sweep that starts from beta = 0 should use
//   version 4, whose on-demand Gram rows handle dense updates better (brr_gpu.cu chooses per sweep).
#pragma once
#include "sweep_dense.cuh"

namespace brr {
namespace v5 {

constexpr int kD = 3;                       // the dot pass of block c sees the residual updates of blocks <= c - kD - 1
constexpr int kRowsPerBlock = 2;            // in-model markers of a block (a third one ends the block before it)
constexpr int kGramRows = (kD + 1) * kRowsPerBlock;   // rows whose Gram counts ride on a dot pass (<= 8 = one MMA N-tile)
static_assert(kGramRows <= 8, "one N-tile of Gram rows");
constexpr int kSeq = 8;                     // blocks in flight inside a CTA (descriptor / metadata ring)
constexpr int kStages = 8;                  // stages of 16 code rows; tile t of block seq s uses stage (4 s + t) % 8
constexpr int kTileRows = 16;
constexpr int kWalkW = 2, kDotW = 12;       // warps 0,1 walk; 2..13 dot; 14,15 produce
constexpr int kDotThreads = 32 * kDotW;
constexpr int kP0 = kWalkW + kDotW, kP1 = kP0 + 1;
constexpr int kStash = kSeq * kRowsPerBlock + 1;      // stash rows: (seq % 8) * 2 + i, + one for a marker entering the model
constexpr uint32_t kSlotRing = 512;         // global slots; slot of block c = c % 512, zeroed again 256 blocks ahead of its next use
constexpr uint32_t kSkip = 16;              // block indices skipped by a restart (more than any CTA can have planned ahead)
constexpr uint32_t kEnd = 0xffffffffu;
constexpr int kLCnt = 0;                    // slot lines (128 bytes each, first 16 bytes used): arrival counter
constexpr int kLDots = 1;                   // [64][2] combined limb sums
constexpr int kLSm = kLDots + 2 * kB;       // [64][2] sums of q over the row's missing individuals
constexpr int kLEps = kLSm + 2 * kB;        // [2] sum of q over all individuals
constexpr int kLGram = kLEps + 2;           // [8 rows][64 columns]: u64 #0 = a | cw << 32, u64 #1 = cy | cz << 32 (gram_entry)
constexpr int kSlotLines = kLGram + 8 * kB + 1;
constexpr size_t kSlotInts5 = size_t(kSlotLines) * kLine;
enum { kActSkip = 0, kActDot = 1, kActExit = 2 };
enum { kPfWPlan = 0, kPfWBar, kPfWRead, kPfWSample, kPfWFix, kPfWFile, kPfDGate, kPfDApply, kPfDTma, kPfDLoop, kPfDPost, kPfDComb,
       kPfBlocks, kPfGramBlocks, kPfRestarts, kPfApplies, kPfP0Plan, kPfP0Issue, kPfN = 24 };

struct Sweep5Params {
    uint32_t n_local, n_total, M, K, nG, nCTA, SL, PW, RS;
    const uint32_t *codes; const double *lut; const uint32_t *counts; const int32_t *group;
    const uint32_t *order; uint32_t n_order;
    const double *rand_uz, *pi, *sigmaG, *cva;
    double sigmaE;
    double *eps, *beta, *comp, *acum;
    int *slots;                  // [kSlotRing][kSlotInts5], zero at launch
    unsigned long long *absmax;  // zero at launch
    unsigned *bar;               // prologue / epilogue grid barrier counter, zero at launch
    double *cta_red;             // [nCTA][2]
    double *part_bsq;            // [nCTA][2][nG]   per-CTA, per-walker-warp sums of beta^2
    int *part_v;                 // [nCTA][nG*K]
    double *v, *betasq, *out_scalars;
    unsigned *out_counts;        // [0] non-skipped, [1] blocks walked, [2] error flags, [3] restarts
    unsigned long long *prof;    // optional [kPfN]
};

struct alignas(16) Meta5 {
    uint32_t id[kB];
    int32_t grp[kB];
    double beta_old[kB];
    double uz[kB][2];
};
struct RowConst { double sj, cj, ej, sumg, nmiss; };
struct Blk5 {
    uint32_t c, pos, n, epoch, nrows, rowk[kRowsPerBlock];
    uint32_t back_seq[kD];       // sequencer: seq of the valid blocks 1..kD back in the chain (kEnd = none)
    uint32_t present;            // sequencer: bit r set = Gram row r = (kD - back) * 2 + i is in this block's slot
};
struct Decision { uint32_t action, n_upd, seq_dummy, pad; };

struct Shared5 {
    GroupConsts gc;
    Meta5 meta[kSeq];
    Blk5 blk[kSeq];
    RowConst rowc[kSeq][kRowsPerBlock];
    double rowlut[kSeq][kRowsPerBlock][4];
    double dbeta[kSeq][kRowsPerBlock];
    // control words (volatile accesses)
    uint32_t planned_seq, walk_done_seq, applied_seq, epoch_w, finished, abort;
    uint32_t rs_pos, rs_c, rs_seq, ne_id, ne_uses;
    double ne_dbeta, ne_lut[4];
    Decision dec[2];
    uint32_t upd_row[12]; double upd_dB[12]; double upd_lut[12][4];
    // sequencer's chain (one thread)
    uint32_t hist[kD + 1], hist_n, seq_epoch;
    unsigned long long wq[kDotW][2], sumq[2];
    uint32_t pubcnt[kSeq], tilecnt[12];
    double scale, inv_scale, warp_tmp[kWarps];
    // walkers
    double d0[kB], G[kGramRows][kB], bcast[2];
    uint32_t w_flag[4];
    int v_s[kMaxGK];
    double betasq_w[2][kMaxGK];
    uint32_t n_nonskipped, n_blocks, n_restarts, err;
    unsigned long long prof[kPfN];
    alignas(8) uint64_t full[kStages], empty[kStages], stash_full[kSeq], ne_full;
};

__host__ __device__ inline size_t sweep5_smem_bytes(uint32_t RS)
{
    size_t s = (sizeof(Shared5) + 15) & ~size_t(15);
    s += size_t(kLimbs) * RS * 16;                    // limbs
    s += size_t(kStages) * kTileRows * RS * 4;        // code stages
    s += size_t(kStash) * RS * 4;                     // stashed rows
    s += size_t(3) * 4 * 256 * 4;                      // tile accumulators (three sets in rotation)
    return s + 16;
}

__device__ __forceinline__ uint32_t ldv(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void stv(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }
__device__ __forceinline__ void bar_named(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ uint32_t atom_acqrel_cta_add(uint32_t *p, uint32_t v)
{
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ double gram_std(const RowConst &rj, const RowConst &rk, double a, double cw, double cy, double cz, double Nt)
{   // gram_entry of sweep_dense.cuh on plain constants: x_j . x_k from the integer counts of the pair
    a = a - 3.0 * cy - 3.0 * cz - 9.0 * cw;
    double t = a - rk.cj * rj.sumg - rj.cj * rk.sumg + Nt * rj.cj * rk.cj;
    t += rk.ej * cy + rj.ej * cz - rj.cj * rk.ej * rk.nmiss - rk.cj * rj.ej * rj.nmiss + rj.ej * rk.ej * cw;
    return rj.sj * rk.sj * t;
}

// the slow half of a guarded spin (every 1024 polls): has the launch been given up, or has this wait lasted ~2 s?
__device__ __noinline__ bool spin_check(uint32_t *abort_flag, unsigned *global_flags, long long &t0)
{
    if (ldv(abort_flag)) return true;
    const long long t = clock64();
    if (t0 == 0) { t0 = t; return false; }
    if (t - t0 > 4000000000LL || (*reinterpret_cast<volatile unsigned *>(global_flags) & 4u)) {
        stv(abort_flag, 1u); atomicOr(global_flags, 4u);
        return true;
    }
    return false;
}

template <bool PROF>
__global__ void __launch_bounds__(kThreads, 1) sweep_dense5_kernel(const Sweep5Params p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared5 &sh = *reinterpret_cast<Shared5 *>(smem_raw);
    const uint32_t RS = p.RS, PW = p.PW, SL = p.SL;
    const size_t off0 = (sizeof(Shared5) + 15) & ~size_t(15);
    uint4 *limbs = reinterpret_cast<uint4 *>(smem_raw + off0);                                      // [8][RS]
    uint32_t *stages = reinterpret_cast<uint32_t *>(smem_raw + off0 + size_t(kLimbs) * RS * 16);     // [8][16][RS]
    uint32_t *stash = stages + size_t(kStages) * kTileRows * RS;                                    // [kStash][RS]
    int *tile_acc = reinterpret_cast<int *>(stash + size_t(kStash) * RS);                           // [3][4 tiles][128 dots + 128 Gram]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t gid = lane >> 2, tig = lane & 3;
    const uint32_t cta = blockIdx.x, K = p.K;
    const uint32_t NP = SL / 64, NI = SL / 4, NF = NP / 2;
    const uint32_t PWT = PW + kTail, CB = PWT * 4;
    const double NM1 = double(p.n_total - 1), Nt = double(p.n_total);
    const uint32_t slice0 = cta * SL;
    const bool is_dot = warp >= kWalkW && warp < kP0;
    const int dt = tid - 32 * kWalkW;                 // dot-thread index 0..383
    const uint32_t dw = uint32_t(warp - kWalkW);      // dot-warp index 0..11
    const bool prof_on = PROF && p.prof != nullptr && cta == 0 && (warp == 0 || warp == kWalkW || warp == kP0);   // one warp per role
    long long t_prev = prof_on ? clock64() : 0;
    auto PHASE = [&](int ph) {                        // lane 0 of the calling warp charges the cycles since its previous mark
        if (PROF && prof_on && lane == 0) { const long long t = clock64(); sh.prof[ph] += (unsigned long long)(t - t_prev); t_prev = t; }
    };
    auto PMARK = [&]() { if (PROF && prof_on && lane == 0) t_prev = clock64(); };
    // Every wait of the pipeline is a spin; a wait that lasts ~2 s means a protocol error (or a lost CTA): give up everywhere -- all
    // other waits of the grid see the flag and fall through, the launch ends, the host reports error flag 4 -- instead of hanging the GPU.
    long long spin_t0 = 0; uint32_t spin_it = 0;
    auto spin_fail = [&]() -> bool {
        __nanosleep(20);                               // a spinning warp must not take issue slots / shared-memory cycles from the dot warps
        if ((++spin_it & 1023u) != 0u) return false;
        return spin_check(&sh.abort, p.out_counts + 2, spin_t0);
    };
    auto spin_ok = [&]() { spin_t0 = 0; spin_it = 0; };
    auto mbar_wait_g = [&](uint64_t *bar, unsigned parity) {
        unsigned ok;
        for (;;) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
            if (ok || spin_fail()) break;
        }
        spin_ok();
    };

    // ---- prologue -----------------------------------------------------------------------------------------------------
    // residual -> registers of the dot warps: item = 4w + c owns the individuals 16w + 4t + c, t = 0..3 (as in version 4)
    double e[3][4];
    double amax = 0.0;
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const uint32_t item = uint32_t(dt) + it * kDotThreads;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint32_t i = 16 * (item >> 2) + 4 * t + (item & 3u);
            const uint32_t gi = slice0 + i;
            e[it][t] = (is_dot && item < NI && gi < p.n_local) ? p.eps[gi] : 0.0;
            amax = fmax(amax, fabs(e[it][t]));
        }
    }
    fill_group_consts(sh.gc, p.nG, K, NM1, p.sigmaE, p.pi, p.sigmaG, p.cva);
    for (int i = tid; i < kMaxGK; i += kThreads) { sh.v_s[i] = 0; sh.betasq_w[0][i] = 0.0; sh.betasq_w[1][i] = 0.0; }
    for (int i = tid; i < 3 * 4 * 256; i += kThreads) tile_acc[i] = 0;
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], 3); }
        for (int i = 0; i < kSeq; ++i) { mbar_init(&sh.stash_full[i], 1); sh.pubcnt[i] = 0; }
        for (int i = 0; i < 12; ++i) sh.tilecnt[i] = 0;
        mbar_init(&sh.ne_full, 1);
        mbar_fence_init();
        sh.abort = 0; sh.planned_seq = 0; sh.walk_done_seq = 0; sh.applied_seq = 0; sh.epoch_w = 0; sh.finished = 0; sh.ne_uses = 0;
        sh.hist_n = 0; sh.seq_epoch = 0;
        sh.n_nonskipped = 0; sh.n_blocks = 0; sh.n_restarts = 0; sh.err = 0;
        if (PROF) for (int i = 0; i < kPfN; ++i) sh.prof[i] = 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) sh.warp_tmp[warp] = amax;
    __syncthreads();
    if (tid == 0) {
        double a = 0.0;
        for (int w = 0; w < kWarps; ++w) a = fmax(a, sh.warp_tmp[w]);
        atomicMax(p.absmax, (unsigned long long)__double_as_longlong(a));
    }
    grid_arrive(p.bar);
    grid_wait(p.bar, p.nCTA);
    {   // common fixed-point scale: |q| <= 2^(62 - headroom) at the start of the sweep
        const double a = __longlong_as_double((long long)__ldcg(p.absmax));
        int ex = 0;
        if (a > 0.0) frexp(a, &ex);
        const int s = 62 - kHeadroomBits - ex;
        if (tid == 0) { sh.scale = ldexp(1.0, s); sh.inv_scale = ldexp(1.0, -s); }
    }
    __syncthreads();
    const double scale = sh.scale, inv_scale = sh.inv_scale;

    // residual registers -> limbs in shared memory and this slice's sum of q (dot threads only; version 4's layout)
    auto quantize = [&]() {
        long long shi = 0; unsigned long long slo = 0;
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const uint32_t item = uint32_t(dt) + it * kDotThreads;
            if (item < NI) {
                uint32_t lo[4], hi[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const double x = e[it][t] * scale;
                    if (!(fabs(x) < 4.0e18)) sh.err = 1u;                       // 2^62 = 4.6e18
                    const long long q = __double2ll_rn(x);
                    shi += q >> 32; slo += (unsigned long long)(q & 0xffffffffLL);
                    const unsigned long long uq = ((unsigned long long)q + 0x8080808080808080ULL) ^ 0x8080808080808080ULL;
                    lo[t] = (uint32_t)uq; hi[t] = (uint32_t)(uq >> 32);
                }
                const uint32_t w = item >> 2, c = item & 3u;
                uint32_t *dst = reinterpret_cast<uint32_t *>(limbs) + c;
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    const uint32_t sel = uint32_t(l) | (uint32_t(4 + l) << 4);
                    const uint32_t x01 = __byte_perm(lo[0], lo[1], sel), x23 = __byte_perm(lo[2], lo[3], sel);
                    dst[(size_t(l) * RS + w) * 4] = __byte_perm(x01, x23, 0x5410);
                    const uint32_t y01 = __byte_perm(hi[0], hi[1], sel), y23 = __byte_perm(hi[2], hi[3], sel);
                    dst[(size_t(4 + l) * RS + w) * 4] = __byte_perm(y01, y23, 0x5410);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            shi += __shfl_xor_sync(0xffffffffu, shi, o);
            slo += __shfl_xor_sync(0xffffffffu, slo, o);
        }
        if (lane == 0) { sh.wq[dw][0] = (unsigned long long)shi; sh.wq[dw][1] = slo; }
    };
    auto sum_wq = [&]() {             // one dot thread, after a barrier among the dot warps
        unsigned long long h = 0, l = 0;
        for (int w = 0; w < kDotW; ++w) { h += sh.wq[w][0]; l += sh.wq[w][1]; }
        sh.sumq[0] = h; sh.sumq[1] = l;
    };
    if (is_dot) quantize();
    __syncthreads();
    if (tid == 32 * kWalkW) sum_wq();
    __syncthreads();

    // =====================================================================================================================
    if (warp == kP0) {
        // ---- planner + TMA issue of tiles 0,1 + stashed rows ------------------------------------------------------------
        uint32_t epoch = 0, pos = 0, next_c = 0;
        uint32_t idA = 0, idB = 0;                              // marker ids of positions pos + lane, pos + 32 + lane
        auto load_ids = [&]() {
            idA = pos + lane < p.n_order ? p.order[pos + lane] : 0u;
            idB = pos + 32 + lane < p.n_order ? p.order[pos + 32 + lane] : 0u;
        };
        auto gather = [&](uint32_t s) {                        // asynchronous gather of block s's per-marker inputs
            Meta5 &m = sh.meta[s & (kSeq - 1)];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t k = lane + 32 * h, id = h ? idB : idA;
                m.id[k] = id;
                if (pos + k < p.n_order) {
                    // the walk reads the marker's value table and class counts straight from L2: pull the lines in now
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.lut + 4 * size_t(id)));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.counts + 4 * size_t(id)));
                    cp_async16(&m.uz[k][0], p.rand_uz + 2 * size_t(id));
                    cp_async8(&m.beta_old[k], p.beta + id);
                    if (p.group) cp_async4(&m.grp[k], p.group + id); else m.grp[k] = 0;
                } else { m.beta_old[k] = 0.0; m.grp[k] = 0; }
            }
        };
        load_ids();
        gather(0);
        for (uint32_t s = 0;; ++s) {
            PMARK();
            // the ring entries of seq s are those of seq s - 8: both its walk and the application of its updates must be over
            if (s >= uint32_t(kSeq)) {
                while (ldv(&sh.walk_done_seq) + kSeq <= s || ldv(&sh.applied_seq) + kSeq <= s) {
                    if (ldv(&sh.finished) || spin_fail()) break;
                }
                spin_ok();
            }
            if (ldv(&sh.finished) || ldv(&sh.abort)) break;
            const uint32_t ew = ldv(&sh.epoch_w);
            if (ew != epoch) {                                  // the walk restarted: plan again from its position
                __threadfence_block();
                epoch = ew; pos = ldv(&sh.rs_pos); next_c = ldv(&sh.rs_c);
                cp_async_wait_all();
                load_ids();
                gather(s);
            }
            cp_async_wait_all();
            __syncwarp();
            Meta5 &m = sh.meta[s & (kSeq - 1)];
            Blk5 &B = sh.blk[s & (kSeq - 1)];
            const uint32_t rem = p.n_order - pos;
            uint32_t n = min(uint32_t(kB), rem);
            const unsigned im0 = __ballot_sync(0xffffffffu, uint32_t(lane) < n && m.beta_old[lane] != 0.0);
            const unsigned im1 = __ballot_sync(0xffffffffu, uint32_t(32 + lane) < n && m.beta_old[32 + lane] != 0.0);
            unsigned long long im = (unsigned long long)im0 | ((unsigned long long)im1 << 32);
            uint32_t rowk[kRowsPerBlock] = {0u, 0u}, nrows = 0;
            for (; im && nrows < uint32_t(kRowsPerBlock); ++nrows) { rowk[nrows] = uint32_t(__ffsll((long long)im) - 1); im &= im - 1; }
            if (im) n = uint32_t(__ffsll((long long)im) - 1);          // a third in-model marker: the block ends before it
            const bool end = rem == 0;
            if (lane == 0) {
                B.c = end ? kEnd : next_c; B.pos = pos; B.n = n; B.epoch = epoch; B.nrows = nrows;
                B.rowk[0] = rowk[0]; B.rowk[1] = rowk[1];
            }
            __syncwarp();
            __threadfence_block();
            if (lane == 0) stv(&sh.planned_seq, s + 1);
            PHASE(kPfP0Plan);
            // ids of the next block (latency hidden behind the TMA issue)
            uint32_t idA_n = 0, idB_n = 0;
            const uint32_t pos_n = pos + n;
            if (!end) {
                idA_n = pos_n + lane < p.n_order ? p.order[pos_n + lane] : 0u;
                idB_n = pos_n + 32 + lane < p.n_order ? p.order[pos_n + 32 + lane] : 0u;
            }
            // stage ring: tile t of seq s -> stage (4 s + t) % 8, use number s / 2
#pragma unroll
            for (uint32_t t = 0; t < 2; ++t) {
                const uint32_t st = (4 * s + t) & (kStages - 1);
                if (s >= 2) mbar_wait_g(&sh.empty[st], ((s >> 1) - 1) & 1u);
                const uint32_t nt = n > 16 * t ? min(16u, n - 16 * t) : 0u;
                fence_proxy_async();
                if (lane == 0) mbar_arrive_expect_tx(&sh.full[st], nt * CB);
                __syncwarp();
                if (uint32_t(lane) < nt)
                    bulk_g2s(stages + (size_t(st) * kTileRows + lane) * RS, p.codes + (size_t(m.id[16 * t + lane]) * p.nCTA + cta) * PWT, CB, &sh.full[st]);
            }
            {   // stashed rows of the block's in-model markers
                if (lane == 0) mbar_arrive_expect_tx(&sh.stash_full[s & (kSeq - 1)], nrows * CB);
                __syncwarp();
                if (uint32_t(lane) < nrows)
                    bulk_g2s(stash + size_t((s & (kSeq - 1)) * kRowsPerBlock + lane) * RS, p.codes + (size_t(m.id[lane ? rowk[1] : rowk[0]]) * p.nCTA + cta) * PWT, CB,
                             &sh.stash_full[s & (kSeq - 1)]);
            }
            PHASE(kPfP0Issue);
            if (end) {                                          // nothing to plan until the sweep finishes or the walk restarts
                while (!ldv(&sh.finished) && ldv(&sh.epoch_w) == epoch) { if (spin_fail()) break; }
                spin_ok();
                continue;
            }
            pos = pos_n; ++next_c;
            idA = idA_n; idB = idB_n;
            // the next block's metadata: its ring entry must be free (see the top of the loop)
            if (s + 1 >= uint32_t(kSeq)) {
                while (ldv(&sh.walk_done_seq) + kSeq <= s + 1 || ldv(&sh.applied_seq) + kSeq <= s + 1) {
                    if (ldv(&sh.finished) || ldv(&sh.epoch_w) != epoch || spin_fail()) break;
                }
                spin_ok();
            }
            if (ldv(&sh.epoch_w) == epoch) gather(s + 1);      // (a restart re-gathers at the top of the loop)
        }
        cp_async_wait_all();
    } else if (warp == kP1) {
        // ---- TMA issue of tiles 2,3; recycling of the global slots ------------------------------------------------------
        uint32_t last_c = 0; bool have_c = false;
        for (uint32_t s = 0;; ++s) {
            while (ldv(&sh.planned_seq) <= s && !ldv(&sh.finished)) { if (spin_fail()) break; }
            spin_ok();
            if (ldv(&sh.planned_seq) <= s || ldv(&sh.abort)) break;
            __threadfence_block();
            const Blk5 &B = sh.blk[s & (kSeq - 1)];
            const Meta5 &m = sh.meta[s & (kSeq - 1)];
            const uint32_t n = B.n, c = B.c;
#pragma unroll
            for (uint32_t t = 2; t < 4; ++t) {
                const uint32_t st = (4 * s + t) & (kStages - 1);
                if (s >= 2) mbar_wait_g(&sh.empty[st], ((s >> 1) - 1) & 1u);
                const uint32_t nt = n > 16 * t ? min(16u, n - 16 * t) : 0u;
                fence_proxy_async();
                if (lane == 0) mbar_arrive_expect_tx(&sh.full[st], nt * CB);
                __syncwarp();
                if (uint32_t(lane) < nt)
                    bulk_g2s(stages + (size_t(st) * kTileRows + lane) * RS, p.codes + (size_t(m.id[16 * t + lane]) * p.nCTA + cta) * PWT, CB, &sh.full[st]);
            }
            if (c != kEnd) {      // block indices (last_c, c] are passed: their slots' next users (256 blocks on) get zeroed by the CTA they map to
                uint32_t i0 = have_c ? last_c + 1 : 0u;
                if (have_c && c < last_c) i0 = c;            // (cannot happen: indices only grow)
                for (uint32_t idx = i0; idx <= c; ++idx) {
                    if (idx % p.nCTA != cta) continue;
                    int *z = p.slots + size_t((idx + kSlotRing / 2) % kSlotRing) * kSlotInts5;
                    for (int ln = lane; ln < kSlotLines; ln += 32) *reinterpret_cast<int4 *>(z + size_t(ln) * kLine) = make_int4(0, 0, 0, 0);
                }
                last_c = c; have_c = true;
            }
        }
    } else if (is_dot) {
        // ---- dot warps ---------------------------------------------------------------------------------------------------
        const uint32_t tile = dw & 3u, kp = dw >> 2;            // this warp: every third word pair (item) of one 16-marker tile
        const uint32_t ti = kp * 32 + lane;                     // thread index inside the tile's 96 threads
        for (uint32_t d = 0;; ++d) {
            PMARK();
            // -- the sequencer: one thread decides what this block is, and which residual updates precede it --
            if (dw == 0 && lane == 0) {
                while (ldv(&sh.planned_seq) <= d) { if (spin_fail()) break; }
                spin_ok();
                __threadfence_block();
                Blk5 &B = sh.blk[d & (kSeq - 1)];
                const bool is_end = B.c == kEnd;
                bool stale = false;
                for (;;) {
                    if (B.epoch != ldv(&sh.epoch_w)) { stale = true; break; }
                    const uint32_t need = is_end ? d : (d >= uint32_t(kD) ? d - kD : 0u);
                    if (ldv(&sh.walk_done_seq) >= need) { __threadfence_block(); stale = B.epoch != ldv(&sh.epoch_w); break; }
                    if (spin_fail()) break;
                }
                spin_ok();
                Decision &D = sh.dec[d & 1];
                uint32_t n_upd = 0;
                auto push_block = [&](uint32_t hs) {            // the updates of a walked block, in marker order
                    const Blk5 &H = sh.blk[hs & (kSeq - 1)];
                    for (uint32_t i = 0; i < H.nrows; ++i) {
                        const double dB = sh.dbeta[hs & (kSeq - 1)][i];
                        if (dB != 0.0) {
                            sh.upd_row[n_upd] = (hs & (kSeq - 1)) * kRowsPerBlock + i; sh.upd_dB[n_upd] = dB;
                            for (int q = 0; q < 4; ++q) sh.upd_lut[n_upd][q] = sh.rowlut[hs & (kSeq - 1)][i][q];
                            ++n_upd;
                        }
                    }
                };
                if (stale) {
                    D.action = kActSkip;
                } else {
                    if (B.epoch != sh.seq_epoch) {              // first block after a restart: everything walked so far goes into the residual
                        const uint32_t rs = ldv(&sh.rs_seq);
                        for (int h = int(sh.hist_n) - 1; h >= 0; --h)
                            if (sh.hist[h] <= rs) push_block(sh.hist[h]);
                        const double dB = sh.ne_dbeta;
                        if (dB != 0.0) {                        // the marker that entered the model: fetch its row
                            uint32_t *dst = stash + size_t(kStash - 1) * RS;
                            fence_proxy_async();
                            mbar_arrive_expect_tx(&sh.ne_full, CB);
                            bulk_g2s(dst, p.codes + (size_t(sh.ne_id) * p.nCTA + cta) * PWT, CB, &sh.ne_full);
                            mbar_wait_g(&sh.ne_full, sh.ne_uses & 1u);
                            ++sh.ne_uses;
                            sh.upd_row[n_upd] = kStash - 1; sh.upd_dB[n_upd] = dB;
                            for (int q = 0; q < 4; ++q) sh.upd_lut[n_upd][q] = sh.ne_lut[q];
                            ++n_upd;
                        }
                        sh.hist_n = 0; sh.seq_epoch = B.epoch;
                    }
                    if (is_end) {
                        for (int h = int(sh.hist_n) - 1; h >= 0; --h) push_block(sh.hist[h]);
                        sh.hist_n = 0;
                        D.action = kActExit;
                    } else {
                        if (sh.hist_n == uint32_t(kD + 1)) { push_block(sh.hist[kD]); sh.hist_n = kD; }
                        uint32_t present = 0;
                        for (uint32_t back = 1; back <= uint32_t(kD); ++back) {
                            const uint32_t hs = back <= sh.hist_n ? sh.hist[back - 1] : kEnd;
                            B.back_seq[back - 1] = hs;
                            if (hs != kEnd) present |= ((1u << sh.blk[hs & (kSeq - 1)].nrows) - 1u) << ((kD - back) * kRowsPerBlock);
                        }
                        present |= ((1u << B.nrows) - 1u) << (kD * kRowsPerBlock);
                        B.present = present;
                        for (int h = int(sh.hist_n); h > 0; --h) sh.hist[h] = sh.hist[h - 1];
                        sh.hist[0] = d; ++sh.hist_n;
                        D.action = kActDot;
                    }
                }
                if (ldv(&sh.abort)) { D.action = kActExit; n_upd = 0; }
                D.n_upd = n_upd;
                // everything up to the oldest block still in the chain has been applied (or is being applied below)
                __threadfence_block();
            }
            bar_named(2, kDotThreads);
            const uint32_t action = sh.dec[d & 1].action, n_upd = sh.dec[d & 1].n_upd;
            PHASE(kPfDGate);
            if (n_upd) {
                // ---- residual update in registers (src/BayesRBase.cpp:375-378 for every update not yet applied, in chain order) ----
#pragma unroll
                for (int it = 0; it < 3; ++it) {
                    const uint32_t item = uint32_t(dt) + it * kDotThreads;
                    if (item < NI) {
                        const uint32_t w = item >> 2, c2 = 2 * (item & 3u);
                        for (uint32_t q = 0; q < n_upd; ++q) {
                            const double dB = sh.upd_dB[q];
                            const uint32_t r = (stash[size_t(sh.upd_row[q]) * RS + w] >> c2) & kM3;
                            const double *L = sh.upd_lut[q];
#pragma unroll
                            for (int t = 0; t < 4; ++t) e[it][t] += dB * L[(r >> (8 * t)) & 3u];
                        }
#pragma unroll
                        for (int t = 0; t < 4; ++t) {      // individuals past n_local carry eps = 0 and code 0: their lut[0] increments must not stick
                            const uint32_t gi = slice0 + 16 * (item >> 2) + 4 * t + (item & 3u);
                            if (gi >= p.n_local) e[it][t] = 0.0;
                        }
                    }
                }
                quantize();
                bar_named(2, kDotThreads);
                if (dt == 0) { sum_wq(); if (PROF && prof_on) sh.prof[kPfApplies] += 1; }
                bar_named(2, kDotThreads);
                PHASE(kPfDApply);
            }
            if (dt == 0) {
                // ring entries older than the oldest block of the chain are dead for the dot warps
                const uint32_t oldest = sh.hist_n ? sh.hist[sh.hist_n - 1] : d + (action == kActExit ? 0u : 1u);
                stv(&sh.applied_seq, oldest);
            }
            if (action == kActExit) break;
            const Blk5 &B = sh.blk[d & (kSeq - 1)];
            const uint32_t st = (4 * d + tile) & (kStages - 1);
            const uint32_t *pk = stages + size_t(st) * kTileRows * RS;
            mbar_wait_g(&sh.full[st], (d >> 1) & 1u);
            if (action == kActSkip) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.empty[st]);
                continue;
            }
            const uint32_t n = B.n, c = B.c, present = B.present;
            const uint32_t nt = n > 16 * tile ? min(16u, n - 16 * tile) : 0u;
            if (B.nrows) mbar_wait_g(&sh.stash_full[d & (kSeq - 1)], (d >> 3) & 1u);
            PHASE(kPfDTma);
            int *slot = p.slots + size_t(c % kSlotRing) * kSlotInts5;
            // Gram row r of this lane (as the B operand's column gid): stashed row of block (kD - r/2) back, its i-th in-model marker
            const uint32_t *urow = stash;
            uint32_t um = 0u;
            {
                const uint32_t r = gid;
                if (r < uint32_t(kGramRows) && ((present >> r) & 1u)) {
                    const uint32_t back = kD - r / kRowsPerBlock, i = r % kRowsPerBlock;
                    const uint32_t hs = back ? B.back_seq[back - 1] : d;
                    urow = stash + size_t((hs & (kSeq - 1)) * kRowsPerBlock + i) * RS;
                    um = 0xffffffffu;
                }
            }
            int acc[4][4], gacc[4][4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                for (int q = 0; q < 4; ++q) { acc[cc][q] = 0; gacc[cc][q] = 0; }
            if (nt) {
                const uint32_t lo = gid * RS + tig;
                uint32_t aR0 = smem_u32(pk + lo + 8 * kp);
                const uint32_t dR1 = 32u * RS;
                uint32_t aL = smem_u32(limbs + lo + 8 * kp);
                uint32_t cnt = NF > kp ? (NF - kp + 2u) / 3u : 0u;
                if (present == 0u) {
                    // ---- dots only: S[m][l] = sum_i g'_mi limb_l(q_i), one MMA per code class (sweep_dense.cuh) ----
                    if (cnt) {
                        uint4 LA = lds_v4<0>(aL), LB = lds_v4<64>(aL);
                        uint32_t a0 = lds_u32<0>(aR0), a1 = lds_u32<0>(aR0 + dR1), b0 = lds_u32<16>(aR0), b1 = lds_u32<16>(aR0 + dR1);
#pragma unroll 1
                        for (;;) {
                            const bool more = --cnt != 0u;
                            if (more) { aR0 += 96u; aL += 384u; }
                            const uint4 LAn = lds_v4<0>(aL), LBn = lds_v4<64>(aL);
                            const uint32_t a0n = lds_u32<0>(aR0), a1n = lds_u32<0>(aR0 + dR1), b0n = lds_u32<16>(aR0), b1n = lds_u32<16>(aR0 + dR1);
                            mma_u8s8(acc[0], a0 & kC0, a1 & kC0, b0 & kC0, b1 & kC0, LA.x, LB.x);
                            mma_u8s8(acc[1], a0 & kC1, a1 & kC1, b0 & kC1, b1 & kC1, LA.y, LB.y);
                            mma_u8s8(acc[2], a0 & kC2, a1 & kC2, b0 & kC2, b1 & kC2, LA.z, LB.z);
                            mma_u8s8(acc[3], a0 & kC3, a1 & kC3, b0 & kC3, b1 & kC3, LA.w, LB.w);
                            if (!more) break;
                            LA = LAn; LB = LBn; a0 = a0n; a1 = a1n; b0 = b0n; b1 = b1n;
                        }
                    }
                } else {
                    // ---- dots + the Gram counts of the stashed rows against this tile's 16 columns (same A fragments) ----
                    uint32_t aU = smem_u32(urow + tig + 8 * kp);
                    if (cnt) {
#pragma unroll 1
                        for (;;) {
                            const uint4 LA = lds_v4<0>(aL), LB = lds_v4<64>(aL);
                            const uint32_t a0 = lds_u32<0>(aR0), a1 = lds_u32<0>(aR0 + dR1), b0 = lds_u32<16>(aR0), b1 = lds_u32<16>(aR0 + dR1);
                            const uint32_t uA = lds_u32<0>(aU) & um, uB = lds_u32<16>(aU) & um;
                            uint32_t x0 = a0 & kC0, x1 = a1 & kC0, x2 = b0 & kC0, x3 = b1 & kC0;
                            mma_u8s8(acc[0], x0, x1, x2, x3, LA.x, LB.x); mma_u8u8(gacc[0], x0, x1, x2, x3, uA & kC0, uB & kC0);
                            x0 = a0 & kC1; x1 = a1 & kC1; x2 = b0 & kC1; x3 = b1 & kC1;
                            mma_u8s8(acc[1], x0, x1, x2, x3, LA.y, LB.y); mma_u8u8(gacc[1], x0, x1, x2, x3, uA & kC1, uB & kC1);
                            x0 = a0 & kC2; x1 = a1 & kC2; x2 = b0 & kC2; x3 = b1 & kC2;
                            mma_u8s8(acc[2], x0, x1, x2, x3, LA.z, LB.z); mma_u8u8(gacc[2], x0, x1, x2, x3, uA & kC2, uB & kC2);
                            x0 = a0 & kC3; x1 = a1 & kC3; x2 = b0 & kC3; x3 = b1 & kC3;
                            mma_u8s8(acc[3], x0, x1, x2, x3, LA.w, LB.w); mma_u8u8(gacc[3], x0, x1, x2, x3, uA & kC3, uB & kC3);
                            if (--cnt == 0u) break;
                            aR0 += 96u; aL += 384u; aU += 96u;
                        }
                    }
                }
                if ((NP & 1u) && NF % 3u == kp) {                                       // the half pair at the end of the slice
                    const uint32_t *r0p = pk + lo + 8 * NF;
                    const uint4 LA = limbs[lo + 8 * NF];
                    const uint32_t a0 = r0p[0], a1 = r0p[8 * RS];
                    const uint32_t uA = urow[tig + 8 * NF] & um;
                    mma_u8s8(acc[0], a0 & kC0, a1 & kC0, 0u, 0u, LA.x, 0u);
                    mma_u8s8(acc[1], a0 & kC1, a1 & kC1, 0u, 0u, LA.y, 0u);
                    mma_u8s8(acc[2], a0 & kC2, a1 & kC2, 0u, 0u, LA.z, 0u);
                    mma_u8s8(acc[3], a0 & kC3, a1 & kC3, 0u, 0u, LA.w, 0u);
                    if (present) {
                        mma_u8u8(gacc[0], a0 & kC0, a1 & kC0, 0u, 0u, uA & kC0, 0u);
                        mma_u8u8(gacc[1], a0 & kC1, a1 & kC1, 0u, 0u, uA & kC1, 0u);
                        mma_u8u8(gacc[2], a0 & kC2, a1 & kC2, 0u, 0u, uA & kC2, 0u);
                        mma_u8u8(gacc[3], a0 & kC3, a1 & kC3, 0u, 0u, uA & kC3, 0u);
                    }
                }
            }
            PHASE(kPfDLoop);
            // ---- sums of q over the missing individuals of the tile's rows (4 lanes per row, warps kp = 0,1: 8 rows each) ----
            if (kp < 2u && nt) {
                const uint32_t r8 = lane >> 2, e4 = lane & 3u, m2 = 8 * kp + r8;
                uint32_t cntm = 0;
                long long qh = 0; unsigned long long ql = 0;
                auto q_of = [&](uint32_t i) -> long long {
                    const uint8_t *lb = reinterpret_cast<const uint8_t *>(limbs) + ((i >> 4) * 4 + (i & 3u)) * 4 + ((i >> 2) & 3u);
                    uint32_t ulo = 0, uhi = 0;
#pragma unroll
                    for (int l = 0; l < 4; ++l) {
                        ulo |= uint32_t(lb[size_t(l) * RS * 16]) << (8 * l);
                        uhi |= uint32_t(lb[size_t(4 + l) * RS * 16]) << (8 * l);
                    }
                    const unsigned long long uq = (unsigned long long)ulo | ((unsigned long long)uhi << 32);
                    return (long long)((uq ^ 0x8080808080808080ULL) - 0x8080808080808080ULL);
                };
                if (m2 < nt) {
                    const uint32_t *row = pk + size_t(m2) * RS;
                    const uint16_t *tail = reinterpret_cast<const uint16_t *>(row + PW);
                    cntm = tail[0];
                    if (cntm <= 15u) {
#pragma unroll 1
                        for (uint32_t en = e4; en < cntm; en += 4) { const long long q = q_of(tail[1 + en]); qh += q >> 32; ql += (unsigned long long)(q & 0xffffffffLL); }
                    } else {                                   // a long list: the four lanes scan the row's words
#pragma unroll 1
                        for (uint32_t w = e4; w < PW; w += 4) {
                            const uint32_t x = row[w];
                            uint32_t mm = x & (x >> 1) & 0x55555555u;
                            while (mm) {
                                const int bit = __ffs(mm) - 1;
                                mm &= mm - 1;
                                const long long q = q_of(16 * w + (bit >> 1)); qh += q >> 32; ql += (unsigned long long)(q & 0xffffffffLL);
                            }
                        }
                    }
                }
#pragma unroll
                for (int o2 = 1; o2 < 4; o2 <<= 1) {
                    qh += __shfl_xor_sync(0xffffffffu, qh, o2);
                    ql += __shfl_xor_sync(0xffffffffu, ql, o2);
                }
                if (e4 == 0 && m2 < nt && cntm) {
                    const uint32_t mg = 16 * tile + m2;
                    if (qh) atomicAdd(reinterpret_cast<unsigned long long *>(slot + size_t(kLSm + 2 * mg) * kLine), (unsigned long long)qh);
                    if (ql) atomicAdd(reinterpret_cast<unsigned long long *>(slot + size_t(kLSm + 2 * mg + 1) * kLine), ql);
                }
            }
            // ---- missing-genotype cross counts of the (Gram row, column) pairs of this tile: one pair per thread ----
            if (present && nt)
            for (uint32_t pidx = ti; pidx < 16u * kGramRows; pidx += 96u) {
                const uint32_t col = pidx & 15u, r = pidx >> 4;
                if (((present >> r) & 1u) && col < nt) {
                    const uint32_t back = kD - r / kRowsPerBlock, i = r % kRowsPerBlock;
                    const uint32_t hs = back ? B.back_seq[back - 1] : d;
                    const bool self = back == 0u && B.rowk[i] == 16 * tile + col;
                    if (!self) {
                        const uint32_t *rowj = stash + size_t((hs & (kSeq - 1)) * kRowsPerBlock + i) * RS;
                        const uint32_t *rowa = pk + size_t(col) * RS;
                        const uint16_t *tlj = reinterpret_cast<const uint16_t *>(rowj + PW), *tla = reinterpret_cast<const uint16_t *>(rowa + PW);
                        const uint32_t cntj = tlj[0], cnta = tla[0];
                        uint32_t z = 0, wv = 0, y = 0;
                        if (cntj <= 15u && cnta <= 15u) {
                            for (uint32_t en = 0; en < cntj; ++en) {
                                const uint32_t ii = tlj[1 + en];
                                const uint32_t cc = (rowa[ii >> 4] >> (2 * (ii & 15u))) & 3u;
                                if (cc == 3u) ++wv; else z += cc;
                            }
                            for (uint32_t en = 0; en < cnta; ++en) {
                                const uint32_t ii = tla[1 + en];
                                const uint32_t cc = (rowj[ii >> 4] >> (2 * (ii & 15u))) & 3u;
                                if (cc != 3u) y += cc;
                            }
                        } else if (cntj | cnta) {
                            for (uint32_t w = 0; w < PW; ++w) {
                                const uint32_t xj = rowj[w], xa = rowa[w];
                                const uint32_t mj = xj & (xj >> 1) & 0x55555555u, ma = xa & (xa >> 1) & 0x55555555u;
                                uint32_t mm = mj | ma;
                                while (mm) {
                                    const int bit = __ffs(mm) - 1;
                                    mm &= mm - 1;
                                    const uint32_t cj2 = (xj >> bit) & 3u, ca2 = (xa >> bit) & 3u;
                                    if (cj2 == 3u) { if (ca2 == 3u) ++wv; else z += ca2; }
                                    else y += cj2;                               // column missing, row present
                                }
                            }
                        }
                        unsigned long long *ln = reinterpret_cast<unsigned long long *>(slot + size_t(kLGram + r * kB + 16 * tile + col) * kLine);
                        if (wv) atomicAdd(ln, (unsigned long long)wv << 32);
                        if (y | z) atomicAdd(ln + 1, (unsigned long long)y | ((unsigned long long)z << 32));
                    }
                }
            }
            // the tile's rows are no longer needed by this warp
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.empty[st]);
            PHASE(kPfDPost);
            // ---- combine the three warps of the tile WITHOUT a barrier: every warp adds its partial sums into the tile's accumulator in
            //      shared memory; the last of the three to arrive adds the totals into the block's slot (and leaves the accumulator zero).
            //      Three accumulator sets in rotation: a warp can be two blocks ahead of its tile mates (stage ring), never three. ----
            {
                const uint32_t ai = (d % 3u) * 4u + tile;
                int *tacc = tile_acc + ai * 256;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int v = acc[0][q] + (acc[1][q] >> 2) + (acc[2][q] >> 4) + (acc[3][q] >> 6);
                    if (v) atomicAdd(&tacc[q * 32 + lane], v);
                    if (present) {
                        const int g = gacc[0][q] + (gacc[1][q] >> 4) + (gacc[2][q] >> 8) + (gacc[3][q] >> 12);
                        if (g) atomicAdd(&tacc[128 + q * 32 + lane], g);
                    }
                }
                __syncwarp();
                unsigned last = 0;
                if (lane == 0) last = atom_acqrel_cta_add(&sh.tilecnt[ai], 1u) == 2u ? 1u : 0u;
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last) {
                    __threadfence_block();
                    {   // (marker m, half h): the limb sums 4h..4h+3 combined into sum_l S_l 2^(8l), exact, as ONE signed 64-bit RED
                        const uint32_t m = lane >> 1, h = lane & 1u;
                        long long v = 0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t l = 4 * h + j, idx = ((l & 1u) + 2 * (m >> 3)) * 32 + (m & 7u) * 4 + (l >> 1);
                            v += (long long)tacc[idx] * (1LL << (8 * j));
                            tacc[idx] = 0;
                        }
                        if (v && m < nt) atomicAdd(reinterpret_cast<unsigned long long *>(slot + size_t(kLDots + 2 * (16 * tile + m) + h) * kLine), (unsigned long long)v);
                    }
                    if (present) {
#pragma unroll
                        for (uint32_t i2 = lane; i2 < 128u; i2 += 32u) {
                            const int a = tacc[128 + i2];
                            if (a) {
                                tacc[128 + i2] = 0;
                                const uint32_t q = i2 >> 5, l2 = i2 & 31u, m = (l2 >> 2) + 8 * (q >> 1), r = 2 * (l2 & 3u) + (q & 1u);
                                if (m < nt && ((present >> r) & 1u))
                                    atomicAdd(reinterpret_cast<unsigned long long *>(slot + size_t(kLGram + r * kB + 16 * tile + m) * kLine), (unsigned long long)(unsigned)a);
                            }
                        }
                    }
                    if (tile == 0 && lane < 2) {
                        const unsigned long long v = sh.sumq[lane];
                        if (v) atomicAdd(reinterpret_cast<unsigned long long *>(slot + size_t(kLEps + lane) * kLine), v);
                    }
                    __syncwarp();
                    // ---- arrive: the last of the four tiles releases the block's counter for the whole grid ----
                    if (lane == 0) {
                        sh.tilecnt[ai] = 0u;
                        if (atom_acqrel_cta_add(&sh.pubcnt[d & (kSeq - 1)], 1u) == 3u) {
                            sh.pubcnt[d & (kSeq - 1)] = 0u;
                            red_release_add(reinterpret_cast<unsigned *>(slot + kLCnt * kLine), 1u);
                        }
                    }
                }
            }
            PHASE(kPfDComb);
        }
    } else {
        // ---- walkers: thread k = marker k of the block ---------------------------------------------------------------------
        const int wk = tid;
        uint32_t my_epoch = 0;
        uint32_t nonskipped = 0, nblocks = 0, nrestarts = 0;
        for (uint32_t s = 0;; ++s) {
            PMARK();
            if (tid == 0) { while (ldv(&sh.planned_seq) <= s) { if (spin_fail()) break; } spin_ok(); }
            bar_named(1, 64);
            __threadfence_block();
            if (ldv(&sh.abort)) break;
            const Blk5 &B = sh.blk[s & (kSeq - 1)];
            if (B.epoch != my_epoch) {                          // planned before the last restart: dropped
                bar_named(1, 64);
                if (tid == 0) stv(&sh.walk_done_seq, s + 1);
                continue;
            }
            const uint32_t c = B.c;
            if (c == kEnd) break;
            const uint32_t n = B.n;
            const Meta5 &m = sh.meta[s & (kSeq - 1)];
            const int *slot = p.slots + size_t(c % kSlotRing) * kSlotInts5;
            PHASE(kPfWPlan);
            if (tid == 0) { while (ld_acquire(reinterpret_cast<const unsigned *>(slot + kLCnt * kLine)) < p.nCTA) { if (spin_fail()) break; } spin_ok(); }
            bar_named(1, 64);
            __threadfence_block();
            PHASE(kPfWBar);
            // (present / back_seq are written by this CTA's sequencer before the block's dot pass, i.e. before this CTA's arrival at the counter)
            const uint32_t present = ldv(&B.present);
            const bool act = uint32_t(wk) < n;
            double bo = 0.0, uu = 2.0, zz = 0.0, base = 0.0; int gg = 0; uint32_t my_id = 0;
            double L[4] = {0.0, 0.0, 0.0, 0.0};
            RowConst rc; rc.sj = rc.cj = rc.ej = rc.sumg = rc.nmiss = 0.0;
            int4 gv[kGramRows];
            if (act) {
                my_id = m.id[wk];
                const double2 l01 = __ldg(reinterpret_cast<const double2 *>(p.lut + 4 * size_t(my_id)));
                const double2 l23 = __ldg(reinterpret_cast<const double2 *>(p.lut + 4 * size_t(my_id) + 2));
                const uint4 cn = __ldg(reinterpret_cast<const uint4 *>(p.counts + 4 * size_t(my_id)));
                const long long tlo = __ldcg(reinterpret_cast<const long long *>(slot + size_t(kLDots + 2 * wk) * kLine));
                const long long thi = __ldcg(reinterpret_cast<const long long *>(slot + size_t(kLDots + 2 * wk + 1) * kLine));
                const long long sm_hi = __ldcg(reinterpret_cast<const long long *>(slot + size_t(kLSm + 2 * wk) * kLine));
                const long long sm_lo = __ldcg(reinterpret_cast<const long long *>(slot + size_t(kLSm + 2 * wk + 1) * kLine));
                const long long e_hi = __ldcg(reinterpret_cast<const long long *>(slot + size_t(kLEps) * kLine));
                const long long e_lo = __ldcg(reinterpret_cast<const long long *>(slot + size_t(kLEps + 1) * kLine));
#pragma unroll
                for (int r = 0; r < kGramRows; ++r)
                    if ((present >> r) & 1u) gv[r] = __ldcg(reinterpret_cast<const int4 *>(slot + size_t(kLGram + r * kB + wk) * kLine));
                L[0] = l01.x; L[1] = l01.y; L[2] = l23.x; L[3] = l23.y;
                const double sx = 0.5 * (L[2] - L[0]);
                const double eps_sum_total = (double(e_hi) * 4294967296.0 + double(e_lo)) * inv_scale;
                const double dprime = (double(thi) * 4294967296.0 + double(tlo)) * inv_scale;
                const double smd = (double(sm_hi) * 4294967296.0 + double(sm_lo)) * inv_scale;
                base = L[0] * eps_sum_total + sx * (dprime - 3.0 * smd) + (L[3] - L[0]) * smd;
                bo = m.beta_old[wk]; uu = m.uz[wk][0]; zz = m.uz[wk][1]; gg = m.grp[wk];
                if (present) {                                  // only blocks with Gram rows need the standardisation constants (a division)
                    const double inv_s = 1.0 / sx, cc = -L[0] * inv_s;
                    rc.sj = sx; rc.cj = cc; rc.ej = L[3] * inv_s + cc;
                    rc.sumg = double(cn.y) + 2.0 * double(cn.z); rc.nmiss = double(cn.w);
                    for (uint32_t i = 0; i < B.nrows; ++i)
                        if (B.rowk[i] == uint32_t(wk)) {
                            sh.rowc[s & (kSeq - 1)][i] = rc;
                            for (int q = 0; q < 4; ++q) sh.rowlut[s & (kSeq - 1)][i][q] = L[q];
                        }
                }
            }
            if (present) {
                bar_named(1, 64);
                // Gram rows: standardise; the rows of earlier blocks complete this block's dots (their updates were not in the residual yet)
                if (act) {
#pragma unroll
                    for (int r = 0; r < kGramRows; ++r) {
                        if (!((present >> r) & 1u)) continue;
                        const uint32_t back = kD - r / kRowsPerBlock, i = r % kRowsPerBlock;
                        const uint32_t hs = back ? B.back_seq[back - 1] : s;
                        const RowConst &rj = sh.rowc[hs & (kSeq - 1)][i];
                        const double g = gram_std(rj, rc, double(unsigned(gv[r].x)), double(unsigned(gv[r].y)), double(unsigned(gv[r].z)), double(unsigned(gv[r].w)), Nt);
                        if (back) { const double dB = sh.dbeta[hs & (kSeq - 1)][i]; if (dB != 0.0) base += dB * g; }
                        else sh.G[i][wk] = g;
                    }
                }
            }
            base += bo * NM1;
            PHASE(kPfWRead);
            Draw dr; dr.beta_new = bo; dr.acum0 = 0.0; dr.comp = -1;
            if (act) dr = sample_marker_k(sh.gc, gg, K, base, p.sigmaE, bo, uu, zz);
            PHASE(kPfWSample);
            // ---- in marker order: the block's in-model markers change the dots of the markers behind them; a marker that enters the
            //      model ends the block right behind itself ----
            uint32_t n_final = n;                               // markers of the block that are final
            bool restart = false;
            {
                uint32_t lo_k = 0;                              // markers below lo_k are final and checked
                for (uint32_t i = 0; i <= B.nrows; ++i) {
                    const uint32_t hi_k = i < B.nrows ? B.rowk[i] : n;      // segment [lo_k, hi_k): no in-model marker inside
                    const bool enters = act && uint32_t(wk) >= lo_k && uint32_t(wk) < hi_k && bo == 0.0 && dr.beta_new != 0.0;
                    const unsigned bal = __ballot_sync(0xffffffffu, enters);
                    uint32_t *fl = sh.w_flag + 2 * ((s + i) & 1u);          // alternating pairs: the next write is two barriers away
                    if (lane == 0) fl[warp] = bal;
                    bar_named(1, 64);
                    const unsigned b0 = fl[0], b1 = fl[1];
                    if (b0 | b1) { n_final = (b0 ? uint32_t(__ffs(b0) - 1) : 32u + uint32_t(__ffs(b1) - 1)) + 1u; restart = true; break; }
                    if (i == B.nrows) break;
                    // marker hi_k (in the model): its draw is final; everything behind it moves by dB x_j.x_k
                    if (uint32_t(wk) == hi_k) sh.bcast[0] = bo - dr.beta_new;
                    bar_named(1, 64);
                    const double dB = sh.bcast[0];
                    if (dB != 0.0 && act && uint32_t(wk) > hi_k) {
                        base += dB * sh.G[i][wk];
                        dr = sample_marker_k(sh.gc, gg, K, base, p.sigmaE, bo, uu, zz);
                    }
                    lo_k = hi_k + 1;
                    if (PROF && prof_on && tid == 0) sh.prof[kPfGramBlocks] += (i == 0);
                }
            }
            PHASE(kPfWFix);
            // ---- file the block: src/BayesRBase.cpp:341, :350-356, the sweep's counts (:355) and sums of squares (:353) ----
            const bool fin = act && uint32_t(wk) < n_final;
            const double dBm = fin ? (bo - dr.beta_new) : 0.0;
            for (uint32_t i = 0; i < B.nrows; ++i)
                if (B.rowk[i] == uint32_t(wk)) sh.dbeta[s & (kSeq - 1)][i] = dBm;      // (0 for a row behind the end of a shortened block)
            const bool nonskip = fin && !(bo == 0.0 && dr.beta_new == 0.0);
            nonskipped += __popc(__ballot_sync(0xffffffffu, nonskip));
            if (c % p.nCTA == cta) {
                double b2 = 0.0; int gk = -1;
                if (fin) {
                    p.acum[my_id] = dr.acum0;
                    if (dr.comp >= 0) {
                        p.beta[my_id] = dr.beta_new;
                        p.comp[my_id] = double(dr.comp);
                        atomicAdd(&sh.v_s[gg * K + dr.comp], 1);
                        if (dr.comp > 0) { b2 = dr.beta_new * dr.beta_new; gk = gg; }
                    }
                }
                unsigned pend = __ballot_sync(0xffffffffu, gk >= 0);
                while (pend) {                                  // one pass per group that gained a non-zero effect (fixed summation order)
                    const int g = __shfl_sync(0xffffffffu, gk, __ffs(pend) - 1);
                    const double s0 = warp_sum(gk == g ? b2 : 0.0);
                    if (lane == 0) sh.betasq_w[warp][g] += s0;
                    pend &= ~__ballot_sync(0xffffffffu, gk == g);
                }
            }
            ++nblocks;
            if (restart) {
                ++nrestarts;
                if (uint32_t(wk) == n_final - 1u) {             // the marker that entered the model
                    sh.ne_id = my_id; sh.ne_dbeta = dBm;
                    for (int q = 0; q < 4; ++q) sh.ne_lut[q] = L[q];
                    sh.rs_pos = B.pos + n_final; sh.rs_c = c + kSkip; sh.rs_seq = s;
                }
                ++my_epoch;
            }
            __threadfence_block();
            bar_named(1, 64);
            if (tid == 0) {
                if (restart) { stv(&sh.epoch_w, my_epoch); __threadfence_block(); }
                stv(&sh.walk_done_seq, s + 1);
            }
            PHASE(kPfWFile);
        }
        if (lane == 0) atomicAdd(&sh.n_nonskipped, nonskipped);           // each warp counted its own lanes
        if (tid == 0) { sh.n_blocks = nblocks; sh.n_restarts = nrestarts; }
    }

    // ---- epilogue: the dot warps have applied every update; eps back to HBM, sweep scalars ------------------------------------
    if (is_dot && dt == 0) stv(&sh.finished, 1u);     // (the walkers left at the END block, the sequencer accepted it: producers may go)
    __syncthreads();
    double sq = 0.0;
    if (is_dot) {
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const uint32_t item = uint32_t(dt) + it * kDotThreads;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const uint32_t gi = slice0 + 16 * (item >> 2) + 4 * t + (item & 3u);
                if (item < NI && gi < p.n_local) { p.eps[gi] = e[it][t]; sq += e[it][t] * e[it][t]; }
            }
        }
    }
    sq = warp_sum(sq);
    if (lane == 0) sh.warp_tmp[warp] = sq;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = kWalkW; w < kP0; ++w) s += sh.warp_tmp[w];
        p.cta_red[2 * cta + 0] = (double((long long)sh.sumq[0]) * 4294967296.0 + double((long long)sh.sumq[1])) * inv_scale;
        p.cta_red[2 * cta + 1] = s;
        if (sh.err) atomicOr(p.out_counts + 2, sh.err);
    }
    for (uint32_t i = tid; i < p.nG * K; i += kThreads) p.part_v[size_t(cta) * p.nG * K + i] = sh.v_s[i];
    for (uint32_t i = tid; i < 2 * p.nG; i += kThreads) p.part_bsq[size_t(cta) * 2 * p.nG + i] = sh.betasq_w[i / p.nG][i % p.nG];
    grid_arrive(p.bar);
    if (cta == 0) {
        grid_wait(p.bar, 2 * p.nCTA);
        if (warp == 0) {
            double s = 0.0, q = 0.0;
            for (uint32_t cc = lane; cc < p.nCTA; cc += 32) { s += ld_cg(p.cta_red + 2 * cc); q += ld_cg(p.cta_red + 2 * cc + 1); }
            s = warp_sum(s); q = warp_sum(q);
            if (lane == 0) {
                p.out_scalars[0] = s; p.out_scalars[1] = q;
                p.out_counts[0] = sh.n_nonskipped; p.out_counts[1] = sh.n_blocks; p.out_counts[3] = sh.n_restarts;
            }
        }
        for (uint32_t i = tid; i < p.nG * K; i += kThreads) {
            int a = 0;
            for (uint32_t cc = 0; cc < p.nCTA; ++cc) a += ld_cg(p.part_v + size_t(cc) * p.nG * K + i);
            p.v[i] = double(a);
        }
        for (uint32_t i = tid; i < p.nG; i += kThreads) {
            double a = 0.0;
            for (uint32_t cc = 0; cc < p.nCTA; ++cc) { a += ld_cg(p.part_bsq + size_t(cc) * 2 * p.nG + i); a += ld_cg(p.part_bsq + size_t(cc) * 2 * p.nG + p.nG + i); }
            p.betasq[i] = a;
        }
        if (PROF && p.prof && tid == 0) for (int i = 0; i < kPfN; ++i) p.prof[i] += sh.prof[i];
    }
}

} // namespace v5
} // namespace brr
