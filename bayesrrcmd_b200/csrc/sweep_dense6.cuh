// Persistent cooperative Gibbs-sweep kernel, dense 2-bit layout (sm_100a), version 6: the steady-state kernel with the walk
// DISTRIBUTED over the CTAs.
//
// One launch = one AnalysisGraph::exec (reference src/sequential.cpp:10-27 driving src/BayesRBase.cpp:287-392 over all
// markers of the sweep, in the given order), synchronous mode, one GPU or several (individuals sharded).  Same arithmetic as
// versions 4 / 5 (exact integer dots of the 2-bit codes against 8 int8 limbs of the fixed-point residual on mma.sync.m16n8k32,
// integer reductions over CTAs and GPUs, fp64 sampling, Gram counts of the in-model rows riding on the dot pass, restart behind a
// marker that enters the model), re-organised around what versions 4 / 5 were measured to be bound by
// (profiles/r2_v5_phase_cycles.txt): every CTA walked every block (read-back of 258 sums + sampling of 64 markers, ~10 k cycles on
// two warps), and the dot warps themselves published, took the missing sums and waited for the walk.
//
//   * The walk of block c is done by ONE CTA, c mod nCTA ("owner").  It reads the block's reduced sums from the block's global
//     slot, samples, and writes a 128-byte RESULT line (markers final, did one enter the model, the residual update of the
//     block's in-model row).  Every CTA consumes the result lines in block order -- one line per block instead of 258 sums
//     and 64 draws.  The head of a line is ONE 16-byte cell, the update with a tag word (block index, markers final, restart
//     flag) in the upper half of both 8-byte words: a reader has everything in one relaxed load (res_store / res_load below).
//     beta / components / acum are written by the owner once its block is known to be part of the chain.
//   * Inside a CTA every stage of a block has its own warps and nothing waits for a later stage (mbarriers between them):
//       warp 11      control: plans the blocks (<= 64 consecutive markers, at most ONE in-model marker each: a second one ends the
//                    block before it), consumes the result lines, decides for every block which residual updates precede its dot
//                    pass (those of the block D + 1 blocks back: a fixed lag, so that all CTAs agree) and which stashed in-model
//                    rows ride on it.  One pass of its loop: the load of the next result line goes out, one block is planned, one
//                    decided, then the line is looked at; its two short lists (chain history, pending updates) are spread over
//                    the lanes;
//       warps 12-15  producers (one per 16-marker tile): TMA bulk copies of the code rows into a ring of 8 stages of 16 rows (one
//                    warp issues one copy per ~100 cycles: profiles/r2_tma_gather_bench.txt); between two issues the TAIL work of
//                    the block before, beside the dot warps' MMA loops: the sums of q over the rows' missing individuals (from the
//                    fp64 residual) and the missing-genotype cross counts of the Gram pairs, two lanes per row;
//       warps 2-9    dot warps: two 16-marker tiles each (one load of the limb fragments feeds both), a quarter of the
//                    individuals each; per-warp partial sums by plain stores into an accumulator set (waited for at the END of
//                    the pass: the publisher has a whole block's time for the set of two blocks back);
//       warp 10      publisher: sums the k-parts of an accumulator set, combines the limbs into 64-bit sums, adds them to the
//                    block's global slot, arrives at the slot's counter (the release fence is paid here, off the dot warps);
//       warps 0-1    walkers: the walk of the blocks this CTA owns.
//   * The residual lives in shared memory (fp64, item order) next to its int8 limbs: updates are applied by the dot warps at
//     the block boundary the control warp names.
//   * Several GPUs: every GPU runs all of the above on its shard; inside the walk the owner pushes the block's locally reduced
//     sums to every peer (LL cells over NVLink) and adds the peers' -- all GPUs walk every block from identical integers.
// A sweep that starts from beta = 0 (many markers entering) should use version 4 (brr_gpu.cu chooses per sweep).
#pragma once
#include "sweep_dense5.cuh"

namespace brr {
namespace v6 {

using v5::ldv; using v5::stv; using v5::bar_named; using v5::mbar_arrive; using v5::gram_std; using v5::RowConst; using v5::spin_check;
using v5::kSlotRing; using v5::kSkip; using v5::kEnd; using v5::kLCnt; using v5::kLDots; using v5::kLSm; using v5::kLEps; using v5::kLGram;
using v5::kSlotLines; using v5::kSlotInts5;

constexpr int kD = 7;                       // LARGEST pipeline depth (array sizes).  The depth of a launch, Sweep6Params::depth = D <= kD: the dot pass of a
                                            // block sees the residual updates of the chain blocks more than D back (4 on one GPU, 7 across GPUs: brr_gpu.cu)
constexpr int kGramRows = kD + 1;           // one in-model row per block: the rows of the last kD blocks and the block's own
static_assert(kGramRows <= 8, "one N-tile of Gram rows");
constexpr int kSeq = 32;                    // ring of block descriptors
constexpr int kAhead = kD + 4;              // blocks planned beyond the last consumed result
constexpr int kIdRing = 16;                 // marker ids of planned blocks (>= kAhead)
constexpr int kStashRing = 32;              // stashed in-model rows, slot = seq % kStashRing.  A row of seq s is last used by the residual update that
                                            // precedes the dot pass of seq s + kD + 1 -- or, behind a restart, of the first block of the new epoch, at most
                                            // seq s + kD + kAhead; its slot is planned again when the result of seq s + kStashRing - kAhead has been consumed
static_assert(kStashRing - kAhead > kD + kAhead, "a stashed row must outlive the update that applies it");
constexpr int kNeSlot = kStashRing;         // + one row for a marker that entered the model
constexpr int kStages = 8, kTileRows = 16;
constexpr int kThreads6 = kThreads;
constexpr int kWalkW = 2, kDotW = 8;
constexpr int kAcc = 2;                     // accumulator sets in rotation
constexpr int kAccDots = kDotW * 256;       // per set: every dot warp's own partial sums [2 tiles][4][32] (plain stores, summed by the publisher),
constexpr int kAccGram = kAccDots;          //          then [4 tiles][128] Gram counts (shared-memory atomics: only blocks that carry Gram rows),
constexpr int kAccSumQ = kAccGram + 4 * 128; //         then sum q as the block's dot pass sees it (2 x u64)
constexpr int kAccInts = kAccSumQ + 4;
constexpr int kMaxUpd = kD + 3;
constexpr uint32_t kResRing = 512;          // result lines (32 ints each)
constexpr int kWPub = kWalkW + kDotW, kPubW = 1, kWCtrl = kWPub + kPubW, kWProd0 = kWCtrl + 1, kProdW = 4;
static_assert(kWProd0 + kProdW == kWarps, "16 warps");
constexpr int kDotThreads = 32 * kDotW;
enum { kActSkip = 0, kActDot = 1, kActExit = 2 };
enum { kPfDDec = 0, kPfDApply, kPfDFull, kPfDLoop, kPfDMiss, kPfDCross, kPfDAcc, kPfPuWait, kPfPuWork, kPfPrPlan, kPfPrEmpty, kPfPrIssue,
       kPfCPoll, kPfCPlan, kPfCDecide, kPfWCnt, kPfWRead, kPfWPred, kPfWSample, kPfWFix, kPfWValid, kPfBlocks, kPfWalks, kPfApplies, kPfRestarts,
       kPfGramBlocks, kPfRowBlocks, kPfDPre, kPfDFullOnly, kPfDLoopGram, kPfCIters, kPfN = 32 };

struct Sweep6Params {
    uint32_t n_local, n_total, M, K, nG, nCTA, SL, PW, RS;
    const uint32_t *codes; const double *lut; const uint32_t *counts; const int32_t *group;
    const uint32_t *order; uint32_t n_order;
    const uint32_t *inmask;      // bit p of word p / 32: beta[order[p]] != 0 at the start of the sweep (inmodel_mask_kernel)
    const double *rand_uz, *pi, *sigmaG, *cva;
    double sigmaE;
    double *eps, *beta, *comp, *acum;
    int *slots;                  // [kSlotRing][kSlotInts5], zero at launch
    int *results;                // [kResRing][32], zero at launch
    GroupConsts *gc;             // [nCTA] scratch: each CTA's table of per-(group, component) constants
    unsigned long long *absmax;  // zero at launch
    unsigned *bar;               // prologue / epilogue grid barrier counter, zero at launch
    double *cta_red;             // [nCTA][2]
    double *part_bsq;            // [nCTA][2][nG]   per-CTA, per-walker-warp sums of beta^2
    int *part_v;                 // [nCTA][nG*K]
    double *v, *betasq, *out_scalars;
    unsigned *out_counts;        // [0] non-skipped, [1] blocks walked, [2] error flags, [3] restarts (zero at launch)
    unsigned long long *prof;    // optional [kPfN + 8 nCTA]: the role counters of CTA 0, then per CTA the first dot warp's wait / work cycles
    // several GPUs (individuals sharded, world > 1): every GPU runs the whole protocol on its shard; the owner CTA of a block pushes
    // the block's locally reduced sums to every peer's IMPORT area over NVLink and adds what the peers pushed before it walks
    int world, rank;
    int *const *peer_xchg;       // [world] base of every rank's exchange buffer (own entry = local buffer)
    unsigned long long imp_off;  // offset (int units) of the version-6 import area inside an exchange buffer
    uint32_t launch_tag;         // number of this launch (identical on all ranks)
    uint32_t depth;              // D, 1..kD
    unsigned *dbg;               // [nCTA][32]: the wait each warp was in when the launch was given up (site | aux << 12 | 0x80000000), zero at launch
};

// Import area of a rank (16-byte cells {lo32 | tag << 32, hi32 | tag << 32}: 8-byte stores arrive whole, a reader that finds the tag in
// both halves has the value -- no flag, no fence, one NVLink trip; as NCCL's LL protocol and the version-4 exchange):
//   header: [world] cells, the peers' max |eps| of this launch;  then [kImpRing blocks][world sources][kImpVals] cells
constexpr int kMaxWorld = 8;                // GPUs the version-6 exchange is written for (more: version 4)
constexpr uint32_t kImpRing = 64;
constexpr uint32_t kImpVals = 4 * kB + 2 + 2 * kGramRows * kB;     // per marker: dot lo / hi, missing hi / lo; sum q hi / lo; Gram words
__host__ __device__ inline size_t imp_u64(int world) { return size_t(2) * world + size_t(2) * kImpRing * world * kImpVals; }
__device__ __forceinline__ void ll_store(unsigned long long *cell, unsigned long long v, uint32_t tag)
{
    st_sys_v2(reinterpret_cast<ulonglong2 *>(cell), (v & 0xffffffffULL) | ((unsigned long long)tag << 32), (v >> 32) | ((unsigned long long)tag << 32));
}
__device__ __forceinline__ bool ll_load(const unsigned long long *cell, uint32_t tag, unsigned long long &v)
{
    const ulonglong2 x = ld_sys_v2(reinterpret_cast<const ulonglong2 *>(cell));
    if (uint32_t(x.x >> 32) != tag || uint32_t(x.y >> 32) != tag) return false;
    v = (x.x & 0xffffffffULL) | (x.y << 32);
    return true;
}

// Sum over all peers of cell `idx` of a block's import entries: every pending peer's cell is loaded in ONE round (a round costs one L2 trip however
// many GPUs there are), the missing ones are polled again.  code: 1 = the block was dropped by a restart, 2 = the launch was given up.
__device__ __noinline__ unsigned long long ll_gather(const unsigned long long *imp, uint32_t idx, uint32_t tag, uint32_t peers, const uint32_t *epoch_w,
                                                     uint32_t b_epoch, uint32_t *abort_flag, unsigned *global_flags, uint32_t &code)
{
    unsigned long long acc = 0;
    uint32_t pending = peers, it = 0;
    long long t0 = 0;
    while (pending) {
        ulonglong2 x[kMaxWorld];
#pragma unroll
        for (int q = 0; q < kMaxWorld; ++q)
            if ((pending >> q) & 1u) x[q] = ld_sys_v2(reinterpret_cast<const ulonglong2 *>(imp + size_t(2) * (size_t(q) * kImpVals + idx)));
#pragma unroll
        for (int q = 0; q < kMaxWorld; ++q)
            if (((pending >> q) & 1u) && uint32_t(x[q].x >> 32) == tag && uint32_t(x[q].y >> 32) == tag) {
                acc += (x[q].x & 0xffffffffULL) | (x[q].y << 32);
                pending &= ~(1u << q);
            }
        if (pending) {
            if (ldv(epoch_w) != b_epoch) { code = 1; break; }
            if (ldv(abort_flag) || ((++it & 255u) == 0u && spin_check(abort_flag, global_flags, t0))) { code = 2; break; }
        }
    }
    return acc;
}

struct Blk6 {
    // planned
    uint32_t c, pos, n, epoch, nrows, rowk, row_id;
    // decided
    uint32_t action, present, n_upd;
    uint32_t back_seq[kD];       // seq of the chain blocks 1..kD back (kEnd = none)
    uint32_t upd_slot[kMaxUpd];
    double upd_dB[kMaxUpd];
};

struct Shared6 {
    Blk6 blk[kSeq];
    uint32_t ids[kIdRing][kB];
    double rowlut[kStashRing + 1][4];
    uint32_t valid_tag[kSeq];
    // control words (volatile accesses)
    uint32_t epoch_w, finished, abort, dots_done, arm_cnt[kStages];
    unsigned long long wq[kWarps][2], sumq[2];
    double scale, inv_scale, warp_tmp[kWarps];
    // walkers
    double bcast[2], w_res;
    uint32_t w_flag[4], w_ctl[4];
    uint32_t n_nonskipped, n_blocks, n_restarts, err;
    unsigned long long prof[kPfN];
    alignas(8) uint64_t full[kStages], empty[kStages], stash_full[kStashRing], ne_full, plan_bar[kSeq], dec_bar[kSeq], acc_full[kAcc], acc_empty[kAcc],
        apply_done;
};

__host__ __device__ inline size_t sweep6_smem_bytes(uint32_t RS, uint32_t SL)
{
    size_t s = (sizeof(Shared6) + 15) & ~size_t(15);
    s += size_t(kLimbs) * RS * 16;                    // limbs
    s += size_t(kStages) * kTileRows * RS * 4;        // code stages
    s += size_t(kStashRing + 1) * RS * 4;             // stashed rows
    s += size_t(kAcc) * kAccInts * 4;                 // accumulator sets
    s += size_t(SL) * 8;                              // residual, fp64
    return s + 16;
}

// bit p of word p / 32 = the marker at position p of the sweep is in the model (src/BayesRBase.cpp:372: its update is never skipped)
__global__ void inmodel_mask_kernel(const uint32_t *__restrict__ order, uint32_t n_order, const double *__restrict__ beta, uint32_t *__restrict__ mask, uint32_t n_words)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = p < n_order && beta[order[p]] != 0.0;
    const unsigned b = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0 && (p >> 5) < n_words) mask[p >> 5] = b;
}

// (timing experiments only, never in a shipped build: parts of the pipeline left out to see what the rest costs -- the results are wrong)
#ifndef BRR_X
#define BRR_X 0
#endif
#ifndef BRR_V6_SUSPEND_NS
#define BRR_V6_SUSPEND_NS 20000
#endif
constexpr uint32_t kSuspendNs = BRR_V6_SUSPEND_NS;
__device__ __forceinline__ bool mbar_test(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    // (with a suspend-time hint: a waiting warp sleeps in the barrier unit instead of spinning through the issue slots of the dot warps)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(kSuspendNs) : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// The head of a block's result line: the update dB of its in-model row with the line's tag word in the upper half of BOTH 8-byte words
// (8-byte accesses are single-copy atomic: a reader that finds the tag in both halves has the whole value, in ONE load).
//   tag word = (c + 1) mod 2^23 | 1 << 23 (a line is never valid while zero) | n_final << 24 | restart << 31
__device__ __forceinline__ uint32_t res_tag(uint32_t c) { return ((c + 1u) & 0x7fffffu) | 0x800000u; }
__device__ __forceinline__ void res_store(int *line, double dB, uint32_t c, uint32_t n_final, bool restart)
{
    const unsigned long long v = (unsigned long long)__double_as_longlong(dB);
    const unsigned long long tw = (unsigned long long)(res_tag(c) | (n_final << 24) | (restart ? 0x80000000u : 0u)) << 32;
    asm volatile("st.release.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(line), "l"((v & 0xffffffffULL) | tw), "l"((v >> 32) | tw) : "memory");
}
// second cell of the line (a restart's "marker that entered the model" update): a plain LL cell, written before the head
__device__ __forceinline__ void res_store_ne(int *line, double dB, uint32_t c)
{
    const unsigned long long v = (unsigned long long)__double_as_longlong(dB);
    const unsigned long long tw = (unsigned long long)res_tag(c) << 32;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(line + 4), "l"((v & 0xffffffffULL) | tw), "l"((v >> 32) | tw) : "memory");
}
__device__ __forceinline__ ulonglong2 res_load(const int *line)
{
    ulonglong2 x;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(x.x), "=l"(x.y) : "l"(line) : "memory");
    return x;
}
// (the load is relaxed so that a poll does not hold back the memory operations behind it.  Everything a reader takes from a line is in its
//  tagged cells; what the release / acquire pair is still needed for is the recycling of the global slots -- a slot is zeroed half a ring,
//  256 blocks, before its next use -- and for that the control warp makes ONE acquire fence every 64 consumed blocks)
#ifdef BRR_V6_POLL_ACQUIRE
__device__ __forceinline__ ulonglong2 res_load_ctrl(const int *line)
{
    ulonglong2 x;
    asm volatile("ld.acquire.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(x.x), "=l"(x.y) : "l"(line) : "memory");
    return x;
}
#else
__device__ __forceinline__ ulonglong2 res_load_ctrl(const int *line) { return res_load(line); }
#endif
#ifndef BRR_V6_IDLE_NS
#define BRR_V6_IDLE_NS 0
#endif
__device__ __forceinline__ void res_acquire() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ bool res_valid(const ulonglong2 &x, uint32_t c)
{
    const uint32_t tw = uint32_t(x.x >> 32);
    return tw == uint32_t(x.y >> 32) && (tw & 0xffffffu) == res_tag(c);
}
__device__ __forceinline__ double res_dB(const ulonglong2 &x) { return __longlong_as_double((long long)((x.x & 0xffffffffULL) | (x.y << 32))); }

template <bool MG, bool PROF>
__global__ void __launch_bounds__(kThreads, 1) sweep_dense6_kernel(const Sweep6Params p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared6 &sh = *reinterpret_cast<Shared6 *>(smem_raw);
    const uint32_t RS = p.RS, PW = p.PW, SL = p.SL;
    const size_t off0 = (sizeof(Shared6) + 15) & ~size_t(15);
    uint4 *limbs = reinterpret_cast<uint4 *>(smem_raw + off0);                                      // [8][RS]
    uint32_t *stages = reinterpret_cast<uint32_t *>(smem_raw + off0 + size_t(kLimbs) * RS * 16);     // [8][16][RS]
    uint32_t *stash = stages + size_t(kStages) * kTileRows * RS;                                    // [kStashRing + 1][RS]
    int *accs = reinterpret_cast<int *>(stash + size_t(kStashRing + 1) * RS);                       // [kAcc][kAccInts]
    double *e_s = reinterpret_cast<double *>(accs + size_t(kAcc) * kAccInts);                       // [SL], item order: e_s[4 * item + t]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t gid = lane >> 2, tig = lane & 3;
    const uint32_t cta = blockIdx.x, K = p.K, D = p.depth;
    const uint32_t NP = SL / 64, NI = SL / 4, NF = NP / 2;
    const uint32_t PWT = PW + kTail, CB = PWT * 4;
    const double NM1 = double(p.n_total - 1), Nt = double(p.n_total);
    const uint32_t slice0 = cta * SL;
    const bool is_dot = warp >= kWalkW && warp < kWPub;
    const int dt = tid - 32 * kWalkW;                 // dot-thread index 0..255
    const uint32_t dw = uint32_t(warp - kWalkW);      // dot-warp index 0..7
    GroupConsts &gc = p.gc[cta];
    const bool prof_on = PROF && p.prof != nullptr && cta == 0 && lane == 0 &&
                         (warp == 0 || warp == kWalkW || warp == kWPub || warp == kWCtrl || warp == kWProd0);   // one warp per role
    long long t_prev = prof_on ? clock64() : 0;
    auto PHASE = [&](int ph) { if (PROF && prof_on) { const long long t = clock64(); sh.prof[ph] += (unsigned long long)(t - t_prev); t_prev = t; } };
    auto PMARK = [&]() { if (PROF && prof_on) t_prev = clock64(); };
    auto PCOUNT = [&](int ph) { if (PROF && prof_on) sh.prof[ph] += 1; };
    // Every wait of the pipeline can be given up: a wait that lasts ~2 s means a protocol error (or a lost CTA).  All other waits of
    // the grid see the flag and fall through, the launch ends, the host reports error flag 4 -- instead of hanging the GPU.
    long long spin_t0 = 0; uint32_t spin_it = 0;
    // (which wait a thread is in -- the index of the mbarrier in Shared6, or 200.. for the other waits -- and at which block: written to
    //  Sweep6Params::dbg when the launch is given up, printed by the host with BRR_GPU_V6_DBG=1)
#ifdef BRR_V6_DEBUG
    uint32_t spin_site = 0, spin_aux = 0;
#define BRR_SITE(site_) (spin_site = (site_))
#define BRR_AUX(aux_) (spin_aux = (aux_))
    auto spin_fail = [&]() -> bool {                  // slow half every 256 polls
        if ((++spin_it & 255u) != 0u) {
            if (ldv(&sh.abort) == 0u) return false;
            if (p.dbg) p.dbg[cta * 32 + warp] = 0x80000000u | spin_site | (spin_aux << 12);
            return true;
        }
        const bool was = ldv(&sh.abort) != 0u || (*reinterpret_cast<volatile unsigned *>(p.out_counts + 2) & 4u);
        const bool r = spin_check(&sh.abort, p.out_counts + 2, spin_t0);
        if (r && p.dbg) p.dbg[cta * 32 + warp] = 0x80000000u | spin_site | (spin_aux << 12);
        if (r && !was) atomicCAS(p.out_counts + 3, 0u, 0x80000000u | spin_site | (uint32_t(warp) << 8) | (cta << 16));
        return r;
    };
#else
#define BRR_SITE(site_) ((void)0)
#define BRR_AUX(aux_) ((void)0)
    auto spin_fail = [&]() -> bool {                  // slow half every 256 polls
        if ((++spin_it & 255u) != 0u) return ldv(&sh.abort) != 0u;
        return spin_check(&sh.abort, p.out_counts + 2, spin_t0);
    };
#endif
    auto spin_ok = [&]() { spin_t0 = 0; spin_it = 0; };
    auto mbw = [&](uint64_t *bar, unsigned parity) -> bool {      // false = the launch was given up
        BRR_SITE((smem_u32(bar) - smem_u32(&sh.full[0])) / 8u);
        for (;;) {
            if (mbar_test(bar, parity)) { spin_ok(); return true; }
            if (spin_fail()) return false;
        }
    };

    // ---- prologue -----------------------------------------------------------------------------------------------------
    double amax = 0.0;
    for (uint32_t idx = tid; idx < SL; idx += kThreads) {          // e_s[4 * item + t] = eps of individual 16 w + 4 t + c, item = 4 w + c
        const uint32_t item = idx >> 2, t = idx & 3u;
        const uint32_t i = 16 * (item >> 2) + 4 * t + (item & 3u);
        const uint32_t gi = slice0 + i;
        const double x = gi < p.n_local ? p.eps[gi] : 0.0;
        e_s[idx] = x;
        amax = fmax(amax, fabs(x));
    }
    fill_group_consts(gc, p.nG, K, NM1, p.sigmaE, p.pi, p.sigmaG, p.cva);
    for (int i = tid; i < kAcc * kAccInts; i += kThreads) accs[i] = 0;
    for (uint32_t i = tid; i < p.nG * K; i += kThreads) p.part_v[size_t(cta) * p.nG * K + i] = 0;
    for (uint32_t i = tid; i < 2 * p.nG; i += kThreads) p.part_bsq[size_t(cta) * 2 * p.nG + i] = 0.0;
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], 5); }      // the 4 dot warps of a tile pair release a stage
        for (int i = 0; i < kStashRing; ++i) mbar_init(&sh.stash_full[i], 1);
        for (int i = 0; i < kSeq; ++i) { mbar_init(&sh.plan_bar[i], 1); mbar_init(&sh.dec_bar[i], 1); sh.valid_tag[i] = 0; }
        for (int i = 0; i < kAcc; ++i) { mbar_init(&sh.acc_full[i], kDotW + kProdW); mbar_init(&sh.acc_empty[i], 1); }
        mbar_init(&sh.ne_full, 1); mbar_init(&sh.apply_done, kDotW);
        mbar_fence_init();
        sh.abort = 0; sh.epoch_w = 0; sh.finished = 0; sh.dots_done = 0;
        for (int i = 0; i < kStages; ++i) sh.arm_cnt[i] = 0;
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
        double a = __longlong_as_double((long long)__ldcg(p.absmax));
        if (MG && tid == 0) {                                           // max |eps| over all shards: pushed by CTA 0, polled by every CTA
            const uint32_t tag = ((p.launch_tag & 0x3ffu) << 21) | 0x1fffffu;
            if (cta == 0)
                for (int q = 0; q < p.world; ++q)
                    if (q != p.rank) ll_store(reinterpret_cast<unsigned long long *>(p.peer_xchg[q] + p.imp_off) + 2 * p.rank, (unsigned long long)__double_as_longlong(a), tag);
            const unsigned long long *hdr = reinterpret_cast<const unsigned long long *>(p.peer_xchg[p.rank] + p.imp_off);
            BRR_SITE(230);
            for (int q = 0; q < p.world; ++q) {
                if (q == p.rank) continue;
                unsigned long long v = 0;
                while (!ll_load(hdr + 2 * q, tag, v)) { if (spin_fail()) break; }
                spin_ok();
                a = fmax(a, __longlong_as_double((long long)v));
            }
            sh.warp_tmp[0] = a;
        }
        if (MG) { __syncthreads(); a = sh.warp_tmp[0]; }
        int ex = 0;
        if (a > 0.0) frexp(a, &ex);
        const int s = 62 - kHeadroomBits - ex;
        if (tid == 0) { sh.scale = ldexp(1.0, s); sh.inv_scale = ldexp(1.0, -s); }
    }
    __syncthreads();
    const double scale = sh.scale, inv_scale = sh.inv_scale;

    // one item (word w, class c: the individuals 16 w + 4 t + c) of the residual -> its 8 limb words; returns q of the four (hi, lo sums)
    auto quantize_item = [&](uint32_t item, const double (&ev)[4], long long &shi, unsigned long long &slo) {
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const double x = ev[t] * scale;
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
    };
    {
        long long shi = 0; unsigned long long slo = 0;
        for (uint32_t item = tid; item < NI; item += kThreads) {
            double ev[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) ev[t] = e_s[4 * item + t];
            quantize_item(item, ev, shi, slo);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { shi += __shfl_xor_sync(0xffffffffu, shi, o); slo += __shfl_xor_sync(0xffffffffu, slo, o); }
        if (lane == 0) { sh.wq[warp][0] = (unsigned long long)shi; sh.wq[warp][1] = slo; }
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long h = 0, l = 0;
        for (int w = 0; w < kWarps; ++w) { h += sh.wq[w][0]; l += sh.wq[w][1]; }
        sh.sumq[0] = h; sh.sumq[1] = l;
    }
    __syncthreads();

    // =====================================================================================================================
    if (warp == kWCtrl) {
        // ---- control warp (warp-uniform state; the two short lists are spread over the lanes: lane q holds entry q) ---------------
        const uint32_t ul = uint32_t(lane);
        uint32_t epoch = 0, pos = 0, next_c = 0;
        uint32_t s_plan = 0, s_cons = 0, s_dec = 0;
        uint32_t pend_n = 0, hist_n = 0, n_restarts = 0;
        bool plan_end = false, restart_pending = false;
        uint32_t ne_user = 0; bool ne_used = false; uint32_t ne_uses = 0;
        uint32_t hreg = kEnd;                                           // lane h: seq of the chain block h back from the last decided one
        uint32_t pslot = 0, pseq = 0; double pdB = 0.0;                 // lane q: pending residual update q (in chain order)
        uint32_t idA = 0, idB = 0, mw0 = 0, mw1 = 0, mw2 = 0;           // prefetched inputs of the block at pos
        const uint32_t n_mw = (p.n_order + 31) / 32;
        auto prefetch = [&]() {
            idA = pos + lane < p.n_order ? __ldg(p.order + pos + lane) : 0u;
            idB = pos + 32 + lane < p.n_order ? __ldg(p.order + pos + 32 + lane) : 0u;
            const uint32_t w = pos >> 5;
            mw0 = w < n_mw ? __ldg(p.inmask + w) : 0u;
            mw1 = w + 1 < n_mw ? __ldg(p.inmask + w + 1) : 0u;
            mw2 = w + 2 < n_mw ? __ldg(p.inmask + w + 2) : 0u;
        };
        prefetch();
        for (;;) {
            if (ldv(&sh.abort)) break;
            bool progress = false;
            PMARK();
            PCOUNT(kPfCIters);
            // ---- (0) the load of the oldest unconsumed block's result line goes out first: its trip to L2 is covered by (1) and (2) ----
            ulonglong2 rx = make_ulonglong2(0ULL, 0ULL);
            bool polled = false;
            if (s_cons < s_plan) {
                const Blk6 &B = sh.blk[s_cons & (kSeq - 1)];
                const uint32_t c = B.c;
                if (B.epoch == epoch && c != kEnd) { rx = res_load_ctrl(p.results + size_t(c % kResRing) * 32); polled = true; }
            }
            // ---- (1) plan the next block -------------------------------------------------------------------------------------
            if (!plan_end && s_plan < s_cons + D + 4u) {
                const uint32_t s = s_plan;
                Blk6 &B = sh.blk[s & (kSeq - 1)];
                const uint32_t rem = p.n_order - pos;
                uint32_t n = min(uint32_t(kB), rem);
                const uint32_t shf = pos & 31u;
                const uint32_t lo = __funnelshift_r(mw0, mw1, shf), hi = __funnelshift_r(mw1, mw2, shf);
                unsigned long long im = (unsigned long long)lo | ((unsigned long long)hi << 32);
                if (n < 64u) im &= (1ULL << n) - 1ULL;
                uint32_t nrows = 0, rowk = 0;
                if (im) {
                    rowk = uint32_t(__ffsll((long long)im) - 1); nrows = 1; im &= im - 1;
                    if (im) n = uint32_t(__ffsll((long long)im) - 1);   // a second in-model marker: the block ends before it
                }
                uint32_t *ids = sh.ids[s & (kIdRing - 1)];
                ids[lane] = idA; ids[32 + lane] = idB;
                const uint32_t row_id = __shfl_sync(0xffffffffu, rowk < 32u ? idA : idB, rowk & 31u);
                const uint32_t slot = s % kStashRing;
                if (nrows && lane < 4) sh.rowlut[slot][lane] = __ldg(p.lut + 4 * size_t(row_id) + lane);
                const bool end = rem == 0;
                if (lane == 0) {
                    B.c = end ? kEnd : next_c; B.pos = pos; B.n = n; B.epoch = epoch; B.nrows = nrows; B.rowk = rowk; B.row_id = row_id;
                    // (the slot's previous phase -- seq s - kStashRing, long complete -- is observed before the barrier is armed again: nobody
                    //  waits on the phase of a block without an in-model row, and a barrier re-armed over an unobserved phase is what
                    //  compute-sanitizer's synccheck calls a missing wait)
                    if (s >= uint32_t(kStashRing)) while (!mbar_test(&sh.stash_full[slot], ((s / kStashRing) - 1u) & 1u)) { }
                    if (nrows) fence_proxy_async();             // (only with a copy to issue: the fence is a MEMBAR, and this thread paces the CTA)
                    mbar_arrive_expect_tx(&sh.stash_full[slot], nrows * CB);
                    if (nrows) bulk_g2s(stash + size_t(slot) * RS, p.codes + (size_t(row_id) * p.nCTA + cta) * PWT, CB, &sh.stash_full[slot]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.plan_bar[s & (kSeq - 1)]);
                ++s_plan; progress = true;
                if (end) plan_end = true;
                else { pos += n; ++next_c; prefetch(); }
            }
            PHASE(kPfCPlan);
            // ---- (2) decide the next block for the dot / post / publisher warps --------------------------------------------
            if (s_dec < s_plan) {
                const uint32_t d = s_dec;
                Blk6 &B = sh.blk[d & (kSeq - 1)];
                const uint32_t b_epoch = B.epoch, b_c = B.c, b_nrows = B.nrows;
                uint32_t action = kActDot, n_take = 0, present = 0;
                bool can = true;
                if (b_epoch != epoch) {
                    action = kActSkip;
                } else if (b_c == kEnd) {
                    if (s_cons == d) { action = kActExit; n_take = pend_n; } else can = false;
                } else {
                    if (restart_pending) { n_take = pend_n; hist_n = 0; }
                    else if (hist_n == D + 1u) {
                        const uint32_t oldest = __shfl_sync(0xffffffffu, hreg, int(D));
                        if (s_cons <= oldest) can = false;
                        else {                                      // the updates of the blocks up to the one that leaves the window
                            const uint32_t due = __ballot_sync(0xffffffffu, ul < pend_n && pseq <= oldest);
                            n_take = uint32_t(__ffs(int(~due))) - 1u;
                            hist_n = D;
                        }
                    }
                }
                if (can) {
                    if (action == kActDot) {
                        restart_pending = false;
                        // lane b = the chain block b back: its seq, and whether it has an in-model row (bit D - b of `present`)
                        const uint32_t up = __shfl_up_sync(0xffffffffu, hreg, 1);
                        const bool is_back = ul >= 1u && ul <= D;
                        const uint32_t hs = (is_back && ul <= hist_n) ? up : kEnd;
                        if (is_back) B.back_seq[ul - 1u] = hs;
                        const bool has_row = hs != kEnd && sh.blk[hs & (kSeq - 1)].nrows != 0u;
                        const uint32_t rows = __ballot_sync(0xffffffffu, has_row);
                        present = (__brev(rows) >> (31u - D)) & ((1u << D) - 1u);
                        if (b_nrows) present |= 1u << D;
                        hreg = ul == 0u ? d : up;                   // (entries beyond hist_n are never read)
                        ++hist_n;
                    }
                    if (ul < n_take) { B.upd_slot[ul] = pslot; B.upd_dB[ul] = pdB; }
                    if (__ballot_sync(0xffffffffu, ul < n_take && pslot == uint32_t(kNeSlot))) { ne_user = d; ne_used = true; }
                    if (n_take) {
                        pslot = __shfl_down_sync(0xffffffffu, pslot, n_take); pseq = __shfl_down_sync(0xffffffffu, pseq, n_take);
                        pdB = __shfl_down_sync(0xffffffffu, pdB, n_take);
                        pend_n -= n_take;
                    }
                    if (lane == 0) { B.action = action; B.present = present; B.n_upd = n_take; }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sh.dec_bar[d & (kSeq - 1)]);
                    ++s_dec; progress = true;
                    if (action == kActExit) {
                        if (lane == 0) { stv(&sh.finished, 1u); sh.n_restarts = n_restarts; }
                        break;
                    }
                }
            }
            PHASE(kPfCDecide);
            // ---- (3) consume the result of the oldest block not yet consumed -------------------------------------------------
            if (s_cons < s_plan) {
                Blk6 &B = sh.blk[s_cons & (kSeq - 1)];
                if (B.epoch != epoch) {                             // planned before a restart: no result to wait for
                    if (ldv(&sh.dots_done) > s_cons) { ++s_cons; progress = true; }
                } else if (polled && res_valid(rx, B.c)) {
                    if ((s_cons & 63u) == 63u) res_acquire();
                    const uint32_t c = B.c;
                    const int *line = p.results + size_t(c % kResRing) * 32;
                    const uint32_t tw = uint32_t(rx.x >> 32);
                    const uint32_t n_final = (tw >> 24) & 0x7fu; const bool restart = (tw >> 31) != 0u;
                    const double dB = res_dB(rx);
                    if (B.nrows && dB != 0.0) {
                        if (ul == pend_n) { pslot = (s_cons % kStashRing) | (((s_cons / kStashRing) & 1u) << 8); pdB = dB; pseq = s_cons; }
                        ++pend_n;
                    }
                    if (lane == 0) stv(&sh.valid_tag[s_cons & (kSeq - 1)], s_cons + 1u);
                    if (restart) {
                        const uint32_t ne_pos = B.pos + n_final - 1u;
                        ulonglong2 nx;
                        while (!res_valid(nx = res_load(line + 4), c)) { if (spin_fail()) break; }
                        spin_ok();
                        const double ne_dB = res_dB(nx);
                        if (ne_dB != 0.0) {                         // the marker that entered the model: fetch its row and value table
                            const uint32_t ne_id = __ldg(p.order + ne_pos);
                            if (ne_used) { while (ldv(&sh.dots_done) <= ne_user) { if (spin_fail()) break; } spin_ok(); }
                            if (lane < 4) sh.rowlut[kNeSlot][lane] = __ldg(p.lut + 4 * size_t(ne_id) + lane);
                            if (lane == 0) {
                                fence_proxy_async();
                                mbar_arrive_expect_tx(&sh.ne_full, CB);
                                bulk_g2s(stash + size_t(kNeSlot) * RS, p.codes + (size_t(ne_id) * p.nCTA + cta) * PWT, CB, &sh.ne_full);
                            }
                            mbw(&sh.ne_full, ne_uses & 1u);
                            ++ne_uses;
                            if (ul == pend_n) { pslot = kNeSlot; pdB = ne_dB; pseq = s_cons; }
                            ++pend_n;
                        }
                        ++epoch; pos = B.pos + n_final; next_c = c + kSkip;
                        restart_pending = true; plan_end = false; ++n_restarts;
                        __syncwarp();
                        if (lane == 0) stv(&sh.epoch_w, epoch);
                        prefetch();
                    }
                    ++s_cons; progress = true;
                }
                __syncwarp();
            }
            PHASE(kPfCPoll);
            BRR_SITE(210 + (s_cons < s_plan ? 1u : 0u) + (s_dec < s_plan ? 2u : 0u) + (plan_end ? 4u : 0u)); BRR_AUX(s_cons | (s_dec << 10));
            if (progress) spin_ok();
            else {
                if (spin_fail()) break;
                if (BRR_V6_IDLE_NS) __nanosleep(BRR_V6_IDLE_NS);     // (an idle pass: leave the issue slots and the result lines to the others)
            }
        }
    } else if (warp >= kWProd0) {
        // ---- producers: producer t takes tile t of every block: stage q = 4 s + t -> ring entry q % 8, i.e. entries t and t + 4 in turn.
        //      The waits on an entry's empty barrier are thus in ONE warp's program order (a parity wait may be at most one phase ahead of
        //      its barrier; with the stages dealt round-robin to three producers this held only because tile pairs are released together and
        //      in block order -- BRR_V6_DEBUG builds check it through sh.arm_cnt) ----
        //      Between two issues the producer does the TAIL work of the block before: the sums of q over the missing individuals of its tile's
        //      rows (q from the fp64 residual: the same rounding as the limbs, one load instead of eight), two lanes per row, and, in blocks that
        //      carry Gram rows, the missing-genotype cross counts of the (stashed row, this row) pairs -- beside the dot warps' MMA loops. ----
        const uint32_t pw = uint32_t(warp - kWProd0);         // tile of this producer
        uint32_t n_apply = 0;
        auto q_of = [&](uint32_t i) -> long long { return __double2ll_rn(e_s[16 * (i >> 4) + 4 * (i & 3u) + ((i >> 2) & 3u)] * scale); };
        bool running = true;
        for (uint32_t s = 0; running; ++s) {
            // ---- issue: tile pw of block s ----
            {
                const uint32_t q = 4 * s + pw, st = q & (kStages - 1);
                PMARK();
                BRR_AUX(s);
                bool ok = true;
                BRR_SITE(220);
                for (;;) {                                          // the block's plan (or the end of the sweep)
                    if (mbar_test(&sh.plan_bar[s & (kSeq - 1)], (s / kSeq) & 1u)) break;
                    if (ldv(&sh.finished) || spin_fail()) { ok = false; break; }
                }
                spin_ok();
                if (!ok) break;
                PHASE(kPfPrPlan);
                const Blk6 &B = sh.blk[s & (kSeq - 1)];
                const uint32_t n = B.n;
                const uint32_t nt = (BRR_X & 8) ? 0u : (n > 16 * pw ? min(16u, n - 16 * pw) : 0u);
                const uint32_t id = sh.ids[s & (kIdRing - 1)][16 * pw + (lane & 15)];
                if (q >= uint32_t(kStages)) {
                    BRR_SITE(221);
                    for (;;) {
                        if (mbar_test(&sh.empty[st], ((q >> 3) - 1u) & 1u)) break;
                        if (ldv(&sh.finished) || spin_fail()) { ok = false; break; }
                    }
                    spin_ok();
                    if (!ok) break;
                }
                PHASE(kPfPrEmpty);
                // (no fence.proxy.async here: the entry's last readers released it through the mbarrier, which orders their generic-proxy
                //  reads before these async-proxy writes)
                if (lane == 0) mbar_arrive_expect_tx(&sh.full[st], nt * CB);
                __syncwarp();
                if (uint32_t(lane) < nt)
                    bulk_g2s(stages + (size_t(st) * kTileRows + lane) * RS, p.codes + (size_t(id) * p.nCTA + cta) * PWT, CB, &sh.full[st]);
                PHASE(kPfPrIssue);
            }
            if (s == 0) continue;
            // ---- tail work of block d = s - 1, tile pw ----
            const uint32_t d = s - 1;
            BRR_AUX(d);
            if (!mbw(&sh.dec_bar[d & (kSeq - 1)], (d / kSeq) & 1u)) break;
            const Blk6 &B = sh.blk[d & (kSeq - 1)];
            const uint32_t action = B.action;
            if (action == kActExit) break;
            if (B.n_upd) { if (!mbw(&sh.apply_done, n_apply & 1u)) break; ++n_apply; }      // the residual this block's dot pass sees
            const uint32_t st = (4 * d + pw) & (kStages - 1);
            if (!mbw(&sh.full[st], (d >> 1) & 1u)) break;
            PHASE(kPfPrEmpty);
            if (action == kActDot && !(BRR_X & 2)) {
                const uint32_t n = B.n, present = B.present;
                const uint32_t *pk = stages + size_t(st) * kTileRows * RS;
                int *slot = p.slots + size_t(B.c % kSlotRing) * kSlotInts5;
                const uint32_t rr = lane >> 1, e2 = lane & 1u, m = 16 * pw + rr;
                const uint32_t *rowa = pk + size_t(rr) * RS;
                const uint16_t *tla = reinterpret_cast<const uint16_t *>(rowa + PW);
                uint32_t cntm = 0;
                long long qh = 0; unsigned long long ql = 0;
                if (m < n) {
                    cntm = tla[0];
                    if (cntm <= 15u) {
#pragma unroll 1
                        for (uint32_t en = e2; en < cntm; en += 2) { const long long q = q_of(tla[1 + en]); qh += q >> 32; ql += (unsigned long long)(q & 0xffffffffLL); }
                    } else {                                   // a long list: the two lanes scan the row's words
#pragma unroll 1
                        for (uint32_t w = e2; w < PW; w += 2) {
                            const uint32_t x = rowa[w];
                            uint32_t mm = x & (x >> 1) & 0x55555555u;
                            while (mm) {
                                const int bit = __ffs(mm) - 1;
                                mm &= mm - 1;
                                const long long q = q_of(16 * w + (bit >> 1)); qh += q >> 32; ql += (unsigned long long)(q & 0xffffffffLL);
                            }
                        }
                    }
                }
                if (__any_sync(0xffffffffu, cntm != 0u)) {
                    qh += __shfl_xor_sync(0xffffffffu, qh, 1);
                    ql += __shfl_xor_sync(0xffffffffu, ql, 1);
                    if (e2 == 0 && m < n && cntm) {
                        if (qh) atomicAdd(reinterpret_cast<unsigned long long *>(slot + size_t(kLSm + 2 * m) * kLine), (unsigned long long)qh);
                        if (ql) atomicAdd(reinterpret_cast<unsigned long long *>(slot + size_t(kLSm + 2 * m + 1) * kLine), ql);
                    }
                }
                PHASE(kPfDMiss);
                if (present) {
                    bool bad = false;
                    for (uint32_t r = 0; r < uint32_t(kGramRows); ++r) {
                        if (!((present >> r) & 1u)) continue;
                        if (pw == 0) PCOUNT(kPfRowBlocks);
                        const uint32_t back = D - r;
                        const uint32_t hs = back ? B.back_seq[back - 1] : d;
                        if (!mbw(&sh.stash_full[hs % kStashRing], (hs / kStashRing) & 1u)) { bad = true; break; }
                        const uint32_t *rowj = stash + size_t(hs % kStashRing) * RS;
                        const uint16_t *tlj = reinterpret_cast<const uint16_t *>(rowj + PW);
                        const uint32_t cntj = tlj[0];
                        const bool pair = m < n && !(back == 0u && B.rowk == m);            // (the row against itself: not needed)
                        uint32_t z = 0, wv = 0, y = 0;
                        if (pair) {
                            if (cntj <= 15u && cntm <= 15u) {
                                for (uint32_t en = e2; en < cntj; en += 2) {
                                    const uint32_t ii = tlj[1 + en];
                                    const uint32_t cc = (rowa[ii >> 4] >> (2 * (ii & 15u))) & 3u;
                                    if (cc == 3u) ++wv; else z += cc;
                                }
                                for (uint32_t en = e2; en < cntm; en += 2) {
                                    const uint32_t ii = tla[1 + en];
                                    const uint32_t cc = (rowj[ii >> 4] >> (2 * (ii & 15u))) & 3u;
                                    if (cc != 3u) y += cc;
                                }
                            } else if (cntj | cntm) {
                                for (uint32_t w = e2; w < PW; w += 2) {
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
                        }
                        unsigned long long pk3 = (unsigned long long)wv | ((unsigned long long)y << 20) | ((unsigned long long)z << 40);
                        pk3 += __shfl_xor_sync(0xffffffffu, pk3, 1);
                        if (e2 == 0 && pair && pk3) {
                            const unsigned long long wv4 = pk3 & 0xfffffULL, y4 = (pk3 >> 20) & 0xfffffULL, z4 = pk3 >> 40;
                            unsigned long long *ln = reinterpret_cast<unsigned long long *>(slot + size_t(kLGram + r * kB + m) * kLine);
                            if (wv4) atomicAdd(ln, wv4 << 32);
                            if (y4 | z4) atomicAdd(ln + 1, y4 | (z4 << 32));
                        }
                    }
                    if (__any_sync(0xffffffffu, bad)) break;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.empty[st]);
            // (this warp arrives at acc_full: the phase of the set's previous block must be over, as for the dot warps)
            if (d >= uint32_t(kAcc) && !mbw(&sh.acc_empty[d % kAcc], ((d / kAcc) - 1u) & 1u)) break;
            if (lane == 0) mbar_arrive(&sh.acc_full[d % kAcc]);
            PHASE(kPfDCross);
        }
    } else if (is_dot) {
        // ---- dot warps: tiles 2 tp, 2 tp + 1 of every block; the 128-individual chunks kp, kp + 4, ... --------------------------------
        const uint32_t tp = dw >> 2, kp = dw & 3u;
        uint32_t n_apply = 0;
        // (profiling builds: where the first dot warp of EVERY CTA spends its cycles -- the CTAs wait for each other through the results)
        const bool cta_prof = PROF && p.prof != nullptr && dw == 0 && lane == 0;
        long long ct[6] = {0, 0, 0, 0, 0, 0}, ct_prev = 0; uint32_t ct_blocks = 0;
        auto CT = [&](int k) { if (PROF && cta_prof) { const long long t = clock64(); ct[k] += t - ct_prev; ct_prev = t; } };
        for (uint32_t d = 0;; ++d) {
            PMARK();
            BRR_AUX(d);
            if (PROF && cta_prof) ct_prev = clock64();
            if (!mbw(&sh.dec_bar[d & (kSeq - 1)], (d / kSeq) & 1u)) break;
            CT(0);
            const Blk6 &B = sh.blk[d & (kSeq - 1)];
            const uint32_t action = B.action, n_upd = B.n_upd;
            PHASE(kPfDDec);
            if (n_upd) {
                // ---- residual update (src/BayesRBase.cpp:375-378 for every update not yet applied, in chain order) ----
                // everything that reads the limbs of the previous block (dot warps, post warp) must be through
                if (d > 0 && !mbw(&sh.acc_full[(d - 1) % kAcc], ((d - 1) / kAcc) & 1u)) break;
                {
                    bool ok = true;
                    for (uint32_t q = 0; q < n_upd; ++q) {
                        const uint32_t us = B.upd_slot[q];
                        if ((us & 0xffu) != uint32_t(kNeSlot)) ok = ok && mbw(&sh.stash_full[us & 0xffu], (us >> 8) & 1u);
                    }
                    if (!ok) break;
                }
                long long shi = 0; unsigned long long slo = 0;
                for (uint32_t item = uint32_t(dt); item < NI; item += kDotThreads) {
                    const uint32_t w = item >> 2, c2 = 2 * (item & 3u);
                    double ev[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) ev[t] = e_s[4 * item + t];
                    long long ohi = 0; unsigned long long olo = 0;
#pragma unroll
                    for (int t = 0; t < 4; ++t) { const long long q = __double2ll_rn(ev[t] * scale); ohi += q >> 32; olo += (unsigned long long)(q & 0xffffffffLL); }
                    for (uint32_t q = 0; q < n_upd; ++q) {
                        const uint32_t slot = B.upd_slot[q] & 0xffu;
                        const double dB = B.upd_dB[q];
                        const uint32_t r = (stash[size_t(slot) * RS + w] >> c2) & kM3;
                        const double *L = sh.rowlut[slot];
#pragma unroll
                        for (int t = 0; t < 4; ++t) ev[t] += dB * L[(r >> (8 * t)) & 3u];
                    }
#pragma unroll
                    for (int t = 0; t < 4; ++t) {      // individuals past n_local carry eps = 0 and code 0: their lut[0] increments must not stick
                        const uint32_t gi = slice0 + 16 * w + 4 * t + (item & 3u);
                        if (gi >= p.n_local) ev[t] = 0.0;
                        e_s[4 * item + t] = ev[t];
                    }
                    quantize_item(item, ev, shi, slo);
                    shi -= ohi; slo -= olo;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { shi += __shfl_xor_sync(0xffffffffu, shi, o); slo += __shfl_xor_sync(0xffffffffu, slo, o); }
                if (lane == 0) { atomicAdd(&sh.sumq[0], (unsigned long long)shi); atomicAdd(&sh.sumq[1], slo); }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.apply_done);
                if (!mbw(&sh.apply_done, n_apply & 1u)) break;
                ++n_apply;
                if (dt == 0) PCOUNT(kPfApplies);
                PHASE(kPfDApply);
            }
            if (action == kActExit) break;
            // sum q as this block's dot pass sees it (no update touches it before every dot warp has arrived at acc_full): kept until the
            // block's accumulator set is free (below)
            unsigned long long sq_snap = 0;
            if (dw == 0 && lane < 2) sq_snap = sh.sumq[lane];
            const uint32_t st0 = (4 * d + 2 * tp) & (kStages - 1);
            const uint32_t *pk = stages + size_t(st0) * kTileRows * RS;      // 32 rows: the two tiles are neighbours in the ring
            PHASE(kPfDPre);
            CT(1);
            if (!mbw(&sh.full[st0], (d >> 1) & 1u)) break;
            if (!mbw(&sh.full[st0 + 1], (d >> 1) & 1u)) break;
            CT(2);
            PHASE(kPfDFullOnly);
            if (action == kActSkip) {
                __syncwarp();
                if (lane == 0) { mbar_arrive(&sh.empty[st0]); mbar_arrive(&sh.empty[st0 + 1]); }
                if (d >= uint32_t(kAcc) && !mbw(&sh.acc_empty[d % kAcc], ((d / kAcc) - 1u) & 1u)) break;
                if (lane == 0) mbar_arrive(&sh.acc_full[d % kAcc]);
                continue;
            }
            const uint32_t n = B.n, present = B.present;
            const bool any = n > 32 * tp;                                   // rows of this tile pair in the block
            // Gram row r of this lane (as the B operand's column gid): the stashed row of the chain block kD - r back
            const uint32_t *urow = stash;
            uint32_t um = 0u;
            if (present) {
                bool bad = false;
                const uint32_t r = gid;
                if ((present >> r) & 1u) {
                    const uint32_t back = D - r;
                    const uint32_t hs = back ? B.back_seq[back - 1] : d;
                    if (!mbw(&sh.stash_full[hs % kStashRing], (hs / kStashRing) & 1u)) bad = true;
                    urow = stash + size_t(hs % kStashRing) * RS;
                    um = 0xffffffffu;
                }
                if (__any_sync(0xffffffffu, bad)) break;
            }
            PHASE(kPfDFull);
            int acc[2][4][4];
#pragma unroll
            for (int tl = 0; tl < 2; ++tl)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[tl][cc][q] = 0;
            int gacc[2][4][4];
            if (any && !(BRR_X & 1)) {
                const uint32_t lo = gid * RS + tig;
                uint32_t aR = smem_u32(pk + lo + 8 * kp);
                const uint32_t dR8 = 32u * RS, dR16 = 64u * RS, dR24 = 96u * RS;       // byte offsets of rows + 8, + 16, + 24
                uint32_t aL = smem_u32(limbs + lo + 8 * kp);
                uint32_t cnt = NF > kp ? (NF - kp + 3u) / 4u : 0u;
                if (present == 0u) {
                    // ---- dots only: S[m][l] = sum_i g'_mi limb_l(q_i), one MMA per code class and tile (sweep_dense.cuh) ----
                    if (cnt) {
                        uint4 LA = lds_v4<0>(aL), LB = lds_v4<64>(aL);
                        uint32_t x0 = lds_u32<0>(aR), x1 = lds_u32<0>(aR + dR8), x2 = lds_u32<16>(aR), x3 = lds_u32<16>(aR + dR8);
                        uint32_t y0 = lds_u32<0>(aR + dR16), y1 = lds_u32<0>(aR + dR24), y2 = lds_u32<16>(aR + dR16), y3 = lds_u32<16>(aR + dR24);
#pragma unroll 1
                        for (;;) {
                            const bool more = --cnt != 0u;
                            if (more) { aR += 128u; aL += 512u; }
                            const uint4 LAn = lds_v4<0>(aL), LBn = lds_v4<64>(aL);
                            const uint32_t x0n = lds_u32<0>(aR), x1n = lds_u32<0>(aR + dR8), x2n = lds_u32<16>(aR), x3n = lds_u32<16>(aR + dR8);
                            const uint32_t y0n = lds_u32<0>(aR + dR16), y1n = lds_u32<0>(aR + dR24), y2n = lds_u32<16>(aR + dR16), y3n = lds_u32<16>(aR + dR24);
                            mma_u8s8(acc[0][0], x0 & kC0, x1 & kC0, x2 & kC0, x3 & kC0, LA.x, LB.x);
                            mma_u8s8(acc[1][0], y0 & kC0, y1 & kC0, y2 & kC0, y3 & kC0, LA.x, LB.x);
                            mma_u8s8(acc[0][1], x0 & kC1, x1 & kC1, x2 & kC1, x3 & kC1, LA.y, LB.y);
                            mma_u8s8(acc[1][1], y0 & kC1, y1 & kC1, y2 & kC1, y3 & kC1, LA.y, LB.y);
                            mma_u8s8(acc[0][2], x0 & kC2, x1 & kC2, x2 & kC2, x3 & kC2, LA.z, LB.z);
                            mma_u8s8(acc[1][2], y0 & kC2, y1 & kC2, y2 & kC2, y3 & kC2, LA.z, LB.z);
                            mma_u8s8(acc[0][3], x0 & kC3, x1 & kC3, x2 & kC3, x3 & kC3, LA.w, LB.w);
                            mma_u8s8(acc[1][3], y0 & kC3, y1 & kC3, y2 & kC3, y3 & kC3, LA.w, LB.w);
                            if (!more) break;
                            LA = LAn; LB = LBn; x0 = x0n; x1 = x1n; x2 = x2n; x3 = x3n; y0 = y0n; y1 = y1n; y2 = y2n; y3 = y3n;
                        }
                    }
                } else {
                    // ---- dots + the Gram counts of the stashed rows against the 32 columns of the pair (same A fragments) ----
#pragma unroll
                    for (int tl = 0; tl < 2; ++tl)
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                            for (int q = 0; q < 4; ++q) gacc[tl][cc][q] = 0;
                    uint32_t aU = smem_u32(urow + tig + 8 * kp);
                    if (cnt) {
#pragma unroll 1
                        for (;;) {
                            const uint4 LA = lds_v4<0>(aL), LB = lds_v4<64>(aL);
                            const uint32_t uA = lds_u32<0>(aU) & um, uB = lds_u32<16>(aU) & um;
#pragma unroll
                            for (int tl = 0; tl < 2; ++tl) {
                                const uint32_t rb = tl ? dR16 : 0u;
                                const uint32_t a0 = lds_u32<0>(aR + rb), a1 = lds_u32<0>(aR + rb + dR8), b0 = lds_u32<16>(aR + rb), b1 = lds_u32<16>(aR + rb + dR8);
                                uint32_t x0 = a0 & kC0, x1 = a1 & kC0, x2 = b0 & kC0, x3 = b1 & kC0;
                                mma_u8s8(acc[tl][0], x0, x1, x2, x3, LA.x, LB.x); mma_u8u8(gacc[tl][0], x0, x1, x2, x3, uA & kC0, uB & kC0);
                                x0 = a0 & kC1; x1 = a1 & kC1; x2 = b0 & kC1; x3 = b1 & kC1;
                                mma_u8s8(acc[tl][1], x0, x1, x2, x3, LA.y, LB.y); mma_u8u8(gacc[tl][1], x0, x1, x2, x3, uA & kC1, uB & kC1);
                                x0 = a0 & kC2; x1 = a1 & kC2; x2 = b0 & kC2; x3 = b1 & kC2;
                                mma_u8s8(acc[tl][2], x0, x1, x2, x3, LA.z, LB.z); mma_u8u8(gacc[tl][2], x0, x1, x2, x3, uA & kC2, uB & kC2);
                                x0 = a0 & kC3; x1 = a1 & kC3; x2 = b0 & kC3; x3 = b1 & kC3;
                                mma_u8s8(acc[tl][3], x0, x1, x2, x3, LA.w, LB.w); mma_u8u8(gacc[tl][3], x0, x1, x2, x3, uA & kC3, uB & kC3);
                            }
                            if (--cnt == 0u) break;
                            aR += 128u; aL += 512u; aU += 128u;
                        }
                    }
                }
                if ((NP & 1u) && (NF & 3u) == kp) {                                     // the half chunk at the end of the slice
                    const uint32_t *r0p = pk + lo + 8 * NF;
                    const uint4 LA = limbs[lo + 8 * NF];
                    const uint32_t uA = urow[tig + 8 * NF] & um;
#pragma unroll
                    for (int tl = 0; tl < 2; ++tl) {
                        const uint32_t a0 = r0p[(16 * tl) * RS], a1 = r0p[(16 * tl + 8) * RS];
                        mma_u8s8(acc[tl][0], a0 & kC0, a1 & kC0, 0u, 0u, LA.x, 0u);
                        mma_u8s8(acc[tl][1], a0 & kC1, a1 & kC1, 0u, 0u, LA.y, 0u);
                        mma_u8s8(acc[tl][2], a0 & kC2, a1 & kC2, 0u, 0u, LA.z, 0u);
                        mma_u8s8(acc[tl][3], a0 & kC3, a1 & kC3, 0u, 0u, LA.w, 0u);
                        if (present) {
                            mma_u8u8(gacc[tl][0], a0 & kC0, a1 & kC0, 0u, 0u, uA & kC0, 0u);
                            mma_u8u8(gacc[tl][1], a0 & kC1, a1 & kC1, 0u, 0u, uA & kC1, 0u);
                            mma_u8u8(gacc[tl][2], a0 & kC2, a1 & kC2, 0u, 0u, uA & kC2, 0u);
                            mma_u8u8(gacc[tl][3], a0 & kC3, a1 & kC3, 0u, 0u, uA & kC3, 0u);
                        }
                    }
                }
            }
            if (present) { PHASE(kPfDLoopGram); } else { PHASE(kPfDLoop); }
            CT(3);
            // the rows are no longer needed by this warp (the tile's producer takes the missing sums / cross counts from them, beside this loop)
            __syncwarp();
            if (lane == 0) { mbar_arrive(&sh.empty[st0]); mbar_arrive(&sh.empty[st0 + 1]); }

            // ---- partial sums into the block's accumulator set: released by the publisher kAcc blocks ago (waited for HERE, not at the top of
            //      the block: the publisher has the whole dot pass of the next block to get through its REDs) ----
            if (d >= uint32_t(kAcc) && !mbw(&sh.acc_empty[d % kAcc], ((d / kAcc) - 1u) & 1u)) break;
            CT(4);
            if (dw == 0 && lane < 2) *(reinterpret_cast<unsigned long long *>(accs + size_t(d % kAcc) * kAccInts + kAccSumQ) + lane) = sq_snap;
            {
                int *aset = accs + size_t(d % kAcc) * kAccInts;
                int *wacc = aset + dw * 256;
#pragma unroll
                for (int tl = 0; tl < 2; ++tl) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int v = any ? acc[tl][0][q] + (acc[tl][1][q] >> 2) + (acc[tl][2][q] >> 4) + (acc[tl][3][q] >> 6) : 0;
                        wacc[(4 * tl + q) * 32 + lane] = v;
                        if (present && any) {
                            const int g = gacc[tl][0][q] + (gacc[tl][1][q] >> 4) + (gacc[tl][2][q] >> 8) + (gacc[tl][3][q] >> 12);
                            if (g) atomicAdd(&aset[kAccGram + (2 * tp + tl) * 128 + q * 32 + lane], g);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.acc_full[d % kAcc]);
            PHASE(kPfDAcc);
            CT(5); ++ct_blocks;
        }
        if (PROF && cta_prof) {
            unsigned long long *o = p.prof + kPfN + size_t(cta) * 8;
            for (int k = 0; k < 6; ++k) o[k] += (unsigned long long)ct[k];
            o[6] += ct_blocks;
            unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            o[7] = smid;
        }
    } else if (warp >= kWPub && warp < kWPub + kPubW) {
        // ---- publishers (two, alternating blocks): accumulator set -> the block's global slot, arrive at its counter; recycling of the global slots ----
        uint32_t last_c = 0; bool have_c = false;
        for (uint32_t d = uint32_t(warp - kWPub);; d += kPubW) {
            PMARK();
            BRR_AUX(d);
            if (d > 0) {                                            // the end of the sweep may be the other publisher's block
                if (!mbw(&sh.dec_bar[(d - 1) & (kSeq - 1)], ((d - 1) / kSeq) & 1u)) break;
                if (sh.blk[(d - 1) & (kSeq - 1)].action == kActExit) break;
            }
            if (!mbw(&sh.dec_bar[d & (kSeq - 1)], (d / kSeq) & 1u)) break;
            const Blk6 &B = sh.blk[d & (kSeq - 1)];
            const uint32_t action = B.action;
            int *aset = accs + size_t(d % kAcc) * kAccInts;
            if (action == kActExit) break;
            if (!mbw(&sh.acc_full[d % kAcc], (d / kAcc) & 1u)) break;
            if (lane == 0) atomicMax(&sh.dots_done, d + 1u);
            PHASE(kPfPuWait);
            if (action == kActDot) {
                const uint32_t n = B.n, c = B.c, present = B.present;
                int *slot = p.slots + size_t(c % kSlotRing) * kSlotInts5;
                {   // (marker m, half h): the limb sums 4h..4h+3, each summed over the four dot warps of the tile pair, combined into
                    // sum_l S_l 2^(8l), exact, as ONE signed 64-bit RED; all loads first
                    int2 pv[4][4][2];
#pragma unroll
                    for (int rep = 0; rep < 4; ++rep) {
                        const uint32_t i = uint32_t(lane) + 32u * rep, m = i >> 1, h = i & 1u, mt = m & 15u, tl = (m >> 4) & 1u, tpm = m >> 5;
                        // warp (tpm, kp): value (tile tl, q, lane gid * 4 + tig) with l = 2 tig + (q & 1), marker gid + 8 (q >> 1)
                        const int *base = aset + (4 * tpm) * 256 + (4 * tl + 2 * (mt >> 3)) * 32 + (mt & 7u) * 4 + 2 * h;
#pragma unroll
                        for (int kp2 = 0; kp2 < 4; ++kp2) {
                            pv[rep][kp2][0] = *reinterpret_cast<const int2 *>(base + kp2 * 256);           // q & 1 = 0: limbs 4h, 4h + 2
                            pv[rep][kp2][1] = *reinterpret_cast<const int2 *>(base + kp2 * 256 + 32);      // q & 1 = 1: limbs 4h + 1, 4h + 3
                        }
                    }
#pragma unroll
                    for (int rep = 0; rep < 4; ++rep) {
                        const uint32_t i = uint32_t(lane) + 32u * rep, m = i >> 1, h = i & 1u;
                        int l0 = 0, l1 = 0, l2 = 0, l3 = 0;
#pragma unroll
                        for (int kp2 = 0; kp2 < 4; ++kp2) { l0 += pv[rep][kp2][0].x; l2 += pv[rep][kp2][0].y; l1 += pv[rep][kp2][1].x; l3 += pv[rep][kp2][1].y; }
                        const long long v = (long long)l0 + (long long)l1 * 256LL + (long long)l2 * 65536LL + (long long)l3 * 16777216LL;
                        if (v && m < n && !(BRR_X & 4)) atomicAdd(reinterpret_cast<unsigned long long *>(slot + size_t(kLDots + 2 * m + h) * kLine), (unsigned long long)v);
                    }
                }
                if (present) {
                    for (uint32_t i2 = lane; i2 < 512u; i2 += 32u) {
                        const uint32_t tile = i2 >> 7, j2 = i2 & 127u;
                        int *tacc = aset + kAccGram + tile * 128;
                        const int a = tacc[j2];
                        if (a) {
                            tacc[j2] = 0;
                            const uint32_t q = j2 >> 5, l2 = j2 & 31u, m = 16 * tile + (l2 >> 2) + 8 * (q >> 1), r = 2 * (l2 & 3u) + (q & 1u);
                            if (m < n && ((present >> r) & 1u))
                                atomicAdd(reinterpret_cast<unsigned long long *>(slot + size_t(kLGram + r * kB + m) * kLine), (unsigned long long)(unsigned)a);
                        }
                    }
                }
                if (lane < 2) {
                    const unsigned long long sq = *(reinterpret_cast<const unsigned long long *>(aset + kAccSumQ) + lane);
                    if (sq) atomicAdd(reinterpret_cast<unsigned long long *>(slot + size_t(kLEps + lane) * kLine), sq);
                }
                __syncwarp();
                if (lane == 0) red_release_add(reinterpret_cast<unsigned *>(slot + kLCnt * kLine), 1u);
                // block indices (last_c, c] are passed: their slots' next users (256 blocks on) get zeroed by the CTA they map to
                // (both publishers do it for the ranges they see: the stores are idempotent)
                uint32_t i0 = have_c ? last_c + 1 : 0u;
                for (uint32_t idx = i0; idx <= c; ++idx) {
                    if (idx % p.nCTA != cta) continue;
                    int *z = p.slots + size_t((idx + kSlotRing / 2) % kSlotRing) * kSlotInts5;
                    for (int ln = lane; ln < kSlotLines; ln += 32) *reinterpret_cast<int4 *>(z + size_t(ln) * kLine) = make_int4(0, 0, 0, 0);
                }
                last_c = c; have_c = true;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.acc_empty[d % kAcc]);
            PHASE(kPfPuWork);
        }
    } else {
        // ---- walkers: the walk of the blocks this CTA owns; thread k = marker k of the block ------------------------------------
        const int wk = tid;
        uint32_t nonskipped = 0, nblocks = 0;
        for (uint32_t s = 0;; ++s) {
            PMARK();
            // the block's plan, or the end of the sweep
            if (tid == 0) {
                uint32_t ctl = 1;
                BRR_SITE(200); BRR_AUX(s);
                for (;;) {
                    if (mbar_test(&sh.plan_bar[s & (kSeq - 1)], (s / kSeq) & 1u)) break;
                    if (ldv(&sh.finished) || spin_fail()) { ctl = 0; break; }
                }
                spin_ok();
                sh.w_ctl[0] = ctl;
            }
            bar_named(1, 64);
            if (!sh.w_ctl[0] || ldv(&sh.abort)) break;
            const Blk6 &B = sh.blk[s & (kSeq - 1)];
            const uint32_t c = B.c, n = B.n, pos = B.pos, b_epoch = B.epoch, nrows = B.nrows, rowk = B.rowk;
            bar_named(1, 64);                                       // (w_ctl is rewritten below)
            if (c == kEnd || c % p.nCTA != cta) continue;
            // ---- per-marker inputs (latency hidden behind the waits below) ----
            const bool act = uint32_t(wk) < n;
            uint32_t my_id = 0;
            double bo = 0.0, uu = 2.0, zz = 0.0; int gg = 0;
            double L[4] = {0.0, 0.0, 0.0, 0.0};
            uint4 cn = make_uint4(0, 0, 0, 0);
            if (act) {
                my_id = __ldg(p.order + pos + wk);
                const double2 l01 = __ldg(reinterpret_cast<const double2 *>(p.lut + 4 * size_t(my_id)));
                const double2 l23 = __ldg(reinterpret_cast<const double2 *>(p.lut + 4 * size_t(my_id) + 2));
                cn = __ldg(reinterpret_cast<const uint4 *>(p.counts + 4 * size_t(my_id)));
                const double2 uz = __ldg(reinterpret_cast<const double2 *>(p.rand_uz + 2 * size_t(my_id)));
                bo = __ldcg(p.beta + my_id);
                gg = p.group ? __ldg(p.group + my_id) : 0;
                L[0] = l01.x; L[1] = l01.y; L[2] = l23.x; L[3] = l23.y; uu = uz.x; zz = uz.y;
            }
            // ---- the decision (Gram rows present, chain blocks behind), then the slot's counter ----
            if (tid == 0) {
                uint32_t ctl = 1;
                if (!mbw(&sh.dec_bar[s & (kSeq - 1)], (s / kSeq) & 1u)) ctl = 0;
                else if (B.action != kActDot) ctl = 2;              // dropped by a restart
                else {
                    const int *slot0 = p.slots + size_t(c % kSlotRing) * kSlotInts5;
                    BRR_SITE(201);
                    while (ld_acquire(reinterpret_cast<const unsigned *>(slot0 + kLCnt * kLine)) < p.nCTA) {
                        if (ldv(&sh.epoch_w) != b_epoch) { ctl = 2; break; }
                        if (spin_fail()) { ctl = 0; break; }
                        __nanosleep(100);
                    }
                    spin_ok();
                }
                sh.w_ctl[1] = ctl;
            }
            bar_named(1, 64);
            {
                const uint32_t ctl = sh.w_ctl[1];
                if (ctl == 0) break;
                if (ctl == 2) continue;
            }
            __threadfence_block();
            PHASE(kPfWCnt);
            const int *slot = p.slots + size_t(c % kSlotRing) * kSlotInts5;
            const uint32_t present = B.present;
            double base = 0.0;
            RowConst rc; rc.sj = rc.cj = rc.ej = rc.sumg = rc.nmiss = 0.0;
            int4 gv[kGramRows];
            long long tlo = 0, thi = 0, sm_hi = 0, sm_lo = 0, e_hi = 0, e_lo = 0;
            if (act || MG) {
                e_hi = __ldcg(reinterpret_cast<const long long *>(slot + size_t(kLEps) * kLine));
                e_lo = __ldcg(reinterpret_cast<const long long *>(slot + size_t(kLEps + 1) * kLine));
            }
            if (act) {
                tlo = __ldcg(reinterpret_cast<const long long *>(slot + size_t(kLDots + 2 * wk) * kLine));
                thi = __ldcg(reinterpret_cast<const long long *>(slot + size_t(kLDots + 2 * wk + 1) * kLine));
                sm_hi = __ldcg(reinterpret_cast<const long long *>(slot + size_t(kLSm + 2 * wk) * kLine));
                sm_lo = __ldcg(reinterpret_cast<const long long *>(slot + size_t(kLSm + 2 * wk + 1) * kLine));
#pragma unroll
                for (int r = 0; r < kGramRows; ++r)
                    if ((present >> r) & 1u) gv[r] = __ldcg(reinterpret_cast<const int4 *>(slot + size_t(kLGram + r * kB + wk) * kLine));
            }
            if (MG) {
                // ---- the other GPUs' sums of this block: push ours to every peer, add theirs (integers: the total does not depend on the order) ----
                const uint32_t tag = ((p.launch_tag & 0x3ffu) << 21) | ((c + 1u) & 0x1fffffu);
                const size_t blk_off = size_t(2) * p.world + size_t(2) * ((size_t(c % kImpRing) * p.world) * kImpVals);
                for (int q = 0; q < p.world; ++q) {
                    if (q == p.rank) continue;
                    unsigned long long *dst = reinterpret_cast<unsigned long long *>(p.peer_xchg[q] + p.imp_off) + blk_off + size_t(2) * p.rank * kImpVals;
                    if (act) {
                        ll_store(dst + 2 * (4 * wk + 0), (unsigned long long)tlo, tag); ll_store(dst + 2 * (4 * wk + 1), (unsigned long long)thi, tag);
                        ll_store(dst + 2 * (4 * wk + 2), (unsigned long long)sm_hi, tag); ll_store(dst + 2 * (4 * wk + 3), (unsigned long long)sm_lo, tag);
#pragma unroll
                        for (int r = 0; r < kGramRows; ++r)
                            if ((present >> r) & 1u) {
                                ll_store(dst + 2 * (4 * kB + 2 + 2 * (r * kB + wk) + 0), (unsigned long long)(unsigned)gv[r].x | ((unsigned long long)(unsigned)gv[r].y << 32), tag);
                                ll_store(dst + 2 * (4 * kB + 2 + 2 * (r * kB + wk) + 1), (unsigned long long)(unsigned)gv[r].z | ((unsigned long long)(unsigned)gv[r].w << 32), tag);
                            }
                    }
                    if (wk < 2) ll_store(dst + 2 * (4 * kB + wk), (unsigned long long)(wk ? e_lo : e_hi), tag);
                }
                uint32_t code = 0;                                  // 1 = dropped by a restart, 2 = given up
                const unsigned long long *imp = reinterpret_cast<const unsigned long long *>(p.peer_xchg[p.rank] + p.imp_off) + blk_off;
                const uint32_t peers = ((1u << p.world) - 1u) & ~(1u << p.rank);
                BRR_SITE(231);
                // cell idx of EVERY peer in one round of loads (ll_gather: out of line, so that its registers are not the kernel's)
                // (two GPUs: the one peer's cells are polled in line -- measured 0.070 s per sweep against 0.079 through the call)
                const unsigned long long *imp1 = imp + size_t(2) * size_t(p.rank ^ 1) * kImpVals;
                auto gather = [&](uint32_t idx) -> unsigned long long {
                    if (code) return 0ULL;
                    if (p.world == 2) {
                        unsigned long long v = 0;
                        while (!code && !ll_load(imp1 + 2 * idx, tag, v)) {
                            if (ldv(&sh.epoch_w) != b_epoch) code = 1;
                            else if (spin_fail()) code = 2;
                        }
                        return v;
                    }
                    return ll_gather(imp, idx, tag, peers, &sh.epoch_w, b_epoch, &sh.abort, p.out_counts + 2, code);
                };
                e_hi += (long long)gather(4 * kB + 0); e_lo += (long long)gather(4 * kB + 1);
                if (act) {
                    tlo += (long long)gather(4 * wk + 0); thi += (long long)gather(4 * wk + 1);
                    sm_hi += (long long)gather(4 * wk + 2); sm_lo += (long long)gather(4 * wk + 3);
#pragma unroll
                    for (int r = 0; r < kGramRows; ++r)
                        if ((present >> r) & 1u) {
                            const unsigned long long w0 = gather(4 * kB + 2 + 2 * (r * kB + wk) + 0), w1 = gather(4 * kB + 2 + 2 * (r * kB + wk) + 1);
                            gv[r].x += int(uint32_t(w0)); gv[r].y += int(uint32_t(w0 >> 32)); gv[r].z += int(uint32_t(w1)); gv[r].w += int(uint32_t(w1 >> 32));
                        }
                }
                spin_ok();
                if (tid == 0) sh.w_ctl[2] = 0;
                bar_named(1, 64);
                if (code) atomicOr(&sh.w_ctl[2], code);
                bar_named(1, 64);
                const uint32_t any = sh.w_ctl[2];
                bar_named(1, 64);
                if (any & 2u) break;
                if (any & 1u) continue;
            }
            if (act) {
                const double sx = 0.5 * (L[2] - L[0]);
                const double eps_sum_total = (double(e_hi) * 4294967296.0 + double(e_lo)) * inv_scale;
                const double dprime = (double(thi) * 4294967296.0 + double(tlo)) * inv_scale;
                const double smd = (double(sm_hi) * 4294967296.0 + double(sm_lo)) * inv_scale;
                base = L[0] * eps_sum_total + sx * (dprime - 3.0 * smd) + (L[3] - L[0]) * smd;
                if (present) {
                    const double inv_s = 1.0 / sx, cc = -L[0] * inv_s;
                    rc.sj = sx; rc.cj = cc; rc.ej = L[3] * inv_s + cc;
                    rc.sumg = double(cn.y) + 2.0 * double(cn.z); rc.nmiss = double(cn.w);
                }
            }
            PHASE(kPfWRead);
            // ---- Gram rows: standardise; the rows of earlier blocks complete this block's dots (their updates were not in the residual yet) ----
            double G_own = 0.0;
            bool dropped = false, given_up = false;
            if (present) {
                for (int r = 0; r < kGramRows; ++r) {
                    if (!((present >> r) & 1u)) continue;
                    const uint32_t back = D - r;
                    const uint32_t hs = back ? B.back_seq[back - 1] : s;
                    const Blk6 &H = sh.blk[hs & (kSeq - 1)];
                    const uint32_t jid = H.row_id;
                    // the row's standardisation constants (every thread derives them: same loads, broadcast by the cache)
                    RowConst rj;
                    {
                        const double2 l01 = __ldg(reinterpret_cast<const double2 *>(p.lut + 4 * size_t(jid)));
                        const double2 l23 = __ldg(reinterpret_cast<const double2 *>(p.lut + 4 * size_t(jid) + 2));
                        const uint4 cj4 = __ldg(reinterpret_cast<const uint4 *>(p.counts + 4 * size_t(jid)));
                        const double sxj = 0.5 * (l23.x - l01.x), inv_s = 1.0 / sxj, cc = -l01.x * inv_s;
                        rj.sj = sxj; rj.cj = cc; rj.ej = l23.y * inv_s + cc;
                        rj.sumg = double(cj4.y) + 2.0 * double(cj4.z); rj.nmiss = double(cj4.w);
                    }
                    double g = 0.0;
                    if (act) g = gram_std(rj, rc, double(unsigned(gv[r].x)), double(unsigned(gv[r].y)), double(unsigned(gv[r].z)), double(unsigned(gv[r].w)), Nt);
                    if (back) {
                        // the update of that row: in the result line of its block (written by that block's owner)
                        const uint32_t hc = H.c;
                        if (tid == 0) {
                            uint32_t ctl = 1;
                            const int *line = p.results + size_t(hc % kResRing) * 32;
                            BRR_SITE(202);
                            ulonglong2 rx;
                            while (!res_valid(rx = res_load(line), hc)) {
                                if (ldv(&sh.epoch_w) != b_epoch) { ctl = 2; break; }
                                if (spin_fail()) { ctl = 0; break; }
                            }
                            spin_ok();
                            if (ctl == 1) { res_acquire(); sh.bcast[1] = res_dB(rx); }
                            sh.w_ctl[2] = ctl;
                        }
                        bar_named(1, 64);
                        const uint32_t ctl = sh.w_ctl[2];
                        const double dBj = sh.bcast[1];
                        bar_named(1, 64);
                        if (ctl == 0) { given_up = true; break; }
                        if (ctl == 2) { dropped = true; break; }
                        if (dBj != 0.0) base += dBj * g;
                    } else G_own = g;
                }
            }
            if (given_up) break;
            if (dropped) continue;
            base += bo * NM1;
            PHASE(kPfWPred);
            Draw dr; dr.beta_new = bo; dr.acum0 = 0.0; dr.comp = -1;
            if (act) dr = sample_marker_k(gc, gg, K, base, p.sigmaE, bo, uu, zz);
            PHASE(kPfWSample);
            // ---- in marker order: the block's in-model marker changes the dots of the markers behind it; a marker that enters the
            //      model ends the block right behind itself ----
            uint32_t n_final = n;
            bool restart = false;
            {
                uint32_t lo_k = 0;
                for (uint32_t i = 0; i <= nrows; ++i) {
                    const uint32_t hi_k = i < nrows ? rowk : n;             // segment [lo_k, hi_k): no in-model marker inside
                    const bool enters = act && uint32_t(wk) >= lo_k && uint32_t(wk) < hi_k && bo == 0.0 && dr.beta_new != 0.0;
                    const unsigned bal = __ballot_sync(0xffffffffu, enters);
                    uint32_t *fl = sh.w_flag + 2 * (i & 1u);
                    if (lane == 0) fl[warp] = bal;
                    bar_named(1, 64);
                    const unsigned b0 = fl[0], b1 = fl[1];
                    if (b0 | b1) { n_final = (b0 ? uint32_t(__ffs(b0) - 1) : 32u + uint32_t(__ffs(b1) - 1)) + 1u; restart = true; break; }
                    if (i == nrows) break;
                    // marker hi_k (in the model): its draw is final; everything behind it moves by dB x_j.x_k
                    if (uint32_t(wk) == hi_k) sh.bcast[0] = bo - dr.beta_new;
                    bar_named(1, 64);
                    const double dB = sh.bcast[0];
                    if (dB != 0.0 && act && uint32_t(wk) > hi_k) {
                        base += dB * G_own;
                        dr = sample_marker_k(gc, gg, K, base, p.sigmaE, bo, uu, zz);
                    }
                    lo_k = hi_k + 1;
                    if (tid == 0) PCOUNT(kPfGramBlocks);
                }
            }
            bar_named(1, 64);                                       // (w_flag / bcast are reused by the next owned block)
            const bool fin = act && uint32_t(wk) < n_final;
            const double dBm = fin ? (bo - dr.beta_new) : 0.0;
            // ---- the result line: what every CTA needs to know about this block ----
            {
                int *line = p.results + size_t(c % kResRing) * 32;
                if (nrows && uint32_t(wk) == rowk) sh.w_res = dBm;
                if (restart && uint32_t(wk) == n_final - 1u) res_store_ne(line, dBm, c);
                bar_named(1, 64);                                   // (the release below is cumulative over the writes this barrier orders before it)
                if (tid == 0) res_store(line, nrows ? sh.w_res : 0.0, c, n_final, restart);
            }
            PHASE(kPfWFix);
            // ---- the block is part of the chain once this CTA's control warp has consumed its result in chain order ----
            if (tid == 0) {
                uint32_t ctl = 1;
                BRR_SITE(203);
                while (ldv(&sh.valid_tag[s & (kSeq - 1)]) != s + 1u) {
                    if (ldv(&sh.epoch_w) != b_epoch) { ctl = ldv(&sh.valid_tag[s & (kSeq - 1)]) == s + 1u ? 1u : 2u; break; }
                    if (spin_fail()) { ctl = 0; break; }
                }
                spin_ok();
                sh.w_ctl[3] = ctl;
            }
            bar_named(1, 64);
            {
                const uint32_t ctl = sh.w_ctl[3];
                if (ctl == 0) break;
                if (ctl == 2) continue;
            }
            PHASE(kPfWValid);
            // ---- file the block: src/BayesRBase.cpp:341, :350-356, the sweep's counts (:355) and sums of squares (:353) ----
            const bool nonskip = fin && !(bo == 0.0 && dr.beta_new == 0.0);
            nonskipped += __popc(__ballot_sync(0xffffffffu, nonskip));
            {
                double b2 = 0.0; int gk = -1;
                if (fin) {
                    p.acum[my_id] = dr.acum0;
                    if (dr.comp >= 0) {
                        p.beta[my_id] = dr.beta_new;
                        p.comp[my_id] = double(dr.comp);
                        atomicAdd(&p.part_v[size_t(cta) * p.nG * K + gg * K + dr.comp], 1);
                        if (dr.comp > 0) { b2 = dr.beta_new * dr.beta_new; gk = gg; }
                    }
                }
                unsigned pend = __ballot_sync(0xffffffffu, gk >= 0);
                while (pend) {                                  // one pass per group that gained a non-zero effect (fixed summation order)
                    const int g = __shfl_sync(0xffffffffu, gk, __ffs(pend) - 1);
                    const double s0 = warp_sum(gk == g ? b2 : 0.0);
                    if (lane == 0) p.part_bsq[size_t(cta) * 2 * p.nG + size_t(warp) * p.nG + g] += s0;
                    pend &= ~__ballot_sync(0xffffffffu, gk == g);
                }
            }
            ++nblocks;
            if (tid == 0) PCOUNT(kPfWalks);
        }
        if (lane == 0) atomicAdd(&sh.n_nonskipped, nonskipped);           // each warp counted its own lanes
        if (tid == 0) sh.n_blocks = nblocks;
    }

    // ---- epilogue: every update has been applied; eps back to HBM, sweep scalars -----------------------------------------
    __syncthreads();
    double sq = 0.0;
    for (uint32_t idx = tid; idx < SL; idx += kThreads) {
        const uint32_t item = idx >> 2, t = idx & 3u;
        const uint32_t gi = slice0 + 16 * (item >> 2) + 4 * t + (item & 3u);
        if (gi < p.n_local) { const double x = e_s[idx]; p.eps[gi] = x; sq += x * x; }
    }
    sq = warp_sum(sq);
    if (lane == 0) sh.warp_tmp[warp] = sq;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += sh.warp_tmp[w];
        p.cta_red[2 * cta + 0] = (double((long long)sh.sumq[0]) * 4294967296.0 + double((long long)sh.sumq[1])) * inv_scale;
        p.cta_red[2 * cta + 1] = s;
        if (sh.err) atomicOr(p.out_counts + 2, sh.err);
        atomicAdd(p.out_counts + 0, sh.n_nonskipped);
        atomicAdd(p.out_counts + 1, sh.n_blocks);
        if (cta == 0 && !(*reinterpret_cast<volatile unsigned *>(p.out_counts + 3) & 0x80000000u)) p.out_counts[3] = sh.n_restarts;
    }
    __threadfence();
    grid_arrive(p.bar);
    if (cta == 0) {
        grid_wait(p.bar, 2 * p.nCTA);
        if (warp == 0) {
            double s = 0.0, q = 0.0;
            for (uint32_t cc = lane; cc < p.nCTA; cc += 32) { s += ld_cg(p.cta_red + 2 * cc); q += ld_cg(p.cta_red + 2 * cc + 1); }
            s = warp_sum(s); q = warp_sum(q);
            if (lane == 0) { p.out_scalars[0] = s; p.out_scalars[1] = q; }
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

} // namespace v6
} // namespace brr
