// Persistent cooperative Gibbs-sweep kernel, dense 2-bit layout (sm_100a), version 4.
//
// One launch = one AnalysisGraph::exec (reference src/sequential.cpp:10-27 driving
// src/BayesRBase.cpp:287-392 over all markers of the sweep, in the given order).
//
// Design
//   * One CTA per SM, all co-resident (cooperative launch).  CTA c owns the individuals
//     [c*SL, (c+1)*SL); their residuals eps live in REGISTERS (fp64, 8 per thread) for the whole sweep,
//     plus an exact fixed-point copy in shared memory: q_i = rint(eps_i * 2^s) split into 8 signed
//     base-256 digits ("limbs", int8).
//   * Genotypes live in HBM as 2-bit codes (16 individuals per 32-bit word).  A block of 64 markers
//     (the next 64 of the permutation) is fetched with TMA 1-D bulk copies (cp.async.bulk + mbarrier),
//     double buffered, two blocks ahead.
//   * All O(N) arithmetic runs on the tensor cores as INTEGER matrix products (mma.sync m16n8k32,
//     u8 x u8 / u8 x s8 -> s32): a register-only 2-bit -> int8 expansion (class c of a code word in place, w & (0x03030303 << 2c),
//     one accumulator per class) feeds
//       - the DOT pass:   64 markers x 8 limbs,   S[m][l] = sum_i g_mi * limb_l(q_i)
//       - the GRAM rows:  u x 64 marker pairs,    GG[j][k] = sum_i g_ji * g_ki, ONLY for the rows j of
//         markers that change the residual (not skipUpdate, src/BayesRBase.cpp:372) -- computed on demand.
//     Both are exact, so the per-CTA partial results are combined over CTAs (and GPUs) with integer
//     reductions (RED.ADD) into a small rotating global slot: the result is bit-identical for every
//     slicing of the individuals and needs no ordered floating-point reduction.
//   * One grid-wide barrier per block, and its latency is hidden: warps 0-1 (the "walkers") wait for the
//     barrier of block b, read the reduced dots back and sample the 64 markers, while warps 2-15 already
//     run the dot pass of block b+1 against the current residual -- speculating that block b leaves the
//     residual untouched (true for most blocks once the chain has left burn-in).  If block b does update
//     the residual, the published sums of block b+1 are completed by a carry taken from the block-(b+1) half of block
//     b's Gram rows: a dot pass is never repeated.
//   * Several GPUs (individuals sharded): the per-block sums of the peers are ADDED into count-carrying accumulators over
//     NVLink and polled at read-back; Gram rows travel as LL cells (see "cross-GPU exchange").
//   * The walk: lane k samples marker k with
//         num_k = x_k.eps_blockstart - sum_{j<k updated} (beta_j,new - beta_j,old) x_j.x_k + beta_old,k (N-1),
//     which is exactly the dot against the residual the one-by-one loop would have seen.  Lanes
//     speculate that no earlier marker of the block changes eps; the markers that do get their Gram rows
//     computed (one more grid-wide round, only then), a triangular solve fixes the betas, and the lanes
//     behind re-sample until every component is confirmed.
//   * Finally every thread applies eps += (beta_old - beta_new) x_j for the block's updated markers to
//     its registers (codes are still resident in shared memory) and re-quantises its limbs.
//   Missing genotypes (code 3) enter the integer products as the value 3 and are corrected exactly by a
//   sparse side pass (sum of q over the missing individuals; cross counts with the other markers).
//   Asynchronous mode (bounded staleness, src/BayesRBase.cpp:394-523 under src/parallelgraph.cpp):
//   same pipeline with B = tokens and no Gram correction.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "sampler.cuh"

namespace brr {

constexpr const char *kDenseKernelVersion = "v4";
constexpr int kB = 64;                 // markers per block = tile structure of the MMA passes
constexpr int kLimbs = 8;
constexpr uint32_t kM3 = 0x03030303u;
// The four 2-bit codes of a byte, left IN PLACE (no shifts): class c of a code word is (w & kCls[c]) = 4^c * code.  An MMA whose
// A (and B) fragments all come from one class yields 4^c (16^c) times the wanted sum; one accumulator per class, recombined by exact
// shifts at the end.  This halves the integer work of expanding the codes, which is what bounds the MMA passes (the integer pipe
// issues one warp instruction per 2 cycles per sub-partition: tools/microbench.cu).
constexpr uint32_t kC0 = 0x03030303u, kC1 = 0x0C0C0C0Cu, kC2 = 0x30303030u, kC3 = 0xC0C0C0C0u;
constexpr int kMissSeg = 256 + 4;      // entries of a warp's missing-list scratch segment (16 words x 16 + sentinels)
constexpr uint32_t kTail = 8;          // tail words of a code row: 16 x uint16 missing list (store_kernels.cuh)
constexpr int kMissCap = kMissSeg * kWarps;
constexpr int kMaxSL = 4096;           // individuals per CTA (2 register items of 4 per thread)
constexpr int kScratchInts = 8192;     // 32 KB shared scratch: dot partials of 16 warps, or the 64x64 Gram rows (doubles)
constexpr int kHeadroomBits = 6;       // |eps| may grow 64x inside a sweep before the fixed point overflows
constexpr int kWalkWarps = 2;          // warps 0,1: barrier wait, read-back, sampling
constexpr int kDotWarps = kWarps - kWalkWarps;   // warps 2..15: dot pass of the next block, TMA issue
static_assert(kDotWarps == 14, "dot_partial / dot_combine split the four marker tiles over 4+4+3+3 dot warps");
constexpr int kMetaWarp0 = kWalkWarps + 10;  // warps 12,13 (dot warps 10,11: 8 dot items, no missing sums) gather and derive the block metadata
constexpr int kTmaLanes = 5;           // rows each dot warp requests per block (14 x 5 >= 64).  One warp issuing all 64 bulk
                                       // copies takes ~3400 cycles (measured, tools/latency_bench2.cu); spread like this ~700.

// rotating global slot (int32 units).  Everything that ALL CTAs add into between two barriers sits on its
// own 128-byte line: RED throughput is per L2 line/slice, and 148 CTAs hammering a few lines was measured to
// cost thousands of cycles per block on the critical path (profiles/).
constexpr int kLine = 32;              // ints per 128-byte line
constexpr int kCols = 2 * kB;          // columns of a Gram row: the 64 markers of the block, then the 64 of the next block
constexpr int kSlotDots = 0;           // [64][2] int64, one per line: the limb sums of a half combined, sum_l S_(4h+l) 2^(8l)
constexpr int kSlotSm = kSlotDots + 128 * kLine;   // [64][2] int64, one per line: sum of q over missing individuals (hi, lo)
constexpr int kSlotEps = kSlotSm + 128 * kLine;    // [2] int64, one per line: sum of q over all individuals (hi, lo)
constexpr int kSlotMainLines = 128 + 128 + 2;
constexpr int kSlotLazy = kSlotEps + 2 * kLine;    // [64 row slots][128 columns], one line each, first 16 bytes:
                                                   //   u64 #0 = a | cw << 32,  u64 #1 = cy | cz << 32   (see gram_entry)
constexpr int kSlotInts = kSlotLazy + kB * kCols * kLine;
// Cross-GPU accumulators carry their own arrival count: every peer adds kCnt1 + its (signed) reduced sum to the word, |total| < 2^55,
// so count = (total + 2^55) >> 56 and value = total - count * 2^56; a reader polls ONE word per value until count == world - 1.
constexpr unsigned long long kCnt1 = 1ULL << 56;
__host__ __device__ inline unsigned line_count(unsigned long long total) { return unsigned((total + (1ULL << 55)) >> 56); }
__host__ __device__ inline long long line_value(unsigned long long total) { return (long long)(total - ((unsigned long long)line_count(total) << 56)); }
constexpr int kSlots = 4;              // blocks b-2 (being recycled), b-1 (still read by slow CTAs), b, b+1 (being published)

struct DenseSweepParams {
    // geometry
    uint32_t n_local, n_total, M, K, nG, B, nCTA, SL, PW, RS;   // PW = SL/16 code words per row (+8 tail words in HBM), RS = smem row stride (words)
    // store
    const uint32_t *codes;   // [M][nCTA][PW + 8]  (store_kernels.cuh)
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
    int *slots;              // [kSlots][kSlotInts], zero at launch
    unsigned long long *absmax;   // bits of max |eps| (zero at launch)
    unsigned *bar;           // grid barrier arrival counter (zero at launch)
    double *cta_red;         // [nCTA][2]
    // outputs
    double *v;               // [nG][K]
    double *betasq;          // [nG]
    double *out_scalars;     // [0]=sum eps, [1]=||eps||^2
    unsigned *out_counts;    // [0]=non-skipped markers, [1]=blocks, [2]=error flags, [3]=grid-wide rounds
    unsigned long long *prof; // optional [kProfPhases]: SM cycles of CTA 0 per phase (nullptr = off)
    // multi-GPU exchange over NVLink peer memory (world > 1): see the "cross-GPU exchange" section of the kernel
    int world, rank;
    int *const *peer_xchg;   // [world] device pointers to every rank's exchange buffer (own entry = local buffer)
    unsigned *bar2;          // local counter: exchange CTAs that have merged the peers' contributions (zero at launch)
    unsigned long long *round_ctr;   // exchange rounds completed by earlier launches (persistent; identical on all ranks)
    uint32_t xchg_ctas;      // CTAs sharing the exchange work of a Gram round: min(kXchgCtas, smallest grid of any rank)
};

// exchange buffer of a rank (uint64 units): mailbox[src][rot] of kMsgU64, then flags[src][kXchgCtas] (round ids)
constexpr int kXchgCtas = 8;           // CTAs that push / merge a share of a Gram round's counts
constexpr int kXchgRot = 4;
constexpr int kMsgLazy = 0;                        // [64 row slots][128 columns][2]
constexpr int kMsgDots = kMsgLazy + 2 * kB * kCols; // [64][2]
constexpr int kMsgSm = kMsgDots + 2 * kB;          // [64][2]
constexpr int kMsgEps = kMsgSm + 2 * kB;           // [2]
constexpr int kMsgAbs = kMsgEps + 2;               // [1]
constexpr int kMsgU64 = kMsgAbs + 2;
__host__ __device__ inline size_t xchg_mailbox_ints(int world) { return size_t(world) * kXchgRot * kMsgU64 * 2; }
// The per-block message (the 258 reduced 64-bit sums of a block: 128 dot halves, 128 missing sums, 2 residual sums, in slot-line
// order) travels in "LL" form, as NCCL's low-latency protocol does: each value is one 16-byte cell {lo32 | id << 32, hi32 | id << 32},
// id = the low 32 bits of the round id, written with ONE vector store.  8-byte stores arrive whole, so a reader that finds the
// round's id in both halves has the value: no flag, no release fence, no second trip over NVLink.
constexpr int kMainVals = 4 * kB + 2;
__host__ __device__ inline size_t xchg_ll_off_ints(int world) { return xchg_mailbox_ints(world) + size_t(world) * kXchgCtas * 2; }
__host__ __device__ inline size_t xchg_ll_ints(int world) { return size_t(world) * kXchgRot * kMainVals * 4; }
// The Gram rows of a lazy round travel the same way when every CTA consumes them at read-back (the usual case: the block's rows
// fit the shared-memory row slots): value i of the round = u64 #(i & 1) of entry (row - r0, column) = i >> 1.
constexpr int kLazyCells = 2 * kB * kCols;
// Alternative for the Gram-row rounds, NOT enabled yet (correct on one device, to be measured on 8 GPUs next round): count-carrying
// accumulators as for the per-block round (one word per value at the head of the lazy LL region, peers add value + kCnt1, the push
// spread over all CTAs, readers poll two words per entry whatever the number of GPUs).  The counts a, cw, cy, cz are non-negative
// and below 2^24 in the upper halves, so count = word >> 56 and value = word & (2^56 - 1).
constexpr bool kLazyAcc = false;
__host__ __device__ inline size_t xchg_lazy_ll_off_ints(int world) { return xchg_ll_off_ints(world) + xchg_ll_ints(world); }
__host__ __device__ inline size_t xchg_total_ints(int world) { return xchg_lazy_ll_off_ints(world) + size_t(world) * kXchgRot * kLazyCells * 4 + 8; }
enum { kXAbs = 0, kXMain = 1, kXLazy = 2 };

// phases of the block loop, as timed by thread 0 of CTA 0 (a walker) when DenseSweepParams::prof is set;
// kProfWorker* are timed by the first lane of warp 2 (a dot-pass warp) and overlap the walker's phases
enum { kProfWait = 0, kProfXchg, kProfReadback, kProfSample, kProfJoin, kProfPublish, kProfLazyCross, kProfLazyMma, kProfLazyBarrier,
       kProfLazyReadback, kProfSolve, kProfResample, kProfFinalize, kProfCarry, kProfAxpy, kProfQuantize, kProfTail, kProfPrologue,
       kProfWalkIters, kProfLazyRounds, kProfLazyRows, kProfBlocksUpdated, kProfWorkerWait, kProfWorkerDot,
       // finer marks inside publish and the Gram-row products (already contained in kProfPublish / kProfLazyMma)
       kProfPubReduce, kProfPubMissing, kProfPubArrive, kProfLzLoop, kProfLzRed, kProfPubPre, kProfLzC0, kProfLzC1, kProfLzC2,
       kProfWDot, kProfWMissing, kProfWCombine, kProfPhases = 40 };

struct alignas(16) BlockMeta {
    uint32_t id[kB];
    int32_t grp[kB];
    double beta_old[kB];
    double uz[kB][2];                // (u, z) of src/BayesRBase.cpp:343-344
    double lut[kB][4];
    uint32_t cnt[kB][4];             // global class counts
    // derived at read-back time
    double sj[kB], cj[kB], ej[kB];   // x = sj * (g - cj + ej*[missing])
    double sumg[kB], nmiss[kB];      // global n1 + 2 n2, global missing count
};

struct BlockResult {                 // what the walk leaves for one block; CTA 0 flushes it to HBM off the critical path
    uint32_t id[kB];
    int32_t grp[kB];
    int comp[kB];
    double beta[kB], acum[kB];
    uint32_t n;
};

struct SweepShared {   // fixed-size part; limbs, code buffers, scratch and the missing list follow (dynamic)
    GroupConsts gc;
    BlockMeta meta[2];               // per code buffer
    BlockResult fin[2];              // by block parity: written by the walk of block b, flushed two passes later
    double d0[kB];
    double dbeta[kB];
    double carry[kB];                // correction of the NEXT block's dots by this block's residual updates
    double w_corr[kB];               // walk: corrections of the second half-block
    int w_comp[kB];                  // walk: tentative components
    uint32_t w_nsm[2], w_chg[2], w_cnt[2], w_cnt2[2];
    uint32_t lz_rows[kB];            // markers (block-local) of the Gram rows being computed in this round
    uint32_t rowslot[kB];            // Gram row slot of a marker (valid for the rows computed so far)
    uint32_t slotrow[kB];            // marker of a Gram row slot
    uint32_t lz_hist[kSlots];        // row slots used by the block that last used each global slot
    double warp_tmp[kWarps];
    double betasq_s[kMaxGK];
    int v_s[kMaxGK];
    unsigned long long wq[kWarps][2]; // per-warp sums of q over the slice, (hi, lo) parts
    double scale, inv_scale;
    uint32_t ns_list[kB];
    uint32_t lza_hist[4];            // kLazyAcc: Gram rows sent through each rotation of the lazy accumulators by its last round
    uint32_t n_ns;
    uint32_t n_nonskipped_total;
    uint32_t err;
    unsigned long long prof[kProfPhases];   // phase cycle counters of this CTA (flushed by CTA 0 at the end)
    alignas(8) uint64_t mbar[2];
};

__host__ __device__ inline size_t sweep_smem_bytes(uint32_t RS)
{
    size_t s = (sizeof(SweepShared) + 15) & ~size_t(15);
    s += size_t(kLimbs) * RS * 16;              // limbs: [8][RS] uint4
    s += size_t(2) * kB * RS * 4;               // two code buffers: [64][RS] words
    s += size_t(kScratchInts) * 4;              // scratch
    s += size_t(kMissCap) * 4;                  // missing list
    return s + 16;
}

// ---- tensor-core primitives ---------------------------------------------------------------------
__device__ __forceinline__ void mma_u8u8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_u8s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// shared-memory loads by 32-bit address with an immediate byte offset: the MMA loops keep three address registers and bump them,
// nothing else (the compiler's own addressing of these loops re-derived the addresses from tid every iteration: ~45 of ~85
// instructions, and the passes are bound by instruction issue)
template <int OFF> __device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF> __device__ __forceinline__ uint4 lds_v4(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_sys_add_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_v2(ulonglong2 *p, unsigned long long a, unsigned long long b)
{
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_sys_v2(const ulonglong2 *p)
{
    ulonglong2 v;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __noinline__ void prof_mark(unsigned long long *slot_cycles, long long *t_prev, bool charge)
{
    const long long t = clock64();
    if (charge) *slot_cycles += (unsigned long long)(t - *t_prev);
    *t_prev = clock64();
}
__device__ __forceinline__ void walker_sync() { asm volatile("bar.sync 2, 64;" ::: "memory"); }   // warps 0,1

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// per-marker inputs of a block: asynchronous gather into shared memory (thread k = marker k of the block) ...
__device__ __forceinline__ void gather_block_meta(const DenseSweepParams &p, BlockMeta &m, uint32_t id, int k)
{
    m.id[k] = id;
    cp_async16(&m.lut[k][0], p.lut + 4 * size_t(id));
    cp_async16(&m.lut[k][2], p.lut + 4 * size_t(id) + 2);
    cp_async16(&m.cnt[k][0], p.counts + 4 * size_t(id));
    cp_async16(&m.uz[k][0], p.rand_uz + 2 * size_t(id));
    cp_async8(&m.beta_old[k], p.beta + id);
    if (p.group) cp_async4(&m.grp[k], p.group + id); else m.grp[k] = 0;
}
// ... and, once it has landed, what the walk derives from it
__device__ __forceinline__ void derive_block_meta(BlockMeta &m, int k)
{
    const double l0 = m.lut[k][0], l2 = m.lut[k][2], l3 = m.lut[k][3];
    const double s = 0.5 * (l2 - l0);          // value table is affine in the allele count
    const double inv_s = 1.0 / s;
    const double cc = -l0 * inv_s;
    m.sj[k] = s; m.cj[k] = cc; m.ej[k] = l3 * inv_s + cc;
    m.sumg[k] = double(m.cnt[k][1]) + 2.0 * double(m.cnt[k][2]);
    m.nmiss[k] = double(m.cnt[k][3]);
}

// x_j . x_k from the reduced integer counts of the pair (j = the Gram row, a marker of block `mj`; k = the column, a
// marker of block `mk` -- the same block or the next one):
//   a  = sum_i g'_j g'_k with missing genotypes entering as the value 3
//   cw = #{i: both missing},  cy = sum over {i: k missing, j present} of g_j,  cz = sum over {i: j missing, k present} of g_k
__device__ __forceinline__ double gram_entry(const BlockMeta &mj, int j, const BlockMeta &mk, int k, double a, double cw, double cy,
                                             double cz, double Nt)
{
    a = a - 3.0 * cy - 3.0 * cz - 9.0 * cw;                  // sum over both present of g_j g_k
    const double cj = mj.cj[j], ck = mk.cj[k], ej = mj.ej[j], ek = mk.ej[k];
    double t = a - ck * mj.sumg[j] - cj * mk.sumg[k] + Nt * cj * ck;
    t += ek * cy + ej * cz - cj * ek * mk.nmiss[k] - ck * ej * mj.nmiss[j] + ej * ek * cw;
    return mj.sj[j] * mk.sj[k] * t;
}

// MG: several GPUs (world > 1) -- the single-GPU instantiation carries no exchange code at all; PROF: phase cycle counters
// (BRR_GPU_PROFILE) -- the production instantiations carry no timing calls.  Smaller code and fewer live registers on the
// walkers' critical path are worth the four instantiations.
template <bool MG, bool PROF>
__global__ void __launch_bounds__(kThreads, 1) sweep_dense_kernel(const DenseSweepParams p)
{
    const int world = MG ? p.world : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SweepShared &sh = *reinterpret_cast<SweepShared *>(smem_raw);
    const uint32_t RS = p.RS, PW = p.PW, SL = p.SL;
    const size_t off0 = (sizeof(SweepShared) + 15) & ~size_t(15);
    uint4 *limbs = reinterpret_cast<uint4 *>(smem_raw + off0);                       // [8][RS]
    uint32_t *pkbuf = reinterpret_cast<uint32_t *>(smem_raw + off0 + size_t(kLimbs) * RS * 16);   // [2][64][RS]
    int *scratch = reinterpret_cast<int *>(pkbuf + size_t(2) * kB * RS);
    uint32_t *miss_list = reinterpret_cast<uint32_t *>(scratch + kScratchInts);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t gid = lane >> 2, tig = lane & 3;
    const uint32_t cta = blockIdx.x;
    const uint32_t B = p.B, K = p.K;
    const uint32_t NP = SL / 64;                     // 64-individual k-step pairs per slice
    const uint32_t NI = SL / 4;                      // register items (4 individuals each) per slice
    const uint32_t PWT = PW + kTail;                 // words per (marker, CTA) row in HBM: codes + missing list
    const uint32_t CB = PWT * 4;                     // bytes per row = one TMA bulk copy
    const double NM1 = double(p.n_total - 1), Nt = double(p.n_total);
    const uint32_t nBlocks = (p.n_order + B - 1) / B;
    const uint32_t slice0 = cta * SL;
    unsigned bar_gen = 0;                            // grid-wide rounds so far (identical in every thread of the grid)
    const bool prof_on = PROF && (p.prof != nullptr) && cta == 0 && tid == 0;
    const bool prof_w = PROF && (p.prof != nullptr) && cta == 0 && tid == 32 * kWalkWarps;
    long long t_prev = (prof_on || prof_w) ? clock64() : 0;
    if (prof_on) for (int i = 0; i < kProfPhases; ++i) sh.prof[i] = 0;
    auto PHASE = [&](int ph) {          // charge the cycles since the previous mark to phase `ph`
        if (prof_on) prof_mark(&sh.prof[ph], &t_prev, true);
    };
    auto PHASE_W = [&](int ph, bool charge) {   // the same for the worker probe
        if (prof_w) prof_mark(&sh.prof[ph], &t_prev, charge);
    };

    // ---- prologue -----------------------------------------------------------------------------
    // eps -> registers.  item = 4w + c owns the individuals 16w + 4t + c, t = 0..3
    double e[2][4];
    double amax = 0.0;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const uint32_t item = tid + it * kThreads;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint32_t i = 16 * (item >> 2) + 4 * t + (item & 3u);
            const uint32_t gi = slice0 + i;
            e[it][t] = (item < NI && gi < p.n_local) ? p.eps[gi] : 0.0;
            amax = fmax(amax, fabs(e[it][t]));
        }
    }
    fill_group_consts(sh.gc, p.nG, K, NM1, p.sigmaE, p.pi, p.sigmaG, p.cva);
    for (int i = tid; i < kMaxGK; i += kThreads) { sh.v_s[i] = 0; sh.betasq_s[i] = 0.0; }
    for (uint32_t i = tid; i < 2 * kB * RS; i += kThreads) pkbuf[i] = 0;
    if (tid < kB) { sh.rowslot[tid] = 0; sh.slotrow[tid] = 0; sh.carry[tid] = 0.0; }
    if (tid == 0) {
        mbar_init(&sh.mbar[0], 1); mbar_init(&sh.mbar[1], 1);
        mbar_fence_init();
        sh.n_nonskipped_total = 0; sh.n_ns = 0; sh.err = 0; sh.fin[0].n = 0; sh.fin[1].n = 0;
        for (int s = 0; s < kSlots; ++s) sh.lz_hist[s] = 0;
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

    // TMA prefetch of one block of code rows: every dot warp requests kTmaLanes rows, from marker ids its lanes loaded
    // earlier (order_ids: the latency of that load is off the critical path).  Nobody waits for the issue: the consumers
    // of the rows wait on the buffer's mbarrier.  Rows past the block's markers (last block only) are zero-filled so
    // that the fixed 64-row tile structure sees g = 0 there; the caller then syncs the CTA.
    uint32_t pf_id = 0u;
    const uint32_t pf_row = (warp >= kWalkWarps && lane < kTmaLanes) ? uint32_t(warp - kWalkWarps) * kTmaLanes + lane : 0xffffffffu;
    auto order_ids = [&](uint32_t b) {
        if (pf_row < B && b < nBlocks) {
            const uint32_t pos = b * B + pf_row;
            pf_id = pos < p.n_order ? p.order[pos] : 0u;
        }
    };
    auto prefetch = [&](uint32_t b) {
        const uint32_t nb = min(B, p.n_order - b * B);
        uint32_t *dst = pkbuf + size_t(b & 1) * kB * RS;
        uint64_t *bar = &sh.mbar[b & 1];
        if (warp == kWalkWarps) {                     // the first dot warp arms the barrier before anybody's bytes can land
            fence_proxy_async();
            if (lane == 0) mbar_arrive_expect_tx(bar, nb * CB);
        }
        if (warp >= kWalkWarps) {
            asm volatile("bar.sync 3, %0;" ::"n"(32 * kDotWarps) : "memory");
            if (pf_row < nb) {
                fence_proxy_async();
                bulk_g2s(dst + size_t(pf_row) * RS, p.codes + (size_t(pf_id) * p.nCTA + cta) * PWT, CB, bar);
            }
            if (nb < B)
                for (uint32_t i = (warp - kWalkWarps) * 32 + lane; i < (B - nb) * PWT; i += kDotWarps * 32)
                    dst[size_t(nb + i / PWT) * RS + i % PWT] = 0;
        }
    };
    if (nBlocks > 0) { order_ids(0); prefetch(0); }      // block 1 follows from the (virtual) first pass of the block loop

    // ---- cross-GPU exchange (world > 1) -----------------------------------------------------------------
    // Individuals are sharded over the GPUs; the integer sums of a grid-wide round must be summed over all of
    // them before they are used.  After the LOCAL grid barrier of a round, the exchange CTAs (CTA 0 for the
    // small per-block message; the first xchg_ctas CTAs share the Gram rows of a lazy round) store their part
    // into their mailbox on every peer (plain stores over NVLink), publish a round id (st.release.sys) in the
    // peer's flag array, wait for the peers' flags, add the peers' parts into the local slot, and arrive on a
    // second local counter that releases everybody.  Sums are integers: the result does not depend on the
    // arrival order or on the number of GPUs.  `t`/`nthr`/`sync` name the thread group that runs the round
    // (the 64 walker threads for the per-block message, the whole CTA otherwise).
    // The per-block message (every block, on the critical path) takes a shorter way: CTA 0 pushes and publishes,
    // the walkers of EVERY CTA wait for the peers' round ids themselves and add the peers' mailboxes to the local
    // sums as they read them (peer_sum) -- no merge, no second counter.
    const unsigned long long round_base = (world > 1) ? *p.round_ctr : 0ULL;
    unsigned xround = 0;                                    // exchange rounds done in this launch
    unsigned bar2_target = 0;
    int main_rot = 0;                                       // mailbox rotation of the current per-block round
    unsigned mcount = 0;                                    // per-block rounds done in this launch (their mailbox rotation)
    unsigned lcount = 0;                                    // Gram-row rounds sent in LL form in this launch
    uint32_t ll_rows = 0;                                   // Gram rows of the current block sent in LL form (not merged in the global slot)
    auto xchg = [&](int kind, int *slot, uint32_t r0, uint32_t nrows, int t, int nthr, auto &&sync) {
        if (world <= 1) return;
        ++xround;
        const unsigned long long rid = round_base + xround;
        const int rot = int(rid % kXchgRot);
        const uint32_t E = (kind == kXLazy) ? p.xchg_ctas : 1u;
        if (kind == kXMain) {
            // Every GPU keeps, per rotation, one 64-bit ACCUMULATOR per value of the round (at the head of the LL region); the peers
            // ADD their reduced sums into it over NVLink (red.add, value + kCnt1), the push spread over all CTAs (value i by CTA
            // i mod nCTA, one thread per (value, peer)).  The walkers of every CTA then poll ONE word per value until it carries the
            // count of all peers (read-back below): the cost of a round does not grow with the number of GPUs.  The accumulators
            // of block b-2 are cleared with its slot; all of them at the start of a launch.
            main_rot = int(mcount++ % kXchgRot);
            {
                const uint32_t np = MG ? uint32_t(world - 1) : 1u;       // (never 0: this code is dead in the single-GPU instantiations)
                const uint32_t nv = cta < uint32_t(kMainVals) ? (uint32_t(kMainVals) - 1u - cta) / p.nCTA + 1u : 0u;
                for (uint32_t idx = t; idx < nv * np; idx += nthr) {
                    const uint32_t k = idx / np, pj = idx - k * np, pr = pj < uint32_t(p.rank) ? pj : pj + 1u, i = cta + k * p.nCTA;
                    const unsigned long long v = __ldcg(reinterpret_cast<const unsigned long long *>(slot + kSlotDots + i * kLine));
                    red_sys_add_u64(reinterpret_cast<unsigned long long *>(p.peer_xchg[pr] + xchg_ll_off_ints(world)) + size_t(main_rot) * kMainVals + i, v + kCnt1);
                }
            }
            return;
        }
        bar2_target += E;
        if (cta < E) {
            unsigned long long *mine = reinterpret_cast<unsigned long long *>(p.peer_xchg[p.rank]);
            uint32_t lo, hi;                                // my range of message elements (uint64 units)
            if (kind == kXAbs) { lo = kMsgAbs; hi = kMsgAbs + 1; }
            else if (kind == kXMain) { lo = kMsgDots; hi = kMsgAbs; }
            else { const uint32_t tot = nrows * 2 * kCols, per = (tot + E - 1) / E; lo = min(tot, cta * per); hi = min(tot, lo + per); }
            auto local_ptr = [&](uint32_t i) -> unsigned long long * {
                if (i < uint32_t(kMsgDots)) return reinterpret_cast<unsigned long long *>(slot + kSlotLazy + (size_t(r0) * kCols + (i >> 1)) * kLine) + (i & 1u);
                if (i < uint32_t(kMsgSm)) return reinterpret_cast<unsigned long long *>(slot + kSlotDots + (i - kMsgDots) * kLine);
                if (i < uint32_t(kMsgEps)) return reinterpret_cast<unsigned long long *>(slot + kSlotSm + (i - kMsgSm) * kLine);
                if (i < uint32_t(kMsgAbs)) return reinterpret_cast<unsigned long long *>(slot + kSlotEps + (i - kMsgEps) * kLine);
                return p.absmax;
            };
            // 1. push
            for (uint32_t i = lo + t; i < hi; i += nthr) {
                const unsigned long long v = __ldcg(local_ptr(i));
                for (int pr = 0; pr < world; ++pr)
                    if (pr != p.rank) reinterpret_cast<unsigned long long *>(p.peer_xchg[pr])[(size_t(p.rank) * kXchgRot + rot) * kMsgU64 + i] = v;
            }
            __threadfence_system();
            sync();
            // 2. publish my round id on every peer, wait for theirs
            if (t < world && t != p.rank) {
                unsigned long long *pf = reinterpret_cast<unsigned long long *>(p.peer_xchg[t] + xchg_mailbox_ints(world)) + size_t(p.rank) * kXchgCtas + cta;
                st_release_sys_u64(pf, rid);
                const unsigned long long *lf = reinterpret_cast<const unsigned long long *>(p.peer_xchg[p.rank] + xchg_mailbox_ints(world)) + size_t(t) * kXchgCtas + cta;
                const long long t0 = clock64();
                while (ld_acquire_sys_u64(lf) < rid) {
                    if ((sh.err & 2u) || clock64() - t0 > 6000000000LL) { sh.err |= 2u; break; }   // ~3 s: a peer is gone; fail, do not hang
                }
            }
            sync();
            // 3. merge the peers' parts into the local slot (integers: any order gives the same sums)
            for (uint32_t i = lo + t; i < hi; i += nthr) {
                unsigned long long *lp = local_ptr(i);
                unsigned long long v = __ldcg(lp);
                for (int pr = 0; pr < world; ++pr)
                    if (pr != p.rank) {
                        const unsigned long long o = __ldcg(mine + (size_t(pr) * kXchgRot + rot) * kMsgU64 + i);
                        if (kind == kXAbs) v = v > o ? v : o;          // bits of non-negative doubles order like integers
                        else v += o;
                    }
                *lp = v;
            }
            __threadfence();
            sync();
            if (t == 0) red_release_add(p.bar2, 1u);
        }
        if (t == 0) {
            const long long t0 = clock64();
            while (ld_acquire(p.bar2) < bar2_target) {
                if (clock64() - t0 > 8000000000LL) { sh.err |= 2u; break; }
            }
        }
        sync();
    };
    auto cta_sync = [&]() { __syncthreads(); };
    auto wlk_sync = [&]() { walker_sync(); };

    if (MG && world > 1 && cta == 0) {                        // all per-block accumulators start at zero (the peers add only after the exchange below)
        unsigned long long *acc = reinterpret_cast<unsigned long long *>(p.peer_xchg[p.rank] + xchg_ll_off_ints(world));
        for (int i = tid; i < kXchgRot * kMainVals; i += kThreads) acc[i] = 0ULL;
        __threadfence_system();
    }
    if (MG && kLazyAcc && world > 1) {                        // ... and the Gram-row accumulators (all CTAs share the 512 KB)
        unsigned long long *accl = reinterpret_cast<unsigned long long *>(p.peer_xchg[p.rank] + xchg_lazy_ll_off_ints(world));
        for (uint32_t i = cta * kThreads + tid; i < uint32_t(kXchgRot * kLazyCells); i += p.nCTA * kThreads) accl[i] = 0ULL;
        if (tid < kXchgRot) sh.lza_hist[tid] = 0u;
        __threadfence_system();
    }
    grid_arrive(p.bar);
    grid_wait(p.bar, (++bar_gen) * p.nCTA);
    xchg(kXAbs, nullptr, 0, 0, tid, kThreads, cta_sync);   // max |eps| over all GPUs
    {   // common fixed-point scale: |q| <= 2^(62 - headroom) at the start of the sweep
        const double a = __longlong_as_double((long long)__ldcg(p.absmax));
        int ex = 0;
        if (a > 0.0) frexp(a, &ex);
        const int s = 62 - kHeadroomBits - ex;
        if (tid == 0) { sh.scale = ldexp(1.0, s); sh.inv_scale = ldexp(1.0, -s); }
    }
    __syncthreads();
    const double scale = sh.scale, inv_scale = sh.inv_scale;

    // eps registers -> limbs in shared memory, and this slice's sum of q
    auto quantize = [&]() {
        long long shi = 0; unsigned long long slo = 0;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const uint32_t item = tid + it * kThreads;
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
        // block-wide integer sums (order independent)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            shi += __shfl_xor_sync(0xffffffffu, shi, o);
            slo += __shfl_xor_sync(0xffffffffu, slo, o);
        }
        if (lane == 0) { sh.wq[warp][0] = (unsigned long long)shi; sh.wq[warp][1] = slo; }
    };
    auto slice_q = [&](unsigned long long &hi, unsigned long long &lo) {     // this slice's sum of q (after a barrier)
        hi = 0; lo = 0;
        for (int w = 0; w < kWarps; ++w) { hi += sh.wq[w][0]; lo += sh.wq[w][1]; }
    };
    quantize();
    __syncthreads();                 // the limbs are read by other warps right away (found by compute-sanitizer racecheck)

    // ---- missing-genotype side pass --------------------------------------------------------------
    // Every code row carries its missing list (count + first 15 slice-local ids) in its tail.  A warp
    // visits the missing entries of a row in chunks copied to its private scratch segment: one chunk from
    // the tail when the count fits, else by scanning the row 16 words (<= 256 entries) at a time.
    auto for_missing_chunks = [&](const uint32_t *row, auto &&fn) {
        uint32_t *seg = miss_list + warp * kMissSeg;
        const uint16_t *tail = reinterpret_cast<const uint16_t *>(row + PW);
        const uint32_t cnt = tail[0];
        if (cnt == 0) return;
        if (cnt <= 15) {
            if (lane < 20) seg[lane] = (uint32_t(lane) < cnt) ? uint32_t(tail[1 + lane]) : 0xffffffffu;   // + sentinels
            __syncwarp();
            fn(seg, cnt);
            __syncwarp();
            return;
        }
        for (uint32_t w0 = 0; w0 < PW; w0 += 16) {
            const uint32_t w = w0 + lane;
            const uint32_t x = (lane < 16 && w < PW) ? row[w] : 0u;
            uint32_t mm = x & (x >> 1) & 0x55555555u;
            const uint32_t n = __popc(mm);
            unsigned bal = __ballot_sync(0xffffffffu, n != 0);
            uint32_t tot = 0;
            while (bal) {
                const int l = __ffs(bal) - 1;
                bal &= bal - 1;
                const uint32_t nl = __shfl_sync(0xffffffffu, n, l);
                if (lane == l) {
                    uint32_t pos = tot;
                    while (mm) {
                        const int bit = __ffs(mm) - 1;
                        mm &= mm - 1;
                        seg[pos++] = 16 * w + (bit >> 1);
                    }
                }
                tot += nl;
            }
            if (lane < 4) seg[tot + lane] = 0xffffffffu;          // sentinels (kMissSeg leaves room for them)
            __syncwarp();
            if (tot) fn(seg, tot);
            __syncwarp();
        }
    };
    // sum of q over this warp's listed entries (lane partials, exact: integers)
    auto missing_q_sum = [&](const uint32_t *ent, uint32_t n, long long &shi, unsigned long long &slo) {
        for (uint32_t en = lane; en < n; en += 32) {
            const uint32_t i = ent[en];
            const uint32_t w = i >> 4, c = i & 3u, t = (i >> 2) & 3u;
            const uint32_t *lw = reinterpret_cast<const uint32_t *>(limbs) + c;
            unsigned long long uq = 0;
#pragma unroll
            for (int l = 0; l < kLimbs; ++l)
                uq |= (unsigned long long)((lw[(size_t(l) * RS + w) * 4] >> (8 * t)) & 0xffu) << (8 * l);
            const long long q = (long long)((uq ^ 0x8080808080808080ULL) - 0x8080808080808080ULL);
            shi += q >> 32; slo += (unsigned long long)(q & 0xffffffffLL);
        }
    };
    // Between a dot pass and the publish of its block the head of the scratch holds the block's sums: the 64 x 8 limb sums
    // (512 ints) and 64 x 2 sums of q over the rows' missing individuals; behind them the dot warps' partial regions.
    // Critical-path code runs on one or two warps at ~10 cycles per instruction (measured), so everything that can be is
    // done by the 14 dot warps right after their pass instead of in the publish.
    int *dot_acc = scratch;
    unsigned long long *msum = reinterpret_cast<unsigned long long *>(scratch + 512);
    int *dot_part = scratch + 1024;                  // [kDotWarps][128]: a dot warp's partial sums of its tile

    // ---- dot pass of one block: S[m][l] = sum_i g'_mi * limb_l(q_i) ---------------------------------------------
    // Work items = (marker tile mi, word pair wp): a word pair is two 64-individual k-steps.  Dot warp `wi` works on ONE tile,
    // mi = wi & 3 (tiles 0,1 have four warps, tiles 2,3 three), and takes every n_t-th word pair of it: one address set-up
    // and one epilogue per warp and block instead of four (these passes are bound by the per-warp instruction chain, not
    // by the tensor pipe), and only 3-4 partials per output to combine.  One item = four MMAs, one per code class
    // (kC0..kC3), each fed by the class-c bytes of the lane's two code words of both rows (A) and the class-c limb bytes of
    // the same individuals (B): four independent chains, 16 LOP3 and no shift per item.  The operands of the next item are
    // loaded before the MMAs of the current one are issued (software pipeline by hand: the compiler does not).
    auto dot_partial = [&](const uint32_t *pk, uint32_t nb, uint32_t wi) {
        const uint32_t NF = NP / 2;                                        // full word pairs; an odd NP leaves a half pair behind them
        const uint32_t mi = wi & 3u, r = wi >> 2, n_t = (mi < 2u) ? 4u : 3u;
        const uint32_t lo = gid * RS + tig;
        int acc[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[c][0] = 0; acc[c][1] = 0; acc[c][2] = 0; acc[c][3] = 0; }
        if (16u * mi < nb) {
            uint32_t aR0 = smem_u32(pk + size_t(16 * mi) * RS + lo + 8 * r);       // row gid of the tile, this lane's word of pair r
            const uint32_t dR1 = 32u * RS;                                         // row gid + 8 (bytes)
            uint32_t aL = smem_u32(limbs + lo + 8 * r);
            const uint32_t stepA = 32u * n_t, stepL = 128u * n_t;
            uint32_t cnt = NF > r ? (mi < 2u ? (NF - r + 3u) >> 2 : (NF - r + 2u) / 3u) : 0u;   // n_t is 3 or 4: no runtime division
            if (cnt) {
                uint4 LA = lds_v4<0>(aL), LB = lds_v4<64>(aL);
                uint32_t a0 = lds_u32<0>(aR0), a1 = lds_u32<0>(aR0 + dR1), b0 = lds_u32<16>(aR0), b1 = lds_u32<16>(aR0 + dR1);
#pragma unroll 1
                for (;;) {
                    const bool more = --cnt != 0u;
                    if (more) { aR0 += stepA; aL += stepL; }                       // the last item re-loads itself (dropped)
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
            if ((NP & 1u) && (mi < 2u ? (NF & 3u) : NF % 3u) == r) {                                      // the half pair at the end of the slice
                const uint32_t *r0p = pk + size_t(16 * mi) * RS + lo + 8 * NF;
                const uint4 LA = limbs[lo + 8 * NF];
                const uint32_t a0 = r0p[0], a1 = r0p[8 * RS];
                mma_u8s8(acc[0], a0 & kC0, a1 & kC0, 0u, 0u, LA.x, 0u);
                mma_u8s8(acc[1], a0 & kC1, a1 & kC1, 0u, 0u, LA.y, 0u);
                mma_u8s8(acc[2], a0 & kC2, a1 & kC2, 0u, 0u, LA.z, 0u);
                mma_u8s8(acc[3], a0 & kC3, a1 & kC3, 0u, 0u, LA.w, 0u);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            dot_part[wi * 128 + q * 32 + lane] = acc[0][q] + (acc[1][q] >> 2) + (acc[2][q] >> 4) + (acc[3][q] >> 6);
    };
    // ... and, once all dot warps have stored theirs (a barrier among them), the sums over the warps of each tile
    auto dot_combine = [&](uint32_t wi) {
        asm volatile("bar.sync 3, %0;" ::"n"(32 * kDotWarps) : "memory");
        const uint32_t o = wi * 32 + lane;               // 448 threads, 512 outputs
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t oo = o + h * 32 * kDotWarps;
            if (oo < 512) {
                const uint32_t mi = oo >> 7, e7 = oo & 127u;
                int sacc = dot_part[mi * 128 + e7] + dot_part[(4 + mi) * 128 + e7] + dot_part[(8 + mi) * 128 + e7];
                if (mi < 2u) sacc += dot_part[(12 + mi) * 128 + e7];
                dot_acc[oo] = sacc;
            }
        }
    };
    // sums of q over the missing individuals of the block's rows (exact: the rows' lists are in their tails): row group rg
    // takes rows 8rg..8rg+7, 4 lanes per row, lane e4 the entries e4, e4+4, ... of the tail's list (<= 15 entries); longer
    // lists go the general way.  Every row < nb gets its pair written.  The 8 groups are taken by the 8 dot warps of marker
    // tiles 0 and 1 (7 dot items each against the 9 of the other six warps) right after their dot pass.
    auto missing_sums = [&](const uint32_t *pkn, uint32_t nbn, uint32_t rg) {
        const uint32_t r8 = lane >> 2, e4 = lane & 3u, m2 = 8 * rg + r8;
        uint32_t cnt = 0;
        long long qh = 0; unsigned long long ql = 0;
        if (m2 < nbn) {
            const uint16_t *tail = reinterpret_cast<const uint16_t *>(pkn + size_t(m2) * RS + PW);
            cnt = tail[0];
            if (cnt <= 15u) {
#pragma unroll 1
                for (uint32_t en = e4; en < cnt; en += 4) {
                    const uint32_t i = tail[1 + en];
                    // limb byte l of individual i: word (l*RS + i/16)*4 + i%4 of the limb array, byte (i/4)%4
                    const uint8_t *lb = reinterpret_cast<const uint8_t *>(limbs) + ((i >> 4) * 4 + (i & 3u)) * 4 + ((i >> 2) & 3u);
                    uint32_t ulo = 0, uhi = 0;
#pragma unroll
                    for (int l = 0; l < 4; ++l) {
                        ulo |= uint32_t(lb[size_t(l) * RS * 16]) << (8 * l);
                        uhi |= uint32_t(lb[size_t(4 + l) * RS * 16]) << (8 * l);
                    }
                    const unsigned long long uq = (unsigned long long)ulo | ((unsigned long long)uhi << 32);
                    const long long q = (long long)((uq ^ 0x8080808080808080ULL) - 0x8080808080808080ULL);
                    qh += q >> 32; ql += (unsigned long long)(q & 0xffffffffLL);
                }
            }
        }
#pragma unroll
        for (int o2 = 1; o2 < 4; o2 <<= 1) {
            qh += __shfl_xor_sync(0xffffffffu, qh, o2);
            ql += __shfl_xor_sync(0xffffffffu, ql, o2);
        }
        if (e4 == 0 && m2 < nbn && cnt <= 15u) { msum[2 * m2] = (unsigned long long)qh; msum[2 * m2 + 1] = ql; }
        unsigned big = __ballot_sync(0xffffffffu, cnt > 15u && e4 == 0);
        while (big) {
            const uint32_t m3 = 8 * rg + uint32_t(__ffs(big) - 1) / 4u;
            big &= big - 1;
            long long shi = 0; unsigned long long slo = 0;
            for_missing_chunks(pkn + size_t(m3) * RS, [&](const uint32_t *ent, uint32_t n) { missing_q_sum(ent, n, shi, slo); });
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                shi += __shfl_xor_sync(0xffffffffu, shi, o);
                slo += __shfl_xor_sync(0xffffffffu, slo, o);
            }
            if (lane == 0) { msum[2 * m3] = (unsigned long long)shi; msum[2 * m3 + 1] = slo; }
        }
    };
    // Publish block bn's sums (limb sums, sums over the missing individuals, the slice's sum of q) into the block's global
    // slot and arrive at the block's grid barrier.  All threads call, after a __syncthreads that follows the dot pass.
    auto publish_dots = [&](uint32_t bn) {
        const uint32_t nbn = min(B, p.n_order - bn * B);
        int *slotn = p.slots + size_t(bn % kSlots) * kSlotInts;
        if (tid < 2 * kB) {
            // (marker m, half h): the four limb sums 4h..4h+3 of the marker combined into sum_l S_l 2^(8l) -- exact, < 2^46 per CTA --
            // as ONE signed 64-bit RED.  dot_acc holds the MMA C fragments: limb 2 tig + (q & 1) of marker 16 mi + gid + 8 (q >> 1).
            const uint32_t m = tid >> 1, h = tid & 1u;
            const int *da = dot_acc + ((m >> 4) * 4 + ((m >> 3) & 1u) * 2) * 32 + (m & 7u) * 4 + 2 * h;
            const long long v = (long long)da[0] + ((long long)da[32] << 8) + ((long long)da[1] << 16) + ((long long)da[33] << 24);
            if (v != 0 && m < nbn)
                atomicAdd(reinterpret_cast<unsigned long long *>(slotn + kSlotDots + tid * kLine), (unsigned long long)v);
        } else if (tid < 4 * kB) {
            const uint32_t i = tid - 2 * kB;              // (marker, hi/lo part)
            if ((i >> 1) < nbn) {
                const unsigned long long v = msum[i];
                if (v) atomicAdd(reinterpret_cast<unsigned long long *>(slotn + kSlotSm + i * kLine), v);
            }
        } else if (warp == 8) {               // lane 2w+h holds part h of warp w's sum of q
            unsigned long long v = sh.wq[lane >> 1][lane & 1];
#pragma unroll
            for (int o2 = 2; o2 < 32; o2 <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
            if (lane < 2) atomicAdd(reinterpret_cast<unsigned long long *>(slotn + kSlotEps + lane * kLine), v);
        }
        PHASE(kProfPubReduce);
        grid_arrive(p.bar);
        PHASE(kProfPubArrive);
    };

    // ---- Gram rows on demand ------------------------------------------------------------------------------------
    // One grid-wide round: the integer pair counts of the rows of the markers in (need0, need1) (bit masks over the
    // two half-blocks of block b) against 128 columns -- the 64 markers of block b and the 64 of block b+1 (whose
    // dots were taken against the residual BEFORE block b's updates and are corrected through these rows) -- reduced
    // over CTAs / GPUs.  The block-b half is standardised into x_j.x_k and stored as doubles in row slots r0, r0+1, ...
    // of the scratch; the block-(b+1) half stays in the global slot for the carry pass.  Every CTA calls with
    // identical arguments.  nbx = markers of block b+1 (0 = there is none).
    auto lazy_round = [&](const uint32_t *pk, const uint32_t *pkx, int *slot, const BlockMeta &meta, const BlockMeta &metax, uint32_t nb,
                          uint32_t nbx, uint32_t need0, uint32_t need1, uint32_t r0, bool keep_next) {
        const uint32_t n0 = __popc(need0), nnew = n0 + __popc(need1);
        if (tid < kB) {
            const bool hi = tid >= 32;
            const uint32_t msk = hi ? need1 : need0, bit = 1u << (tid & 31);
            if (msk & bit) {
                const uint32_t pos = (hi ? n0 : 0u) + __popc(msk & (bit - 1u));
                sh.lz_rows[pos] = tid; sh.rowslot[tid] = r0 + pos; sh.slotrow[r0 + pos] = tid;
            }
        }
        __syncthreads();
        PHASE(kProfLzC0);
        // (1) missing-genotype cross counts, all warps: warp w takes the columns a = 8w..8w+7, lane = (column, row subset);
        //     a lane handles the pairs (row idx0 + s4, column a):
        //       from the row's list:     cz += g_a(i) (a present), cw += 1 (a missing too),  i in missing(j)
        //       from the column's list:  cy += g_j(i) (j present),                           i in missing(a)
        //     both read straight from the rows' tails (<= 15 entries); pairs with a longer list are redone by the whole warp.
        {
            const uint32_t a = 8 * warp + (lane >> 2), s4 = lane & 3u;
            const bool okcol = (a < uint32_t(kB)) ? (a < nb) : (a - kB < nbx);
            const uint32_t *rowa = (a < uint32_t(kB)) ? pk + size_t(a) * RS : pkx + size_t(a - kB) * RS;
            const uint16_t *tla = reinterpret_cast<const uint16_t *>(rowa + PW);
            const uint32_t cnta = okcol ? uint32_t(tla[0]) : 0u;
            // the four lanes of a column take R rows at a time, 4/R lanes per (row, column) pair sharing the list entries:
            // R = 1 for the one or two rows of a typical round (short dependent chains), R = 4 when many rows are new
            auto cross = [&](auto rc) {
                constexpr uint32_t R = decltype(rc)::value, ES = 4u / R;
                const uint32_t rsub = s4 / ES, e0 = s4 % ES;
                for (uint32_t idx0 = 0; idx0 < nnew; idx0 += R) {
                    const uint32_t idx = idx0 + rsub;
                    const bool on = okcol && idx < nnew;
                    const uint32_t j = on ? sh.lz_rows[idx] : 0u;
                    const bool pair = on && a != j;                            // a >= 64 never equals a row marker
                    const uint32_t *rowj = pk + size_t(j) * RS;
                    const uint16_t *tlj = reinterpret_cast<const uint16_t *>(rowj + PW);
                    const uint32_t cntj = pair ? uint32_t(tlj[0]) : 0u;
                    const bool slow = pair && (cntj > 15u || cnta > 15u);
                    uint32_t zwy = 0;                                          // z | wv << 8 | y << 16 (each <= 2 * 15)
                    if (pair && !slow) {
                        for (uint32_t en = e0; en < cntj; en += ES) {
                            const uint32_t i = tlj[1 + en];
                            const uint32_t c = (rowa[i >> 4] >> (2 * (i & 15u))) & 3u;
                            zwy += (c == 3u) ? 0x100u : c;
                        }
                        for (uint32_t en = e0; en < cnta; en += ES) {
                            const uint32_t i = tla[1 + en];
                            const uint32_t c = (rowj[i >> 4] >> (2 * (i & 15u))) & 3u;
                            zwy += (c == 3u) ? 0u : (c << 16);
                        }
                    }
#pragma unroll
                    for (uint32_t o2 = 1; o2 < ES; o2 <<= 1) zwy += __shfl_xor_sync(0xffffffffu, zwy, o2);
                    int z = int(zwy & 0xffu), wv = int((zwy >> 8) & 0xffu), y = int(zwy >> 16);
                    unsigned slow_m = __ballot_sync(0xffffffffu, slow && e0 == 0u);
                    while (slow_m) {
                        const int l = __ffs(slow_m) - 1;
                        slow_m &= slow_m - 1;
                        const uint32_t a2 = __shfl_sync(0xffffffffu, a, l), j2 = __shfl_sync(0xffffffffu, j, l);
                        const uint32_t *ra = (a2 < uint32_t(kB)) ? pk + size_t(a2) * RS : pkx + size_t(a2 - kB) * RS;
                        const uint32_t *rj = pk + size_t(j2) * RS;
                        int zz = 0, ww = 0, yy = 0;
                        for_missing_chunks(rj, [&](const uint32_t *ent, uint32_t n) {
                            for (uint32_t en = lane; en < n; en += 32) {
                                const uint32_t i = ent[en];
                                const uint32_t c = (ra[i >> 4] >> (2 * (i & 15u))) & 3u;
                                zz += (c == 3u) ? 0 : int(c);
                                ww += (c == 3u) ? 1 : 0;
                            }
                        });
                        for_missing_chunks(ra, [&](const uint32_t *ent, uint32_t n) {
                            for (uint32_t en = lane; en < n; en += 32) {
                                const uint32_t i = ent[en];
                                const uint32_t c = (rj[i >> 4] >> (2 * (i & 15u))) & 3u;
                                yy += (c == 3u) ? 0 : int(c);
                            }
                        });
                        zz = __reduce_add_sync(0xffffffffu, zz); ww = __reduce_add_sync(0xffffffffu, ww); yy = __reduce_add_sync(0xffffffffu, yy);
                        if (lane == l) { z = zz; wv = ww; y = yy; }
                    }
                    if (pair && e0 == 0u) {
                        unsigned long long *ln = reinterpret_cast<unsigned long long *>(slot + kSlotLazy + (size_t(r0 + idx) * kCols + a) * kLine);
                        if (wv) atomicAdd(ln, (unsigned long long)(unsigned)wv << 32);
                        if (y | z) atomicAdd(ln + 1, (unsigned long long)(unsigned)y | ((unsigned long long)(unsigned)z << 32));
                    }
                }
            };
            if (nnew <= 2u) cross(std::integral_constant<uint32_t, 1u>{}); else cross(std::integral_constant<uint32_t, 4u>{});
        }
        PHASE(kProfLzC2);
        __syncthreads();                       // the warps' missing-list segments are free: reused below
        PHASE(kProfLazyCross);
        // (2) integer products, shaped like the dot pass: A = 16 column markers (tile mt of the 128 columns), B = 8 new
        //     rows (a group) in place of the 8 limbs.  warp = (column tile mt, half kh of the word pairs); one group per
        //     pass; per word pair four MMAs, one per code class, fed by in-place class bytes of both operands (16^c scaling).
        {
            const uint32_t mt = warp & 7u, kh = warp >> 3;
            const uint32_t NF = NP / 2;                                        // full word pairs (see dot_partial)
            const uint32_t first_row = need0 ? uint32_t(__ffs(need0) - 1) : 32u + uint32_t(__ffs(need1) - 1);   // its own column and all before it are not needed
            const bool mt_ok = (mt < 4) ? (16 * mt < nb && 16 * mt + 15 > first_row) : (16 * (mt - 4) < nbx);
            const uint32_t *c0p = ((mt < 4) ? pk + size_t(16 * mt) * RS : pkx + size_t(16 * (mt - 4)) * RS) + gid * RS + tig;
            const uint32_t dC1 = 32u * RS;                                     // row gid + 8 (bytes)
            int *sc2 = reinterpret_cast<int *>(miss_list);
            for (uint32_t g0 = 0; g0 * 8 < nnew; ++g0) {
                const uint32_t i0 = g0 * 8 + gid;
                const bool v0 = i0 < nnew;
                const uint32_t *u0p = pk + size_t(v0 ? sh.lz_rows[i0] : 0u) * RS + tig;
                const uint32_t m0 = v0 ? kC0 : 0u, m1 = v0 ? kC1 : 0u, m2 = v0 ? kC2 : 0u, m3 = v0 ? kC3 : 0u;   // an empty row slot feeds zeros
                int acc[4][4];
#pragma unroll
                for (int c = 0; c < 4; ++c) { acc[c][0] = 0; acc[c][1] = 0; acc[c][2] = 0; acc[c][3] = 0; }
                if (mt_ok) {
                    uint32_t aC = smem_u32(c0p + 8 * kh), aU = smem_u32(u0p + 8 * kh);
                    uint32_t cnt = NF > kh ? (NF - kh + 1u) >> 1 : 0u;
                    if (cnt) {
                        uint32_t a0 = lds_u32<0>(aC), a1 = lds_u32<0>(aC + dC1), b0 = lds_u32<16>(aC), b1 = lds_u32<16>(aC + dC1);
                        uint32_t uA = lds_u32<0>(aU), uB = lds_u32<16>(aU);
#pragma unroll 1
                        for (;;) {
                            const bool more = --cnt != 0u;
                            if (more) { aC += 64u; aU += 64u; }                    // the last item re-loads itself (dropped)
                            const uint32_t a0n = lds_u32<0>(aC), a1n = lds_u32<0>(aC + dC1), b0n = lds_u32<16>(aC), b1n = lds_u32<16>(aC + dC1);
                            const uint32_t uAn = lds_u32<0>(aU), uBn = lds_u32<16>(aU);
                            mma_u8u8(acc[0], a0 & kC0, a1 & kC0, b0 & kC0, b1 & kC0, uA & m0, uB & m0);
                            mma_u8u8(acc[1], a0 & kC1, a1 & kC1, b0 & kC1, b1 & kC1, uA & m1, uB & m1);
                            mma_u8u8(acc[2], a0 & kC2, a1 & kC2, b0 & kC2, b1 & kC2, uA & m2, uB & m2);
                            mma_u8u8(acc[3], a0 & kC3, a1 & kC3, b0 & kC3, b1 & kC3, uA & m3, uB & m3);
                            if (!more) break;
                            a0 = a0n; a1 = a1n; b0 = b0n; b1 = b1n; uA = uAn; uB = uBn;
                        }
                    }
                    if ((NP & 1u) && (NF & 1u) == kh) {                            // the half pair at the end of the slice
                        const uint32_t a0 = c0p[8 * NF], a1 = c0p[8 * NF + 8 * RS], uA = u0p[8 * NF];
                        mma_u8u8(acc[0], a0 & kC0, a1 & kC0, 0u, 0u, uA & m0, 0u);
                        mma_u8u8(acc[1], a0 & kC1, a1 & kC1, 0u, 0u, uA & m1, 0u);
                        mma_u8u8(acc[2], a0 & kC2, a1 & kC2, 0u, 0u, uA & m2, 0u);
                        mma_u8u8(acc[3], a0 & kC3, a1 & kC3, 0u, 0u, uA & m3, 0u);
                    }
                }
                int val[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) val[r] = acc[0][r] + (acc[1][r] >> 4) + (acc[2][r] >> 8) + (acc[3][r] >> 12);
                PHASE(kProfLzLoop);
                if (kh == 1) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) sc2[(mt * 4 + r) * 32 + lane] = val[r];
                }
                __syncthreads();
                if (kh == 0 && mt_ok) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int v = val[r] + sc2[(mt * 4 + r) * 32 + lane];
                        const uint32_t col = 16 * mt + gid + 8 * (r >> 1);                     // 0..127
                        const uint32_t ui = g0 * 8 + 2 * tig + (r & 1);                         // index of the new row
                        if (ui < nnew && v) {
                            const uint32_t rowm = sh.lz_rows[ui];
                            const bool okc = (col < uint32_t(kB)) ? (col < nb && col != rowm) : (col - kB < nbx);
                            if (okc)
                                atomicAdd(reinterpret_cast<unsigned long long *>(slot + kSlotLazy + (size_t(r0 + ui) * kCols + col) * kLine),
                                          (unsigned long long)(unsigned)v);
                        }
                    }
                }
                __syncthreads();
                PHASE(kProfLzRed);
            }
        }
        PHASE(kProfLazyMma);
        // (3) the grid-wide round
        if (MG && kLazyAcc && world > 1 && keep_next) {
            // the rotation this round's successor-but-one will use was last used two rounds ago: every CTA of this GPU has left that
            // round's read-back (it passed the previous round's barrier), and no peer adds into it before it has seen this GPU's next
            // round.  One CTA clears the rows used then; the arrive below publishes the zeros.
            const uint32_t zr = (lcount + 2u) % kXchgRot;
            if (cta == lcount % p.nCTA) {
                unsigned long long *accl = reinterpret_cast<unsigned long long *>(p.peer_xchg[p.rank] + xchg_lazy_ll_off_ints(world)) + size_t(zr) * kLazyCells;
                for (uint32_t i = tid; i < sh.lza_hist[zr] * 2u * kCols; i += kThreads) accl[i] = 0ULL;
            }
        }
        grid_arrive(p.bar);
        grid_wait(p.bar, (++bar_gen) * p.nCTA);
        // several GPUs: when the rows keep their next-block half in shared memory (keep_next) nobody reads the global slot after
        // this read-back, so the peers' counts are added as they are read (LL cells, one NVLink trip); else the merging exchange
        bool ll_round = false;
        uint32_t lazy_id = 0u;
        int lrot = 0;
        if (MG && world > 1 && keep_next) {
            ll_round = true;
            ++xround;
            lazy_id = uint32_t(round_base + 1u + lcount);
            lrot = int(lcount++ % kXchgRot);
            ll_rows += nnew;
            const uint32_t E = p.xchg_ctas;
            if (kLazyAcc) {
                const uint32_t totv = nnew * 2 * kCols, np = MG ? uint32_t(world - 1) : 1u;
                for (uint32_t idx = cta * kThreads + tid; idx < totv * np; idx += p.nCTA * kThreads) {
                    const uint32_t i = idx / np, pj = idx - i * np, pr = pj < uint32_t(p.rank) ? pj : pj + 1u;
                    const unsigned long long v = __ldcg(reinterpret_cast<const unsigned long long *>(slot + kSlotLazy + (size_t(r0) * kCols + (i >> 1)) * kLine) + (i & 1u));
                    red_sys_add_u64(reinterpret_cast<unsigned long long *>(p.peer_xchg[pr] + xchg_lazy_ll_off_ints(world)) + size_t(lrot) * kLazyCells + i, v + kCnt1);
                }
                if (tid == 0) sh.lza_hist[lrot] = nnew;         // (read two rounds later, behind CTA barriers)
            } else if (cta < E) {
                const uint32_t totv = nnew * 2 * kCols, per = (totv + E - 1) / E, vlo = min(totv, cta * per), vhi = min(totv, vlo + per);
                const unsigned long long idh = (unsigned long long)lazy_id << 32;
                for (uint32_t i = vlo + tid; i < vhi; i += kThreads) {
                    const unsigned long long v = __ldcg(reinterpret_cast<const unsigned long long *>(slot + kSlotLazy + (size_t(r0) * kCols + (i >> 1)) * kLine) + (i & 1u));
                    for (int pr = 0; pr < world; ++pr)
                        if (pr != p.rank)
                            st_sys_v2(reinterpret_cast<ulonglong2 *>(p.peer_xchg[pr] + xchg_lazy_ll_off_ints(world)) + (size_t(p.rank) * kXchgRot + lrot) * kLazyCells + i,
                                      (v & 0xffffffffULL) | idh, (v >> 32) | idh);
                }
            }
        } else {
            xchg(kXLazy, slot, r0, nnew, tid, kThreads, cta_sync);
        }
        PHASE(kProfLazyBarrier);
        // (4) read back and standardise into x_j . x_k (all loads of a thread in flight at once): the block-b half into row slot
        //     r of G; the block-(b+1) half, when the block's rows fit 32 slots (keep_next), into row slot 32 + r -- the carry pass
        //     then needs no second trip to L2
        double *G = reinterpret_cast<double *>(scratch);
        const uint32_t ncol = keep_next ? uint32_t(kCols) : uint32_t(kB), csh = keep_next ? 7u : 6u;
        const uint32_t tot = nnew * ncol;
        for (uint32_t base = 0; base < tot; base += 4 * kThreads) {
            int4 v[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const uint32_t en = base + tid + t * kThreads;
                if (en < tot) v[t] = __ldcg(reinterpret_cast<const int4 *>(slot + kSlotLazy + (size_t(r0 + (en >> csh)) * kCols + (en & (ncol - 1u))) * kLine));
            }
            if (MG && ll_round) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint32_t en = base + tid + t * kThreads;
                    if (en >= tot) continue;
                    const size_t ci = 2 * (size_t(en >> csh) * kCols + (en & (ncol - 1u)));
                    if (kLazyAcc) {
                        const unsigned long long *accl = reinterpret_cast<const unsigned long long *>(p.peer_xchg[p.rank] + xchg_lazy_ll_off_ints(world)) + size_t(lrot) * kLazyCells + ci;
                        unsigned long long c0, c1;
                        const long long t0 = clock64();
                        for (;;) {
                            c0 = ld_sys_u64(accl); c1 = ld_sys_u64(accl + 1);
                            if (unsigned(c0 >> 56) == unsigned(world - 1) && unsigned(c1 >> 56) == unsigned(world - 1)) break;
                            if ((sh.err & 2u) || clock64() - t0 > 6000000000LL) { sh.err |= 2u; break; }   // ~3 s: a peer is gone; fail, do not hang
                        }
                        c0 &= kCnt1 - 1ULL; c1 &= kCnt1 - 1ULL;
                        v[t].x += int(uint32_t(c0)); v[t].y += int(uint32_t(c0 >> 32)); v[t].z += int(uint32_t(c1)); v[t].w += int(uint32_t(c1 >> 32));
                        continue;
                    }
                    for (int pr = 0; pr < world; ++pr) {
                        if (pr == p.rank) continue;
                        const ulonglong2 *cell = reinterpret_cast<const ulonglong2 *>(p.peer_xchg[p.rank] + xchg_lazy_ll_off_ints(world)) + (size_t(pr) * kXchgRot + lrot) * kLazyCells + ci;
                        ulonglong2 c0, c1;
                        const long long t0 = clock64();
                        for (;;) {
                            c0 = ld_sys_v2(cell); c1 = ld_sys_v2(cell + 1);
                            if (uint32_t(c0.x >> 32) == lazy_id && uint32_t(c0.y >> 32) == lazy_id && uint32_t(c1.x >> 32) == lazy_id && uint32_t(c1.y >> 32) == lazy_id) break;
                            if ((sh.err & 2u) || clock64() - t0 > 6000000000LL) { sh.err |= 2u; break; }   // ~3 s: a peer is gone; fail, do not hang
                        }
                        v[t].x += int(uint32_t(c0.x)); v[t].y += int(uint32_t(c0.y)); v[t].z += int(uint32_t(c1.x)); v[t].w += int(uint32_t(c1.y));
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const uint32_t en = base + tid + t * kThreads;
                if (en < tot) {
                    const uint32_t idx = en >> csh, k = en & (ncol - 1u), j = sh.lz_rows[idx];
                    const double a = double(unsigned(v[t].x)), cw = double(unsigned(v[t].y)), cy = double(unsigned(v[t].z)), cz = double(unsigned(v[t].w));
                    if (k < uint32_t(kB)) G[(r0 + idx) * kB + k] = (k != j && k < nb) ? gram_entry(meta, int(j), meta, int(k), a, cw, cy, cz, Nt) : 0.0;
                    else G[(32 + r0 + idx) * kB + (k - kB)] = (k - kB < nbx) ? gram_entry(meta, int(j), metax, int(k - kB), a, cw, cy, cz, Nt) : 0.0;
                }
            }
        }
        __syncthreads();
        if (prof_on) { sh.prof[kProfLazyRounds] += 1; sh.prof[kProfLazyRows] += nnew; }
        PHASE(kProfLazyReadback);
    };

    // CTA 0 only, one warp: the previous block's results go to HBM (src/BayesRBase.cpp:341, :350-356 / :492-493) and into
    // the sweep's counts (:355 / :511) and sums of squares (:353 / :522, fixed summation order) -- off the critical path.
    auto flush_results = [&](BlockResult &fin) {
        const uint32_t n = fin.n;
        double b2[2]; int gk[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t k = lane + 32 * h;
            b2[h] = 0.0; gk[h] = -1;
            if (k < n) {
                const uint32_t id = fin.id[k];
                const int comp = fin.comp[k];
                if (!p.async_mode) p.acum[id] = fin.acum[k];
                if (comp >= 0) {
                    const double bn = fin.beta[k];
                    p.beta[id] = bn;
                    p.comp[id] = double(comp);
                    atomicAdd(&sh.v_s[fin.grp[k] * K + comp], 1);
                    if (comp > 0) { b2[h] = bn * bn; gk[h] = fin.grp[k]; }
                }
            }
        }
        unsigned pend0 = __ballot_sync(0xffffffffu, gk[0] >= 0), pend1 = __ballot_sync(0xffffffffu, gk[1] >= 0);
        while (pend0 | pend1) {                    // one pass per group that gained a non-zero effect in this block
            const int g = pend0 ? __shfl_sync(0xffffffffu, gk[0], __ffs(pend0) - 1) : __shfl_sync(0xffffffffu, gk[1], __ffs(pend1) - 1);
            const double s0 = warp_sum(gk[0] == g ? b2[0] : 0.0), s1 = warp_sum(gk[1] == g ? b2[1] : 0.0);
            if (lane == 0) sh.betasq_s[g] += s0 + s1;
            pend0 &= ~__ballot_sync(0xffffffffu, gk[0] == g);
            pend1 &= ~__ballot_sync(0xffffffffu, gk[1] == g);
        }
        __syncwarp();
        if (lane == 0) fin.n = 0;
    };
    constexpr int kFlushWarp = kWarps - 2;             // a dot warp of CTA 0

    // ---- the block loop.  Pass bb works on block bb while the dot warps run ahead on block bb+1; the first pass
    // (bb = -1) is a virtual empty block that only starts block 0, so that every piece of the pipeline exists once.
    unsigned gen_wait = 0;                                      // the grid-wide round that carries the current block's dots
    int zero_turn = int(cta);                                   // the next block whose pass recycles a slot on this CTA
    while (zero_turn < 2) zero_turn += int(p.nCTA);
    uint32_t gid_next = 0u;                                     // threads of warps kMetaWarp0, +1: marker id of their row of the next block to gather
    {
        const int k2 = tid - 32 * kMetaWarp0;
        if (uint32_t(k2) < min(B, p.n_order)) gid_next = p.order[k2];
    }
    PHASE(kProfPrologue);
    PHASE_W(0, false);

    for (int bb = -1; bb < int(nBlocks); ++bb) {
        const bool real = bb >= 0;
        const uint32_t b = real ? uint32_t(bb) : 0u;
        const uint32_t nb = real ? min(B, p.n_order - b * B) : 0u;
        const uint32_t buf = uint32_t(bb) & 1u;
        const uint32_t *pk = pkbuf + size_t(buf) * kB * RS;
        int *slot = p.slots + size_t(b % kSlots) * kSlotInts;
        BlockMeta &meta = sh.meta[buf];
        BlockMeta &meta_next = sh.meta[buf ^ 1u];
        const uint32_t bn = uint32_t(bb + 1);                   // the block the dot warps work on
        const bool have_next = bn < nBlocks;
        const uint32_t nb_next = have_next ? min(B, p.n_order - bn * B) : 0u;
        const uint32_t *pk_next = pkbuf + size_t(buf ^ 1u) * kB * RS;

        // ---- A. walkers: wait for the block's round, read it back, sample.  Workers: dot pass of block b+1 ------
        const bool walker = warp < kWalkWarps;
        const int wk = tid;                                   // marker of a walker thread
        const bool act = walker && uint32_t(wk) < nb;
        double bo = 0.0, uu = 2.0, zz = 0.0, base = 0.0; int gg = 0;
        double bo1 = 0.0, zz1 = 0.0, base1 = 0.0; int gg1 = 0; bool act1 = false;    // warp 0: marker 32+lane
        double corr_used = 0.0;
        uint32_t my_id = 0u;
        Draw dr; dr.beta_new = 0.0; dr.acum0 = 0.0; dr.comp = -1;
        if (walker && !real) {
            if (lane == 0) sh.w_nsm[warp] = 0u;
        } else if (walker) {
            if (tid == 0) { while (ld_acquire(p.bar) < gen_wait * p.nCTA) { } }
            walker_sync();
            PHASE(kProfWait);
            xchg(kXMain, slot, 0, 0, tid, 32 * kWalkWarps, wlk_sync);
            PHASE(kProfXchg);
            if (act) {
                // all six loads in flight at once, then the peers' parts
                long long tlo = __ldcg(reinterpret_cast<const long long *>(slot + kSlotDots + (2 * wk) * kLine));
                long long thi = __ldcg(reinterpret_cast<const long long *>(slot + kSlotDots + (2 * wk + 1) * kLine));
                long long sm_hi = __ldcg(reinterpret_cast<const long long *>(slot + kSlotSm + (2 * wk) * kLine));
                long long sm_lo = __ldcg(reinterpret_cast<const long long *>(slot + kSlotSm + (2 * wk + 1) * kLine));
                long long e_hi = __ldcg(reinterpret_cast<const long long *>(slot + kSlotEps));
                long long e_lo = __ldcg(reinterpret_cast<const long long *>(slot + kSlotEps + kLine));
                if (MG && world > 1) {                              // + the peers' sums: one accumulator word per value, complete at count world-1
                    const unsigned long long *acc = reinterpret_cast<const unsigned long long *>(p.peer_xchg[p.rank] + xchg_ll_off_ints(world)) + size_t(main_rot) * kMainVals;
                    const uint32_t ci[6] = {2u * wk, 2u * wk + 1u, 2u * kB + 2u * wk, 2u * kB + 2u * wk + 1u, 4u * kB, 4u * kB + 1u};
                    unsigned long long c[6];
                    const long long t0 = clock64();
                    for (;;) {
                        bool ok = true;
#pragma unroll
                        for (int q6 = 0; q6 < 6; ++q6) { c[q6] = ld_sys_u64(acc + ci[q6]); ok = ok && line_count(c[q6]) == unsigned(world - 1); }
                        if (ok) break;
                        if ((sh.err & 2u) || clock64() - t0 > 6000000000LL) { sh.err |= 2u; break; }   // ~3 s: a peer is gone; fail, do not hang
                    }
                    tlo += line_value(c[0]); thi += line_value(c[1]); sm_hi += line_value(c[2]); sm_lo += line_value(c[3]);
                    e_hi += line_value(c[4]); e_lo += line_value(c[5]);
                }
                const double eps_sum_total = (double(e_hi) * 4294967296.0 + double(e_lo)) * inv_scale;
                // tlo = sum_l S_l 2^(8l) over limbs 0..3, thi over limbs 4..7 (combined per CTA before the reduction)
                const double dprime = (double(thi) * 4294967296.0 + double(tlo)) * inv_scale;    // sum_i g'_i eps_i
                const double smd = (double(sm_hi) * 4294967296.0 + double(sm_lo)) * inv_scale;
                const double *L = meta.lut[wk];
                // x.eps = L0 sum(eps) + s (S1 + 2 S2) + (L3 - L0) Smissing ; the products saw missing as 3.  The sums were
                // taken before the previous block's residual updates: its carry completes them.
                sh.d0[wk] = (L[0] * eps_sum_total + meta.sj[wk] * (dprime - 3.0 * smd) + (L[3] - L[0]) * smd) + sh.carry[wk];
                sh.carry[wk] = 0.0;
            }
            walker_sync();
            PHASE(kProfReadback);
            // (a) every marker is sampled against its dot at the start of the block
            if (act) {
                bo = meta.beta_old[wk]; uu = meta.uz[wk][0]; zz = meta.uz[wk][1]; gg = meta.grp[wk]; my_id = meta.id[wk];
                base = sh.d0[wk] + bo * NM1;
            }
            if (warp == 0 && uint32_t(32 + lane) < nb) {
                act1 = true; bo1 = meta.beta_old[32 + lane]; zz1 = meta.uz[32 + lane][1]; gg1 = meta.grp[32 + lane];
                base1 = sh.d0[32 + lane] + bo1 * NM1;
            }
            dr.beta_new = bo;
            if (act) dr = sample_marker_k(sh.gc, gg, K, base, p.sigmaE, bo, uu, zz);
            const unsigned nsm = __ballot_sync(0xffffffffu, act && !(bo == 0.0 && dr.beta_new == 0.0));
            if (lane == 0) sh.w_nsm[warp] = nsm;
            sh.w_comp[wk] = dr.comp;
            PHASE(kProfSample);
        } else {
            order_ids(bn + 1);
            if (cta == 0 && warp == kFlushWarp && sh.fin[buf].n) flush_results(sh.fin[buf]);
            if (have_next) {
                const int k2 = tid - 32 * kMetaWarp0;
                const uint32_t my_gid = gid_next;
                if (uint32_t(k2) < uint32_t(kB)) {                 // the id for the NEXT pass's gather: its latency hides behind this pass
                    const uint32_t pos = (bn + 1) * B + k2;
                    gid_next = (uint32_t(k2) < B && pos < p.n_order) ? p.order[pos] : 0u;
                }
                if (uint32_t(k2) < nb_next) gather_block_meta(p, meta_next, my_gid, k2);
                mbar_wait(&sh.mbar[buf ^ 1u], (bn >> 1) & 1u);
                PHASE_W(kProfWorkerWait, true);
                dot_partial(pk_next, nb_next, warp - kWalkWarps);
                if (prof_w) { const long long t = clock64(); sh.prof[kProfWDot] += (unsigned long long)(t - t_prev); }
                { const uint32_t wi = warp - kWalkWarps; if ((wi & 3u) < 2u) missing_sums(pk_next, nb_next, (wi >> 2) * 2 + (wi & 1u)); }
                if (prof_w) { const long long t = clock64(); sh.prof[kProfWMissing] += (unsigned long long)(t - t_prev); }
                dot_combine(warp - kWalkWarps);
                if (prof_w) { const long long t = clock64(); sh.prof[kProfWCombine] += (unsigned long long)(t - t_prev); }
                cp_async_wait_all();
                if (uint32_t(k2) < nb_next) derive_block_meta(meta_next, k2);
                PHASE_W(kProfWorkerDot, true);
            }
        }
        __syncthreads();
        PHASE(kProfJoin);

        // ---- B. the round of block b+1 goes out NOW: its latency hides behind the rest of block b -----------------
        // recycle the slot of block b-2 first (every CTA finished with it before it arrived at block b's round, and
        // nobody adds into it again before passing block b+1's round)
        if (bb == zero_turn) {                                   // block b with b >= 2 and b % nCTA == cta, without a division
            zero_turn += int(p.nCTA);
            int *old = p.slots + size_t((b - 2) % kSlots) * kSlotInts;
            for (int i = tid; i < kSlotMainLines; i += kThreads) *reinterpret_cast<unsigned long long *>(old + i * kLine) = 0ULL;
            if (MG && world > 1) {
                unsigned long long *acc = reinterpret_cast<unsigned long long *>(p.peer_xchg[p.rank] + xchg_ll_off_ints(world)) + size_t((b - 2) % kXchgRot) * kMainVals;
                for (int i = tid; i < kMainVals; i += kThreads) acc[i] = 0ULL;
            }
            const uint32_t used = sh.lz_hist[(b - 2) % kSlots];
            for (uint32_t i = tid; i < used * kCols; i += kThreads) *reinterpret_cast<int4 *>(old + kSlotLazy + size_t(i) * kLine) = make_int4(0, 0, 0, 0);
        }
        // a block that leaves the residual untouched needs its rows no more: request block b+2 into its buffer already
        const bool quiet = (sh.w_nsm[0] | sh.w_nsm[1]) == 0u;
        const bool pf_early = quiet && bn + 1 < nBlocks;
        if (pf_early) prefetch(bn + 1);
        PHASE(kProfPubPre);
        if (have_next) {
            publish_dots(bn);
            gen_wait = ++bar_gen;
        }
        PHASE(kProfPublish);

        // ---- C. the walk -----------------------------------------------------------------------------------------------
        // The one-by-one loop is reproduced exactly by iterating to a fixed point:
        //   (a) [done above] every marker is sampled against its current dot (speculating no earlier marker changes eps);
        //   (b) the non-skipped markers (src/BayesRBase.cpp:372) get their Gram rows (one grid-wide round, only for rows
        //       not yet known) and are solved in order: beta_j from its corrected dot, then the dots of all later markers
        //       are corrected by (beta_old_j - beta_j) x_j.x_k.  A triangular solve with O(1) flops per step -- no exp;
        //   (c) every marker whose dot moved is sampled again; if all components (hence all betas, bit for bit) are
        //       confirmed, the state equals the sequential chain's, else (b) restarts with the new ones.
        // Thread k < 64 samples marker k of the block; warp 0 runs the solve for both halves.
        // Asynchronous mode: no in-block correction; the rows are only needed for the next block's carry.
        uint32_t have0 = 0, have1 = 0, lz_used = 0;             // Gram rows known so far (identical in every CTA)
        ll_rows = 0;
        bool g2_ok = true;                                       // all rows so far kept their block-(b+1) half in shared memory
        // A quiet block (nothing sampled away from zero) ends here for the dot warps: they go straight to the dot pass of
        // block b+2 while the walkers file the results and wait for block b+1's round; no CTA-wide barrier in between.
        if (!quiet && have_next) mbar_wait(&sh.mbar[buf ^ 1u], (bn >> 1) & 1u);   // the walkers have not observed the rows of block b+1 yet
        for (int iter = 0; iter < 2 * kB && !quiet; ++iter) {
            const unsigned nsm0 = sh.w_nsm[0], nsm1 = sh.w_nsm[1];
            if ((nsm0 | nsm1) == 0) break;
            if (prof_on) sh.prof[kProfWalkIters] += 1;
            uint32_t need0 = nsm0 & ~have0, need1 = nsm1 & ~have1;
            if (!have_next) {                                    // the last marker of the sweep corrects nobody
                if (nb - 1 < 32) need0 &= ~(1u << (nb - 1)); else need1 &= ~(1u << (nb - 33));
                if (p.async_mode) { need0 = 0; need1 = 0; }
            }
            if (need0 | need1) {
                const uint32_t nnew = __popc(need0) + __popc(need1);
                const bool g2_was = g2_ok;
                g2_ok = g2_ok && have_next && lz_used + nnew <= 32u;
                if (MG && world > 1 && g2_was && !g2_ok && ll_rows) {
                    // the block outgrows the shared-memory row slots: its carry will read ALL rows from the global slot, where the
                    // rows sent in LL form so far hold local sums only -- merge them now (after a grid barrier: no CTA of this GPU
                    // may still be reading the local values).  Chain start only.
                    grid_arrive(p.bar);
                    grid_wait(p.bar, (++bar_gen) * p.nCTA);
                    xchg(kXLazy, slot, 0, ll_rows, tid, kThreads, cta_sync);
                    ll_rows = 0;
                }
                lazy_round(pk, pk_next, slot, meta, meta_next, nb, nb_next, need0, need1, lz_used, g2_ok);
                have0 |= need0; have1 |= need1; lz_used += nnew;
            }
            if (p.async_mode) break;
            double corr = 0.0;
            if (warp == 0) {
                const double *G = reinterpret_cast<const double *>(scratch);
                const int comp1 = sh.w_comp[32 + lane];
                double c0 = 0.0, c1 = 0.0;
                unsigned m = nsm0;
                while (m) {
                    const int jl = __ffs(m) - 1;
                    m &= m - 1;
                    const double mine = bo - beta_given_comp(sh.gc, gg, K, dr.comp, base + c0, bo, zz);
                    const double dB = __shfl_sync(0xffffffffu, mine, jl);          // eps += dB * x_j
                    if (dB != 0.0) {
                        const double *Gr = G + size_t(sh.rowslot[jl]) * kB;
                        if (act && lane > jl) c0 += dB * Gr[lane];
                        if (act1) c1 += dB * Gr[32 + lane];
                    }
                }
                m = nsm1;
                while (m) {
                    const int jl = __ffs(m) - 1;
                    m &= m - 1;
                    const double mine = bo1 - beta_given_comp(sh.gc, gg1, K, comp1, base1 + c1, bo1, zz1);
                    const double dB = __shfl_sync(0xffffffffu, mine, jl);
                    if (dB != 0.0 && act1 && lane > jl) c1 += dB * (G + size_t(sh.rowslot[32 + jl]) * kB)[32 + lane];
                }
                corr = c0;
                sh.w_corr[32 + lane] = c1;
            }
            __syncthreads();
            PHASE(kProfSolve);
            if (walker) {
                if (warp == 1) corr = sh.w_corr[wk];
                // (c) re-sample where the dot moved; confirm the components
                bool changed = false;
                if (act && corr != corr_used) {
                    const Draw nd = sample_marker_k(sh.gc, gg, K, base + corr, p.sigmaE, bo, uu, zz);
                    changed = nd.comp != dr.comp;
                    dr = nd;
                    corr_used = corr;
                }
                const unsigned chg = __ballot_sync(0xffffffffu, changed);
                const unsigned nsm = __ballot_sync(0xffffffffu, act && !(bo == 0.0 && dr.beta_new == 0.0));
                if (lane == 0) { sh.w_chg[warp] = chg; sh.w_nsm[warp] = nsm; }
                sh.w_comp[wk] = dr.comp;
            }
            __syncthreads();
            PHASE(kProfResample);
            if ((sh.w_chg[0] | sh.w_chg[1]) == 0) break;
        }
        if (walker) {
            const double dB = act ? (bo - dr.beta_new) : 0.0;
            const bool upd = act && dB != 0.0;
            const bool nonskip = act && !(bo == 0.0 && dr.beta_new == 0.0);
            BlockResult &fin = sh.fin[buf];
            if (act) {
                sh.dbeta[wk] = dB;
                fin.id[wk] = my_id; fin.grp[wk] = gg; fin.comp[wk] = dr.comp;
                fin.beta[wk] = dr.beta_new; fin.acum[wk] = dr.acum0;
            }
            if (quiet) {
                if (tid == 0) { if (real) sh.lz_hist[b % kSlots] = 0u; fin.n = nb; }
            } else {
                const unsigned um = __ballot_sync(0xffffffffu, upd);
                const unsigned nsk = __ballot_sync(0xffffffffu, nonskip);
                if (lane == 0) { sh.w_cnt[warp] = __popc(um); sh.w_cnt2[warp] = __popc(nsk); }
                walker_sync();
                const uint32_t off = warp ? sh.w_cnt[0] : 0u;
                if (upd) sh.ns_list[off + __popc(um & ((1u << lane) - 1u))] = wk;
                if (tid == 0) {
                    sh.n_ns = sh.w_cnt[0] + sh.w_cnt[1]; sh.n_nonskipped_total += sh.w_cnt2[0] + sh.w_cnt2[1];
                    if (real) sh.lz_hist[b % kSlots] = lz_used;
                    fin.n = nb;
                }
            }
        }
        if (!quiet) __syncthreads();
        PHASE(kProfFinalize);

        // ---- D. carry: the dots of block b+1 were taken before this block's updates -------------------------------
        // carry_k = sum over the updated markers j (in row-slot order) of (beta_old_j - beta_j) x_j . x_k, from the block-(b+1)
        // half of the Gram rows (still in the global slot).  512 threads = 64 columns x 8 row groups, fixed-order sums:
        // every CTA computes bit-identical values.
        const uint32_t n_ns = quiet ? 0u : sh.n_ns;
        if (n_ns && lz_used && have_next && g2_ok) {
            if (tid < kB) {                                      // fixed order over the row slots: identical in every CTA
                const double *G2 = reinterpret_cast<const double *>(scratch) + 32 * kB;
                double sacc = 0.0;
                for (uint32_t r = 0; r < lz_used; ++r) {
                    const double dB = sh.dbeta[sh.slotrow[r]];
                    if (dB != 0.0) sacc += dB * G2[r * kB + tid];
                }
                sh.carry[tid] = sacc;
            }
            PHASE(kProfCarry);
        } else if (n_ns && lz_used && have_next) {
            const uint32_t kc = tid & 63u, rg = tid >> 6;
            int4 v[kB / 8];
#pragma unroll
            for (int t = 0; t < kB / 8; ++t) {
                const uint32_t r = rg + 8 * t;
                if (r < lz_used) v[t] = __ldcg(reinterpret_cast<const int4 *>(slot + kSlotLazy + (size_t(r) * kCols + kB + kc) * kLine));
            }
            double acc = 0.0;
#pragma unroll
            for (int t = 0; t < kB / 8; ++t) {
                const uint32_t r = rg + 8 * t;
                if (r < lz_used && kc < nb_next) {
                    const uint32_t j = sh.slotrow[r];
                    const double dB = sh.dbeta[j];
                    if (dB != 0.0)
                        acc += dB * gram_entry(meta, int(j), meta_next, int(kc), double(unsigned(v[t].x)), double(unsigned(v[t].y)),
                                               double(unsigned(v[t].z)), double(unsigned(v[t].w)), Nt);
                }
            }
            double *part = reinterpret_cast<double *>(miss_list);
            part[rg * kB + kc] = acc;
            __syncthreads();
            if (tid < kB) {
                double s = 0.0;
#pragma unroll
                for (int g = 0; g < 8; ++g) s += part[g * kB + tid];
                sh.carry[tid] = s;
            }
            PHASE(kProfCarry);
        }

        // ---- E. residual update in registers, re-quantise --------------------------------------------------
        if (n_ns) {
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const uint32_t item = tid + it * kThreads;
                if (item < NI) {
                    const uint32_t w = item >> 2, c2 = 2 * (item & 3u);
                    uint32_t q = 0;
                    for (; q + 1 < n_ns; q += 2) {          // two markers per step: independent loads, same order of adds
                        const uint32_t m0 = sh.ns_list[q], m1 = sh.ns_list[q + 1];
                        const double dB0 = sh.dbeta[m0], dB1 = sh.dbeta[m1];
                        const uint32_t r0 = (pk[size_t(m0) * RS + w] >> c2) & kM3, r1 = (pk[size_t(m1) * RS + w] >> c2) & kM3;
                        const double *L0 = meta.lut[m0], *L1 = meta.lut[m1];
                        double x0[4], x1[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) { x0[t] = L0[(r0 >> (8 * t)) & 3u]; x1[t] = L1[(r1 >> (8 * t)) & 3u]; }
#pragma unroll
                        for (int t = 0; t < 4; ++t) { e[it][t] += dB0 * x0[t]; e[it][t] += dB1 * x1[t]; }   // eps += (beta_old-beta) * x
                    }
                    if (q < n_ns) {
                        const uint32_t m = sh.ns_list[q];
                        const double dB = sh.dbeta[m];
                        const uint32_t r = (pk[size_t(m) * RS + w] >> c2) & kM3;
                        const double *L = meta.lut[m];
#pragma unroll
                        for (int t = 0; t < 4; ++t) e[it][t] += dB * L[(r >> (8 * t)) & 3u];
                    }
                }
            }
            PHASE(kProfAxpy);
            // individuals past n_local carry eps = 0 and code 0, so their lut[0] increments must not stick
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const uint32_t item = tid + it * kThreads;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint32_t gi = slice0 + 16 * (item >> 2) + 4 * t + (item & 3u);
                    if (gi >= p.n_local) e[it][t] = 0.0;
                }
            }
            quantize();
            if (prof_on) sh.prof[kProfBlocksUpdated] += 1;
            __syncthreads();
            PHASE(kProfQuantize);
        }

        // ---- F. block b's buffer is free: prefetch block b+2 into it ----------------------------------------------
        if (!pf_early && bn + 1 < nBlocks) prefetch(bn + 1);
        if (!quiet && bn + 1 < nBlocks && p.n_order - (bn + 1) * B < B) __syncthreads();   // zero rows of a partial last block
        PHASE(kProfTail);
        PHASE_W(0, false);
    }
    __syncthreads();
    if (cta == 0 && warp == kFlushWarp) {                       // the last two blocks, in block order
        const uint32_t older = nBlocks & 1u;                    // parity of block nBlocks-2
        if (sh.fin[older].n) flush_results(sh.fin[older]);
        if (sh.fin[older ^ 1u].n) flush_results(sh.fin[older ^ 1u]);
    }

    // ---- epilogue: eps back to HBM, sweep scalars ------------------------------------------------
    double sq = 0.0;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const uint32_t item = tid + it * kThreads;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint32_t gi = slice0 + 16 * (item >> 2) + 4 * t + (item & 3u);
            if (item < NI && gi < p.n_local) { p.eps[gi] = e[it][t]; sq += e[it][t] * e[it][t]; }
        }
    }
    sq = warp_sum(sq);
    if (lane == 0) sh.warp_tmp[warp] = sq;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += sh.warp_tmp[w];
        // this slice's sum(eps) from the exact fixed-point sum
        unsigned long long qhi, qlo;
        slice_q(qhi, qlo);
        p.cta_red[2 * cta + 0] = (double((long long)qhi) * 4294967296.0 + double((long long)qlo)) * inv_scale;
        p.cta_red[2 * cta + 1] = s;
        if (sh.err) atomicOr(p.out_counts + 2, sh.err);
    }
    grid_arrive(p.bar);
    if (cta == 0) {
        grid_wait(p.bar, (++bar_gen) * p.nCTA);
        if (warp == 0) {
            double s = 0.0, q = 0.0;
            for (uint32_t c = lane; c < p.nCTA; c += 32) { s += ld_cg(p.cta_red + 2 * c); q += ld_cg(p.cta_red + 2 * c + 1); }
            s = warp_sum(s); q = warp_sum(q);
            if (lane == 0) {
                p.out_scalars[0] = s; p.out_scalars[1] = q;
                p.out_counts[0] = sh.n_nonskipped_total; p.out_counts[1] = nBlocks; p.out_counts[3] = bar_gen;
                if (world > 1) *p.round_ctr = round_base + xround;
            }
        }
        __syncthreads();
        if (p.prof && tid == 0) for (int i = 0; i < kProfPhases; ++i) p.prof[i] += sh.prof[i];
        for (uint32_t i = tid; i < p.nG * K; i += kThreads) p.v[i] = double(sh.v_s[i]);
        for (uint32_t i = tid; i < p.nG; i += kThreads) p.betasq[i] = sh.betasq_s[i];
    }
}

} // namespace brr
