// Device helpers shared by the sweep kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace brr {

constexpr int kThreads = 512;          // 16 warps per CTA, one CTA per SM
constexpr int kWarps = kThreads / 32;
constexpr int kMaxB = 64;              // dense: markers per block (two half-blocks of 32 sampling lanes)
constexpr int kMaxW = 32;              // ragged: markers per window (lanes of the sampling warp)
constexpr int kMaxK = 16;              // mixture components
constexpr int kMaxGK = 256;            // n_groups * K held in shared memory

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- deterministic reductions ------------------------------------------------------------------
// xor-butterfly: every lane ends with the bit-identical sum (a+b == b+a in IEEE-754).
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// acc += e if (q & mask) != 0, as ONE predicated DADD (the compiler otherwise emits DADD + 2 FSEL)
__device__ __forceinline__ void add_if_bit(double &acc, double e, uint32_t q, uint32_t mask)
{
    asm("{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 t;\n\t"
        "and.b32 t, %2, %3;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "@p add.f64 %0, %0, %1;\n\t"
        "}"
        : "+d"(acc)
        : "d"(e), "r"(q), "r"(mask));
}

// ---- grid-wide barrier: monotonically increasing arrival counter --------------------------------
__device__ __forceinline__ void red_release_add(unsigned *p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// system-scope variants for flags living in (or written into) peer memory over NVLink
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All threads of the CTA call both halves. Arrive publishes every earlier global write of the CTA.
__device__ __forceinline__ void grid_arrive(unsigned *bar)
{
    __syncthreads();
    if (threadIdx.x == 0) red_release_add(bar, 1u);
}
__device__ __forceinline__ void grid_wait(const unsigned *bar, unsigned target)
{
    if (threadIdx.x == 0) {
        while (ld_acquire(bar) < target) { }
    }
    __syncthreads();
}

// L2-only loads for data other CTAs (or GPUs) produced during this launch.
__device__ __forceinline__ double ld_cg(const double *p) { return __ldcg(p); }
__device__ __forceinline__ int ld_cg(const int *p) { return __ldcg(p); }

// ---- mbarrier + TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) -----------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared, 16-byte aligned on both sides, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

} // namespace brr
