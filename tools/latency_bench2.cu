// More single-SM latencies behind the dense sweep's design (B200, sm_100a): legacy IMMA chains, TMA bulk-copy issue cost for
// 64 x 880-byte rows vs one 56 KB copy vs cp.async, and the cost of a release after a burst of REDs.  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/latency_bench2 tools/latency_bench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mma_u8s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int ROWB = 880, NROWS = 64;

__global__ void k(const unsigned char *src, size_t row_pitch, int *redbuf, long long *out, int mode)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    long long t0, t1, t2;
    int kx = 0;
    // ---- IMMA: one chain, then four chains (one warp) ----
    if (warp == 0) {
        int c[4][4] = {};
        const uint32_t a = 0x01020301u * (lane + 1), b = 0x01010101u;
        t0 = clock64();
#pragma unroll 8
        for (int i = 0; i < 128; ++i) mma_u8s8(c[0], a, a, a, a, b, b);
        t1 = clock64();
#pragma unroll 2
        for (int i = 0; i < 128; ++i) { mma_u8s8(c[0], a, a, a, a, b, b); mma_u8s8(c[1], a, a, a, a, b, b); mma_u8s8(c[2], a, a, a, a, b, b); mma_u8s8(c[3], a, a, a, a, b, b); }
        t2 = clock64();
        if (lane == 0) { out[0] = (t1 - t0) / 128; out[1] = (t2 - t1) / 128; }
        if (c[0][0] + c[1][1] + c[2][2] + c[3][3] == 123456789) out[63] = 1;
    }
    __syncthreads();
    // ---- all 16 warps: 4 chains each: IMMA throughput per SM ----
    {
        int c[4][4] = {};
        const uint32_t a = 0x01020301u * (lane + 1), b = 0x01010101u;
        __syncthreads();
        t0 = clock64();
#pragma unroll 2
        for (int i = 0; i < 128; ++i) { mma_u8s8(c[0], a, a, a, a, b, b); mma_u8s8(c[1], a, a, a, a, b, b); mma_u8s8(c[2], a, a, a, a, b, b); mma_u8s8(c[3], a, a, a, a, b, b); }
        __syncthreads();
        t1 = clock64();
        if (tid == 0) out[2] = (t1 - t0);      // cycles for 16 warps x 512 MMAs
        if (c[0][0] + c[1][1] + c[2][2] + c[3][3] == 123456789) out[63] = 1;
    }
    __syncthreads();
    // ---- TMA: 64 rows of 880 B (two per lane of warp 15), issue time and completion time ----
    for (int rep = 0; rep < 3; ++rep) {
        __syncthreads();
        if (warp == 15) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            t0 = clock64();
            if (lane == 0) mbar_expect(&bar, NROWS * ROWB);
            __syncwarp();
            for (int h = 0; h < 2; ++h) {
                const int r = lane + 32 * h;
                bulk_g2s(sm + r * 896, src + size_t((r * 7919 + rep * 131) % 4096) * row_pitch, ROWB, &bar);
            }
            t1 = clock64();
            mbar_wait(&bar, rep & 1);
            t2 = clock64();
            if (lane == 0) { out[4 + 2 * rep] = t1 - t0; out[5 + 2 * rep] = t2 - t0; }
        }
    }
    __syncthreads();
    // ---- TMA: the same rows, issued by 16 warps (4 rows each, lanes 0..3) ----
    {
        __syncthreads();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        t0 = clock64();
        if (tid == 0) mbar_expect(&bar, NROWS * ROWB);
        __syncthreads();
        if (lane < 4) { const int r = warp * 4 + lane; bulk_g2s(sm + r * 896, src + size_t((r * 7919 + 977) % 4096) * row_pitch, ROWB, &bar); }
        t1 = clock64();
        mbar_wait(&bar, 1);
        t2 = clock64();
        if (tid == 0) { out[10] = t1 - t0; out[11] = t2 - t0; }
    }
    __syncthreads();
    // ---- TMA: ONE contiguous 56320-byte copy ----
    if (warp == 15) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        t0 = clock64();
        if (lane == 0) { mbar_expect(&bar, NROWS * ROWB); bulk_g2s(sm, src + 5000 * row_pitch, NROWS * ROWB, &bar); }
        t1 = clock64();
        mbar_wait(&bar, 0);
        t2 = clock64();
        if (lane == 0) { out[12] = t1 - t0; out[13] = t2 - t0; }
    }
    __syncthreads();
    // ---- cp.async 16 B: the same 64 rows by 13 warps (rows w, w+13, ..; 55 chunks per row) ----
    {
        __syncthreads();
        t0 = clock64();
        if (warp >= 2 && warp < 15) {
            for (int r = warp - 2; r < NROWS; r += 13) {
                const unsigned char *g = src + size_t((r * 7919 + 4242) % 4096 + 8000) * row_pitch;
                for (int cch = lane; cch < 55; cch += 32)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sm + r * 896 + cch * 16)), "l"(g + cch * 16) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            t1 = clock64();
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else t1 = clock64();
        __syncthreads();
        t2 = clock64();
        if (tid == 64) { out[14] = t1 - t0; out[15] = t2 - t0; }
    }
    __syncthreads();
    // ---- burst of REDs (one per thread, own 128-byte line) followed by red.release.gpu by thread 0 ----
    for (int rep = 0; rep < 2; ++rep) {
        __syncthreads();
        t0 = clock64();
        if (mode) atomicAdd(reinterpret_cast<unsigned long long *>(redbuf + (tid + 512 * rep) * 32), 1ULL);
        __syncthreads();
        t1 = clock64();
        if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(redbuf + 40000), "r"(1u) : "memory");
        t2 = clock64();
        if (tid == 0) { out[16 + 2 * rep] = t1 - t0; out[17 + 2 * rep] = t2 - t1; }
    }
    // ---- red.release alone, then a relaxed RED + __threadfence ----
    __syncthreads();
    t0 = clock64();
    if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(redbuf + 40000), "r"(1u) : "memory");
    t1 = clock64();
    if (tid == 0) { out[20] = t1 - t0; }
    if (sm[tid] == 77 && sm[tid + 30000] == 99) out[62] = 1;
}

int main()
{
    const size_t pitch = 880 * 148;           // rows of one CTA slice of consecutive markers are 148 rows apart
    unsigned char *src; CK(cudaMalloc(&src, pitch * 16384)); CK(cudaMemset(src, 1, pitch * 16384));
    int *red; CK(cudaMalloc(&red, 4 * 41000 * 32)); CK(cudaMemset(red, 0, 4 * 41000));
    long long *out; CK(cudaMalloc(&out, 64 * 8)); CK(cudaMemset(out, 0, 64 * 8));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 896));
    for (int grid : {1, 148}) {
        k<<<grid, 512, 64 * 896>>>(src, pitch, red, out, 1);
        CK(cudaDeviceSynchronize());
        long long h[64]; CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
        printf("-- grid %d (numbers of the last CTA to write)\n", grid);
        printf("IMMA m16n8k32 dependent chain: %lld cycles/MMA; 4 independent chains: %lld cycles per 4 MMAs (one warp)\n", h[0], h[1]);
        printf("IMMA 16 warps x 4 chains x 128: %lld cycles -> %.2f cycles/MMA/SM\n", h[2], double(h[2]) / (16 * 512));
        for (int r = 0; r < 3; ++r) printf("TMA 64 x 880 B by one warp, rep %d (%s): issue %lld, complete %lld cycles\n", r, r ? "other rows" : "cold", h[4 + 2 * r], h[5 + 2 * r]);
        printf("TMA 64 x 880 B by 16 warps x 4 lanes: issue %lld, complete %lld cycles\n", h[10], h[11]);
        printf("TMA 1 x 56320 B: issue %lld, complete %lld cycles\n", h[12], h[13]);
        printf("cp.async.cg 16 B, 64 x 880 B by 13 warps: issue %lld, complete (+CTA sync) %lld cycles\n", h[14], h[15]);
        printf("512 REDs (own lines) + syncthreads: %lld / %lld cycles; red.release.gpu after them: %lld / %lld cycles; red.release alone: %lld\n", h[16], h[18], h[17], h[19], h[20]);
    }
    return 0;
}
