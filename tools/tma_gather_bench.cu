// How fast can 148 CTAs gather random 880-byte rows (one per (marker, CTA)) from HBM with TMA bulk copies, as the dense sweep does?
// Pure data movement: each CTA keeps `depth` groups of 16 rows in flight into a shared-memory ring and does nothing else.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_gather_bench tools/tma_gather_bench.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
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
// group = 16 rows; `depth` groups in flight; warps 0..W-1 issue (one group each, round robin), everybody else idle
template <int ROWB>
__global__ void __launch_bounds__(512, 1) gather(const unsigned char *codes, const uint32_t *order, uint32_t n_order, uint32_t nCTA, int depth, int issue_warps,
                                                 size_t marker_stride, size_t cta_stride, unsigned long long *sink)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t bars[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1);
    __syncthreads();
    const uint32_t n_groups = n_order / 16;
    if (warp < issue_warps) {
        unsigned long long acc = 0;
        // warp w handles groups w, w + issue_warps, ...; ring slot = group % depth; slot reuse waits for the slot's previous group
        for (uint32_t g = warp; g < n_groups; g += issue_warps) {
            const int slot = g % depth;
            const uint32_t use = g / depth;
            if (use > 0) { mbar_wait(&bars[slot], (use - 1) & 1u); acc += sm[size_t(slot) * 16 * ROWB + lane * 8]; }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (lane == 0) mbar_expect(&bars[slot], 16 * ROWB);
            __syncwarp();
            if (lane < 16) {
                const uint32_t id = order[g * 16 + lane];
                bulk_g2s(sm + (size_t(slot) * 16 + lane) * ROWB, codes + size_t(id) * marker_stride + size_t(blockIdx.x) * cta_stride, ROWB, &bars[slot]);
            }
        }
        if (lane == 0) atomicAdd(sink, acc);
    }
    // NOTE: a slot is owned by the warp (group % depth) only if depth is a multiple of issue_warps; the host picks such pairs
}
int main(int argc, char **argv)
{
    const uint32_t nCTA = 148, M = argc > 1 ? atoi(argv[1]) : 400000;
    constexpr int ROWB = 880;
    const size_t bytes = size_t(M) * nCTA * ROWB;
    unsigned char *codes; CK(cudaMalloc(&codes, bytes)); CK(cudaMemset(codes, 1, bytes));
    std::vector<uint32_t> order(M); for (uint32_t i = 0; i < M; ++i) order[i] = i;
    std::mt19937 rng(1); std::shuffle(order.begin(), order.end(), rng);
    uint32_t *dorder; CK(cudaMalloc(&dorder, M * 4)); CK(cudaMemcpy(dorder, order.data(), M * 4, cudaMemcpyHostToDevice));
    std::vector<uint32_t> seq(M); for (uint32_t i = 0; i < M; ++i) seq[i] = i;
    uint32_t *dseq; CK(cudaMalloc(&dseq, M * 4)); CK(cudaMemcpy(dseq, seq.data(), M * 4, cudaMemcpyHostToDevice));
    unsigned long long *sink; CK(cudaMalloc(&sink, 8));
    CK(cudaFuncSetAttribute(gather<ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 16 * ROWB));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("%u markers x %u CTAs x %d B = %.1f GB\n", M, nCTA, ROWB, bytes / 1e9);
    for (int layout = 0; layout < 2; ++layout)             // 0: [marker][cta][row] (the store's), 1: [cta][marker][row]
        for (int random = 1; random >= 0; --random)
            for (int w : {1, 2, 4, 8})
                for (int depth : {w, 2 * w, 4 * w}) {
                    if (depth > 16 || depth < 2) continue;
                    const size_t ms = layout == 0 ? size_t(nCTA) * ROWB : size_t(ROWB), cs = layout == 0 ? size_t(ROWB) : size_t(M) * ROWB;
                    for (int rep = 0; rep < 2; ++rep) {
                        cudaEventRecord(e0);
                        gather<ROWB><<<nCTA, 512, 16 * 16 * ROWB>>>(codes, random ? dorder : dseq, M, nCTA, depth, w, ms, cs, sink);
                        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                    }
                    float t; cudaEventElapsedTime(&t, e0, e1);
                    printf("layout %d %s order, %d issuing warps, %2d groups (%3d KB) in flight per SM: %7.2f ms  %7.1f GB/s  %.2f us per 64-marker block\n", layout, random ? "random" : "sequential",
                           w, depth, depth * 16 * ROWB / 1024, t, bytes / t / 1e6, t * 1e3 / (M / 64));
                }
    return 0;
}
