// Microbenchmark of the dot pass of the dense sweep kernels on shared-memory resident data (one CTA on one SM): cycles per 64-marker
// block of a 3392-individual slice for several ways of splitting the block over warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I bayesrrcmd_b200/csrc -o tools/dot_bench tools/dot_bench.cu
#include <cstdio>
#include <cstdlib>
#include "sweep_dense.cuh"
using namespace brr;

constexpr uint32_t SL = 3392, PW = SL / 16, RS = PW + 8, NP = SL / 64, NF = NP / 2;

// MODE 0: full loop; 1: no LOP3 (raw words into the MMAs); 2: no MMAs (LOP3 results xor-ed into the accumulators); 3: loads only
template <int TILES, int KPARTS, int MODE>
__global__ void __launch_bounds__(512, 1) dot_kernel(int reps, long long *out, int *sink)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *limbs = reinterpret_cast<uint4 *>(smem_raw);                       // [8][RS]
    uint32_t *codes = reinterpret_cast<uint32_t *>(smem_raw + size_t(8) * RS * 16);   // [64][RS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < 8 * RS * 4; i += blockDim.x) reinterpret_cast<uint32_t *>(limbs)[i] = i * 2654435761u;
    for (uint32_t i = tid; i < 64 * RS; i += blockDim.x) codes[i] = i * 40503u + 17u;
    __syncthreads();
    const uint32_t gid = lane >> 2, tig = lane & 3;
    constexpr int GROUPS = 4 / TILES;                 // tile groups of a block
    const uint32_t tg = warp / KPARTS, kp = warp % KPARTS;
    const bool active = warp < GROUPS * KPARTS;
    int acc[TILES][4][4];
#pragma unroll
    for (int tl = 0; tl < TILES; ++tl)
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[tl][c][q] = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (active)
    for (int rep = 0; rep < reps; ++rep) {
        const uint32_t *pk = codes + size_t(tg) * TILES * 16 * RS;
        const uint32_t lo = gid * RS + tig;
        uint32_t aR = smem_u32(pk + lo + 8 * kp);
        uint32_t aL = smem_u32(limbs + lo + 8 * kp);
        uint32_t cnt = NF > kp ? (NF - kp + KPARTS - 1) / KPARTS : 0u;
        uint4 LA = lds_v4<0>(aL), LB = lds_v4<64>(aL);
        uint32_t x[TILES][4];
#pragma unroll
        for (int tl = 0; tl < TILES; ++tl) {
            x[tl][0] = lds_u32<0>(aR + tl * 64u * RS); x[tl][1] = lds_u32<0>(aR + tl * 64u * RS + 32u * RS);
            x[tl][2] = lds_u32<16>(aR + tl * 64u * RS); x[tl][3] = lds_u32<16>(aR + tl * 64u * RS + 32u * RS);
        }
#pragma unroll 1
        for (;;) {
            const bool more = --cnt != 0u;
            if (more) { aR += 32u * KPARTS; aL += 128u * KPARTS; }
            const uint4 LAn = lds_v4<0>(aL), LBn = lds_v4<64>(aL);
            uint32_t xn[TILES][4];
#pragma unroll
            for (int tl = 0; tl < TILES; ++tl) {
                xn[tl][0] = lds_u32<0>(aR + tl * 64u * RS); xn[tl][1] = lds_u32<0>(aR + tl * 64u * RS + 32u * RS);
                xn[tl][2] = lds_u32<16>(aR + tl * 64u * RS); xn[tl][3] = lds_u32<16>(aR + tl * 64u * RS + 32u * RS);
            }
            if (MODE == 0) {
#pragma unroll
                for (int tl = 0; tl < TILES; ++tl) {
                    mma_u8s8(acc[tl][0], x[tl][0] & kC0, x[tl][1] & kC0, x[tl][2] & kC0, x[tl][3] & kC0, LA.x, LB.x);
                    mma_u8s8(acc[tl][1], x[tl][0] & kC1, x[tl][1] & kC1, x[tl][2] & kC1, x[tl][3] & kC1, LA.y, LB.y);
                    mma_u8s8(acc[tl][2], x[tl][0] & kC2, x[tl][1] & kC2, x[tl][2] & kC2, x[tl][3] & kC2, LA.z, LB.z);
                    mma_u8s8(acc[tl][3], x[tl][0] & kC3, x[tl][1] & kC3, x[tl][2] & kC3, x[tl][3] & kC3, LA.w, LB.w);
                }
            } else if (MODE == 1) {
#pragma unroll
                for (int tl = 0; tl < TILES; ++tl) {
                    mma_u8s8(acc[tl][0], x[tl][0], x[tl][1], x[tl][2], x[tl][3], LA.x, LB.x);
                    mma_u8s8(acc[tl][1], x[tl][0], x[tl][1], x[tl][2], x[tl][3], LA.y, LB.y);
                    mma_u8s8(acc[tl][2], x[tl][0], x[tl][1], x[tl][2], x[tl][3], LA.z, LB.z);
                    mma_u8s8(acc[tl][3], x[tl][0], x[tl][1], x[tl][2], x[tl][3], LA.w, LB.w);
                }
            } else if (MODE == 2) {
#pragma unroll
                for (int tl = 0; tl < TILES; ++tl) {
                    acc[tl][0][0] ^= (x[tl][0] & kC0) + LA.x; acc[tl][0][1] ^= (x[tl][1] & kC0) + LB.x; acc[tl][0][2] ^= x[tl][2] & kC0; acc[tl][0][3] ^= x[tl][3] & kC0;
                    acc[tl][1][0] ^= (x[tl][0] & kC1) + LA.y; acc[tl][1][1] ^= (x[tl][1] & kC1) + LB.y; acc[tl][1][2] ^= x[tl][2] & kC1; acc[tl][1][3] ^= x[tl][3] & kC1;
                    acc[tl][2][0] ^= (x[tl][0] & kC2) + LA.z; acc[tl][2][1] ^= (x[tl][1] & kC2) + LB.z; acc[tl][2][2] ^= x[tl][2] & kC2; acc[tl][2][3] ^= x[tl][3] & kC2;
                    acc[tl][3][0] ^= (x[tl][0] & kC3) + LA.w; acc[tl][3][1] ^= (x[tl][1] & kC3) + LB.w; acc[tl][3][2] ^= x[tl][2] & kC3; acc[tl][3][3] ^= x[tl][3] & kC3;
                }
            } else {
#pragma unroll
                for (int tl = 0; tl < TILES; ++tl) { acc[tl][0][0] ^= x[tl][0] ^ x[tl][1] ^ x[tl][2] ^ x[tl][3] ^ LA.x ^ LB.y ^ LA.z ^ LB.w; }
            }
            if (!more) break;
            LA = LAn; LB = LBn;
#pragma unroll
            for (int tl = 0; tl < TILES; ++tl)
#pragma unroll
                for (int q = 0; q < 4; ++q) x[tl][q] = xn[tl][q];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int tl = 0; tl < TILES; ++tl)
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) s ^= acc[tl][c][q];
    if (s == 0x12345678) sink[tid] = s;
    if (tid == 0) out[0] = t1 - t0;
}

template <int TILES, int KPARTS, int MODE>
static void run(const char *what)
{
    long long *out; int *sink;
    cudaMalloc(&out, 8); cudaMalloc(&sink, 4096);
    const size_t smem = size_t(8) * RS * 16 + size_t(64) * RS * 4;
    cudaFuncSetAttribute(dot_kernel<TILES, KPARTS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int reps = 200;
    dot_kernel<TILES, KPARTS, MODE><<<1, 512, smem>>>(reps, out, sink);
    dot_kernel<TILES, KPARTS, MODE><<<1, 512, smem>>>(reps, out, sink);
    long long h = 0;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-72s %2d warps  %7.0f cycles per 64-marker block%s\n", what, (4 / TILES) * KPARTS, double(h) / reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out); cudaFree(sink);
}

int main()
{
    run<2, 4, 0>("v6: 2 tiles per warp, 4 k-parts, full loop");
    run<2, 4, 1>("v6: 2 tiles per warp, 4 k-parts, no LOP3");
    run<2, 4, 2>("v6: 2 tiles per warp, 4 k-parts, no MMA");
    run<2, 4, 3>("v6: 2 tiles per warp, 4 k-parts, loads only");
    run<2, 8, 0>("2 tiles per warp, 8 k-parts, full loop");
    run<2, 2, 0>("2 tiles per warp, 2 k-parts, full loop");
    run<1, 4, 0>("1 tile per warp, 4 k-parts, full loop");
    run<1, 3, 0>("v5: 1 tile per warp, 3 k-parts, full loop");
    run<1, 2, 0>("1 tile per warp, 2 k-parts, full loop");
    run<1, 4, 1>("1 tile per warp, 4 k-parts, no LOP3");
    run<1, 4, 2>("1 tile per warp, 4 k-parts, no MMA");
    run<4, 4, 0>("4 tiles per warp, 4 k-parts, full loop");
    run<4, 2, 0>("4 tiles per warp, 2 k-parts, full loop");
    run<4, 8, 0>("4 tiles per warp, 8 k-parts, full loop");
    run<4, 16, 0>("4 tiles per warp, 16 k-parts, full loop");
    return 0;
}
