// Is the 4-bit integer MMA usable on B200?  mma.sync m16n8k64 u4 x s4 vs m16n8k32 u8 x s8: latency and per-SM rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/latency_bench3 tools/latency_bench3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)
__device__ __forceinline__ void mma_u8s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_u4s4(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k64.row.col.s32.u4.s4.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int MODE> __global__ void k(long long *out, int *sink)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int c[4][4] = {};
    const uint32_t a = 0x01020301u * (lane + 1), b = 0x11111111u;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (warp == 0) {
        t0 = clock64();
#pragma unroll 8
        for (int i = 0; i < 128; ++i) { if (MODE) mma_u4s4(c[0], a, a, a, a, b, b); else mma_u8s8(c[0], a, a, a, a, b, b); }
        t1 = clock64();
    }
    __syncthreads();
    t2 = clock64();
#pragma unroll 2
    for (int i = 0; i < 128; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) { if (MODE) mma_u4s4(c[q], a, a, a, a, b, b); else mma_u8s8(c[q], a, a, a, a, b, b); }
    __syncthreads();
    const long long t3 = clock64();
    if (threadIdx.x == 0) { out[0] = (t1 - t0) / 128; out[1] = t3 - t2; }
    if (c[0][0] + c[1][1] + c[2][2] + c[3][3] == 123456789) sink[0] = 1;
    // correctness probe: lane 0's c[0][0] after ONE fresh MMA with known operands
    int d[4] = {0, 0, 0, 0};
    if (MODE) mma_u4s4(d, 0x33333333u, 0x33333333u, 0x33333333u, 0x33333333u, 0x77777777u, 0x77777777u);
    else mma_u8s8(d, 0x03030303u, 0x03030303u, 0x03030303u, 0x03030303u, 0x07070707u, 0x07070707u);
    if (threadIdx.x == 0) out[2] = d[0];
}
int main()
{
    long long *out; int *sink; CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&sink, 64));
    for (int mode = 0; mode < 2; ++mode) {
        if (mode) k<1><<<148, 512>>>(out, sink); else k<0><<<148, 512>>>(out, sink);
        CK(cudaDeviceSynchronize());
        long long h[3]; CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
        printf("%s: dependent chain %lld cycles/MMA; 16 warps x 4 chains x 128: %lld cycles -> %.2f cycles/MMA/SM; probe c00 = %lld (expect %d)\n",
               mode ? "m16n8k64 u4 x s4" : "m16n8k32 u8 x s8", h[0], h[1], double(h[1]) / (16 * 512), h[2], mode ? 64 * 3 * 7 : 32 * 3 * 7);
    }
    return 0;
}
