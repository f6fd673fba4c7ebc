// Dependent-chain latencies of the operations on the walkers' critical path (B200, sm_100a): one warp, clock64().
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/latency_bench tools/latency_bench.cu
// Not part of the product.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int N = 256;

__global__ void lat_kernel(double *io, int *gi, long long *out, int nwarps_active)
{
    __shared__ double sm[1024];
    __shared__ int smi[1024];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) { sm[i] = 1.0 + 1e-9 * i; smi[i] = (i * 37 + 11) & 1023; }
    __syncthreads();
    if ((threadIdx.x >> 5) >= nwarps_active) return;
    double x = io[lane], y = io[32 + lane];
    long long t0, t1;
    int k = 0;
    // DFMA dependent chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
    t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
    // DADD dependent chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x + y;
    t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
    // DMUL
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x * y;
    t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
    x = io[lane] * 1e-3;
    // exp()
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = exp(x) - 1.0;
    t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
    // 1/x
    x = io[lane] + 1.5;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = 1.0 / x + 1.0;
    t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
    // three independent exp + sum + reciprocal (the sampler's shape)
    t0 = clock64();
#pragma unroll 2
    for (int i = 0; i < N; ++i) { const double a = exp(x * 0.1), b = exp(x * 0.2), c = exp(x * 0.3); x = 1.0 / (1.0 + a + b + c); }
    t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
    // FFMA chain
    float f = float(x), g = float(y);
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) f = fmaf(f, g, 1e-9f);
    t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
    // ex2.approx.f32 chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) f = exp2f(f) - 1.0f;
    t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
    // LDS pointer chase
    int p = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) p = smi[p];
    t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
    // LDS.64 dependent (address from value)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { const double v = sm[p]; p = (p + int(v)) & 1023; }
    t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
    // ld.global.cg pointer chase (L2)
    int q = gi[lane];
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) q = __ldcg(gi + q);
    t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
    // redux.sync add
    int r = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) r = __reduce_add_sync(0xffffffffu, r) & 7;
    t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
    // shfl.sync (double = 2 shuffles)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
    t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
    // double -> int64 -> double (quantise round trip)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = double(__double2ll_rn(x * 1.5));
    t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
    // bar.sync of the whole CTA
    if (nwarps_active == int(blockDim.x >> 5)) {
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < N; ++i) __syncthreads();
        t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0;
    }
    ++k;
    io[64 + lane] = x + f + p + q + r;
}

int main()
{
    double *io; int *gi; long long *out;
    CK(cudaMalloc(&io, 4096)); CK(cudaMalloc(&gi, 1 << 22)); CK(cudaMalloc(&out, 256));
    double h[96]; for (int i = 0; i < 96; ++i) h[i] = 1.0 + 1e-6 * i;
    CK(cudaMemcpy(io, h, sizeof(h), cudaMemcpyHostToDevice));
    const int n = 1 << 20;
    int *hg = new int[n];
    for (int i = 0; i < n; ++i) hg[i] = int((uint64_t(i) * 2654435761ull + 12345) % n);
    CK(cudaMemcpy(gi, hg, n * 4, cudaMemcpyHostToDevice));
    const char *names[] = {"DFMA", "DADD", "DMUL", "exp(double)", "1.0/x (double)", "3 x exp + sum + 1/s", "FFMA", "ex2.approx.f32",
                           "LDS.32 chase", "LDS.64 + IADD chase", "ld.global.cg chase (L2)", "redux.sync.add", "shfl.sync f64", "f64->s64->f64", "bar.sync (512 thr)"};
    for (int warps : {1, 2, 16}) {
        CK(cudaMemset(out, 0, 256));
        lat_kernel<<<1, 512>>>(io, gi, out, warps);
        CK(cudaDeviceSynchronize());
        long long ho[16];
        CK(cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost));
        printf("-- %d active warp(s) of a 512-thread CTA: cycles per dependent step\n", warps);
        for (int k = 0; k < 15; ++k) if (ho[k]) printf("   %-28s %8.1f\n", names[k], double(ho[k]) / N);
    }
    return 0;
}
