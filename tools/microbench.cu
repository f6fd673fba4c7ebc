// Microbenchmarks behind the design decisions in DESIGN.md (B200, sm_100a).  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench tools/microbench.cu
// Each test prints one line: name, per-SM per-clock rate (using the measured SM clock) or latency.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static int g_sms = 148;
static double g_clk_hz = 1.9e9;

template <class F> static float time_ms(F f, int reps = 3)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

__global__ void clk_kernel(long long *out, int iters)
{
    long long t0 = clock64();
    float x = threadIdx.x;
    for (int i = 0; i < iters; ++i) x = x * 1.0001f + 0.5f;
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)x; }
}

// ---- A. DADD -------------------------------------------------------------------------------------
__global__ void dadd_kernel(double *out, int iters, unsigned maskbits)
{
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
    double e = out[threadIdx.x & 7];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += e;
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;
}
__global__ void dadd_pred_kernel(double *out, int iters, unsigned q)
{
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
    double e = out[threadIdx.x & 7];
    q ^= threadIdx.x * 2654435761u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 p, t, 0;\n\t@p add.f64 %0, %0, %1;\n\t}"
                : "+d"(a[i]) : "d"(e), "r"(q), "r"(1u << (i + (it & 15))));
        }
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;
}
__global__ void dfma_kernel(double *out, int iters)
{
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
    double e = out[threadIdx.x & 7], f = out[(threadIdx.x + 1) & 7];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], f, e);
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;
}

// ---- B. POPC / LOP3 / DP4A -----------------------------------------------------------------------
__global__ void popc_kernel(unsigned *out, int iters)
{
    unsigned a[8], acc[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 2654435761u + i; acc[i] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i] += __popc(a[i] ^ acc[i]); }
    }
    unsigned s = 0; for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 0x12345678u) out[0] = s;
}
__global__ void dp4a_kernel(int *out, int iters)
{
    int a[8], acc[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 2654435761u + i; acc[i] = 0; }
    int b = out[threadIdx.x & 7];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = __dp4a(a[i], b, acc[i]);
    }
    int s = 0; for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 0x12345678) out[0] = s;
}
__global__ void lop3_kernel(unsigned *out, int iters)
{
    unsigned a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 2654435761u + i;
    unsigned b = out[threadIdx.x & 7], c = out[(threadIdx.x + 3) & 7];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = (a[i] & b) ^ (c | a[(i + 1) & 7]);
    }
    unsigned s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 0x12345678u) out[0] = s;
}

// ---- C. legacy tensor-core paths: IMMA s8 m16n8k32, BMMA b1 m16n8k256 ------------------------------
__global__ void imma_kernel(int *out, int iters)
{
    int c[4][4];
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = threadIdx.x * 11, b1 = threadIdx.x * 13;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int s = 0; for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    if (s == 0x12345678) out[0] = s;
}
#ifdef WITH_BMMA
__global__ void bmma_kernel(int *out, int iters)
{
    int c[4][4];
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = threadIdx.x * 11, b1 = threadIdx.x * 13;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int s = 0; for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    if (s == 0x12345678) out[0] = s;
}
#endif

// ---- D. grid barrier round trip --------------------------------------------------------------------
__global__ void gridbar_kernel(unsigned *bar, int iters, long long *cycles)
{
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
            unsigned target = (it + 1) * gridDim.x, v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
// barrier + every CTA publishes 32 doubles and reads all nCTA x 32 partials back (what a block exchange does)
__global__ void gridbar_xchg_kernel(unsigned *bar, double *part, int iters, long long *cycles, double *sink)
{
    long long t0 = clock64();
    double acc = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int it = 0; it < iters; ++it) {
        double *p = part + size_t(it & 1) * gridDim.x * 32;
        if (threadIdx.x < 32) p[blockIdx.x * 32 + threadIdx.x] = acc + threadIdx.x;
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
            unsigned target = (it + 1) * gridDim.x, v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
        }
        __syncthreads();
        double v[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) { int c = warp + i * nw; v[i] = c < gridDim.x ? __ldcg(p + c * 32 + lane) : 0.0; }
#pragma unroll
        for (int i = 0; i < 12; ++i) acc += v[i];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    if (acc == 1.2345) sink[0] = acc;
}

// ---- E. shared-memory table lookups: all lanes read 8 B from the SAME 128-byte table row -----------
__global__ void lds_table_kernel(double *out, int iters, int groups)
{
    extern __shared__ double tab[];   // groups x 16 doubles
    for (int i = threadIdx.x; i < groups * 16; i += blockDim.x) tab[i] = i;
    __syncthreads();
    unsigned bits = threadIdx.x * 2654435761u;
    double s1 = 0, s2 = 0;
    const int warp = threadIdx.x >> 5;
    for (int it = 0; it < iters; ++it) {
        int g = (warp * 7 + it * 4) % (groups - 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s1 += tab[(g + k) * 16 + ((bits >> (4 * k)) & 15)];
            s2 += tab[(g + k) * 16 + ((bits >> (16 + 4 * k)) & 15)];
        }
        bits = bits * 1664525u + 1013904223u;
    }
    if (s1 + s2 == 1.2345) out[0] = s1;
}

// ---- F. L2 reductions: red.global.add.f32 scalar vs v4 ------------------------------------------------
__global__ void red_f32_kernel(float *dst, int n, int iters)
{
    for (int it = 0; it < iters; ++it)
        for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(dst + i, 1.0f);
}
__global__ void red_v4_kernel(float *dst, int n, int iters)
{
    for (int it = 0; it < iters; ++it)
        for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "f"(1.0f), "f"(1.0f), "f"(1.0f), "f"(1.0f) : "memory");
}
__global__ void red_s32_kernel(int *dst, int n, int iters)
{
    for (int it = 0; it < iters; ++it)
        for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(dst + i, 1);
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, g_sms);
    long long *dcyc; CK(cudaMalloc(&dcyc, 64));
    double *dd; CK(cudaMalloc(&dd, 1 << 20)); CK(cudaMemset(dd, 0, 1 << 20));
    {   // SM clock under load
        long long h[2];
        float ms = time_ms([&] { clk_kernel<<<g_sms, 256>>>(dcyc, 4000000); }, 2);
        CK(cudaMemcpy(h, dcyc, 16, cudaMemcpyDeviceToHost));
        g_clk_hz = double(h[0]) / (ms * 1e-3);
        printf("sm clock %.0f MHz\n", g_clk_hz / 1e6);
    }
    const int T = 512, IT = 20000;
    auto rate = [&](const char *name, float ms, double ops_per_thread) {
        double per_sm_clk = ops_per_thread * T / (ms * 1e-3 * g_clk_hz);
        printf("%-28s %8.3f ms  %8.2f lane-ops/clk/SM\n", name, ms, per_sm_clk);
    };
    rate("DADD", time_ms([&] { dadd_kernel<<<g_sms, T>>>(dd, IT, 0); }), 8.0 * IT);
    rate("DADD predicated(and+setp)", time_ms([&] { dadd_pred_kernel<<<g_sms, T>>>(dd, IT, 0x5a5a5a5a); }), 8.0 * IT);
    rate("DFMA", time_ms([&] { dfma_kernel<<<g_sms, T>>>(dd, IT); }), 8.0 * IT);
    rate("POPC(+xor+add)", time_ms([&] { popc_kernel<<<g_sms, T>>>((unsigned *)dd, IT); }), 8.0 * IT);
    rate("DP4A", time_ms([&] { dp4a_kernel<<<g_sms, T>>>((int *)dd, IT); }), 8.0 * IT);
    rate("LOP3 x2", time_ms([&] { lop3_kernel<<<g_sms, T>>>((unsigned *)dd, IT); }), 8.0 * IT);
    {
        float ms = time_ms([&] { imma_kernel<<<g_sms, T>>>((int *)dd, IT); });
        double macs = 4.0 * IT * (T / 32) * 16 * 8 * 32;
        printf("%-28s %8.3f ms  %8.1f MAC/clk/SM\n", "IMMA.s8 m16n8k32 (legacy)", ms, macs / (ms * 1e-3 * g_clk_hz));
    }
#ifdef WITH_BMMA
    {
        float ms = time_ms([&] { bmma_kernel<<<g_sms, T>>>((int *)dd, IT); });
        double macs = 4.0 * IT * (T / 32) * 16 * 8 * 256;
        printf("%-28s %8.3f ms  %8.1f bit-MAC/clk/SM\n", "BMMA.b1 m16n8k256 and.popc", ms, macs / (ms * 1e-3 * g_clk_hz));
    }
#endif
    {
        unsigned *bar; CK(cudaMalloc(&bar, 256));
        for (int threads : {128, 512}) {
            CK(cudaMemset(bar, 0, 256));
            const int iters = 2000;
            void *args[] = {&bar, (void *)&iters, &dcyc};
            float ms = time_ms([&] { cudaMemsetAsync(bar, 0, 4); cudaLaunchCooperativeKernel((void *)gridbar_kernel, dim3(g_sms), dim3(threads), args, 0, 0); }, 2);
            printf("grid barrier (%d CTAs x %d thr)  %.3f us per barrier\n", g_sms, threads, ms * 1e3 / iters);
            double *part; CK(cudaMalloc(&part, sizeof(double) * 2 * g_sms * 32));
            void *args2[] = {&bar, &part, (void *)&iters, &dcyc, &dd};
            ms = time_ms([&] { cudaMemsetAsync(bar, 0, 4); cudaLaunchCooperativeKernel((void *)gridbar_xchg_kernel, dim3(g_sms), dim3(threads), args2, 0, 0); }, 2);
            printf("grid barrier + 32-double exchange (%d thr)  %.3f us per round\n", threads, ms * 1e3 / iters);
            cudaFree(part);
        }
        cudaFree(bar);
    }
    {
        const int groups = 848, iters = 20000;
        CK(cudaFuncSetAttribute(lds_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, groups * 128));
        float ms = time_ms([&] { lds_table_kernel<<<g_sms, T, groups * 128>>>(dd, iters, groups); });
        rate("LDS.64 table lookup (+DADD)", ms, 8.0 * iters);
    }
    {
        float *dst; const int n = 8192; CK(cudaMalloc(&dst, n * 4)); CK(cudaMemset(dst, 0, n * 4));
        const int iters = 200;
        float ms = time_ms([&] { red_f32_kernel<<<g_sms, 512>>>(dst, n, iters); });
        printf("RED.f32 scalar: %d CTAs x %d values  %.3f us per round  (%.1f REDs/clk chip)\n", g_sms, n, ms * 1e3 / iters, double(g_sms) * n * iters / (ms * 1e-3 * g_clk_hz));
        ms = time_ms([&] { red_v4_kernel<<<g_sms, 512>>>(dst, n, iters); });
        printf("RED.v4.f32:     %d CTAs x %d values  %.3f us per round  (%.1f values/clk chip)\n", g_sms, n, ms * 1e3 / iters, double(g_sms) * n * iters / (ms * 1e-3 * g_clk_hz));
        ms = time_ms([&] { red_s32_kernel<<<g_sms, 512>>>((int *)dst, n, iters); });
        printf("RED.s32 scalar: %d CTAs x %d values  %.3f us per round  (%.1f REDs/clk chip)\n", g_sms, n, ms * 1e3 / iters, double(g_sms) * n * iters / (ms * 1e-3 * g_clk_hz));
        cudaFree(dst);
    }
    return 0;
}
