// Kernels around the device-resident genotype store: PLINK -> bit planes, class counts, value
// tables, synthetic data, and the marker-level parity hooks (computeNum / epsilon change).
#pragma once
#include "common.cuh"

namespace brr {

// ---- counter-based synthetic genotypes (shared by host and device) ------------------------------
__host__ __device__ inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}
// thr = {t0, t1, tmiss}: class 0 if r < t0, 1 if r < t1, else 2; missing if r2 < tmiss
__host__ __device__ inline uint32_t synth_class(uint64_t seed, uint32_t marker, uint32_t ind, const uint32_t *thr)
{
    const uint64_t h = mix64((seed ^ (uint64_t(marker) * 0x9E3779B97F4A7C15ULL)) + uint64_t(ind) * 0xD1B54A32D192ED03ULL);
    const uint32_t r = uint32_t(h >> 32), r2 = uint32_t(h);
    if (r2 < thr[2]) return 3u;
    return r < thr[0] ? 0u : (r < thr[1] ? 1u : 2u);
}

// Store layout: planes[((marker*nCTA + cta)*2 + plane)*PB + byte], bit i of a slice = individual
// cta*SL + i.  P1 = {class 1 or missing}, P2 = {class 2 or missing}.

// PLINK payload -> planes. grid = (nCTA, markers in chunk), threads over plane bytes.
__global__ void bed_to_planes_kernel(const uint8_t *__restrict__ bed, size_t col_bytes, uint32_t first_marker,
                                     uint32_t n_local, uint32_t nCTA, uint32_t SL, uint32_t PB,
                                     uint8_t *__restrict__ planes)
{
    const uint32_t cta = blockIdx.x, mk = blockIdx.y;
    const uint8_t *col = bed + size_t(mk) * col_bytes;
    uint8_t *dst = planes + (size_t(first_marker + mk) * nCTA + cta) * 2 * PB;
    for (uint32_t u = threadIdx.x; u < PB; u += blockDim.x) {
        const uint32_t i0 = cta * SL + 8 * u;
        uint32_t p1 = 0, p2 = 0;
        if (8 * u < SL && i0 < n_local) {
            uint32_t two16 = col[i0 >> 2];
            if ((i0 >> 2) + 1 < col_bytes) two16 |= uint32_t(col[(i0 >> 2) + 1]) << 8;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (i0 + t < n_local) {
                    const uint32_t b0 = (two16 >> (2 * t)) & 1u, b1 = (two16 >> (2 * t + 1)) & 1u;
                    // PLINK 00->2, 10->1 (b1=1,b0=0), 11->0, 01->missing (src/preprocessgraph.cpp:50-51)
                    p1 |= (b0 ^ b1) << t;        // class 1 or missing
                    p2 |= (b1 ^ 1u) << t;        // class 2 or missing
                }
            }
        }
        dst[u] = uint8_t(p1);
        dst[PB + u] = uint8_t(p2);
    }
}

// synthetic genotypes straight into the planes. grid = (nCTA, min(M, 65535)) with a marker loop.
__global__ void synth_planes_kernel(uint64_t seed, const uint32_t *__restrict__ thr, uint32_t M, uint32_t n_local,
                                    uint32_t ind_offset, uint32_t nCTA, uint32_t SL, uint32_t PB,
                                    uint8_t *__restrict__ planes)
{
    const uint32_t cta = blockIdx.x;
    for (uint32_t mk = blockIdx.y; mk < M; mk += gridDim.y) {
        uint8_t *dst = planes + (size_t(mk) * nCTA + cta) * 2 * PB;
        const uint32_t t3[3] = {thr[3 * mk], thr[3 * mk + 1], thr[3 * mk + 2]};
        for (uint32_t u = threadIdx.x; u < PB; u += blockDim.x) {
            const uint32_t i0 = cta * SL + 8 * u;
            uint32_t p1 = 0, p2 = 0;
            if (8 * u < SL) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    if (i0 + t < n_local) {
                        const uint32_t c = synth_class(seed, mk, ind_offset + i0 + t, t3);
                        p1 |= uint32_t(c == 1u || c == 3u) << t;
                        p2 |= uint32_t(c == 2u || c == 3u) << t;
                    }
                }
            }
            dst[u] = uint8_t(p1);
            dst[PB + u] = uint8_t(p2);
        }
    }
}

// class counts of this shard: one warp per marker. counts[marker] = {n0, n1, n2, nmiss}
__global__ void count_classes_kernel(const uint8_t *__restrict__ planes, uint32_t first, uint32_t count,
                                     uint32_t n_local, uint32_t nCTA, uint32_t PB, uint32_t *__restrict__ counts)
{
    const uint32_t warps_per_block = blockDim.x / 32, lane = threadIdx.x & 31;
    const uint32_t mk = first + blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (mk >= first + count) return;
    const uint32_t PW = PB / 4;
    uint32_t n1 = 0, n2 = 0, nm = 0;
    for (uint32_t i = lane; i < nCTA * PW; i += 32) {
        const uint32_t cta = i / PW, w = i % PW;
        const uint32_t *base = reinterpret_cast<const uint32_t *>(planes + (size_t(mk) * nCTA + cta) * 2 * PB);
        const uint32_t x1 = base[w], x2 = base[PW + w];
        n1 += __popc(x1 & ~x2); n2 += __popc(x2 & ~x1); nm += __popc(x1 & x2);
    }
    n1 = __reduce_add_sync(0xffffffffu, n1); n2 = __reduce_add_sync(0xffffffffu, n2); nm = __reduce_add_sync(0xffffffffu, nm);
    if (lane == 0) {
        counts[4 * size_t(mk) + 0] = n_local - n1 - n2 - nm;
        counts[4 * size_t(mk) + 1] = n1; counts[4 * size_t(mk) + 2] = n2; counts[4 * size_t(mk) + 3] = nm;
    }
}

// DenseMarkerBuilder::endColumn (reference src/densemarkerbuilder.cpp:43-60) in closed form on the
// class counts: impute missing with the mean of the non-missing, centre with the mean of the
// imputed column, scale by 1/sqrt(||.||^2/(N-1)).
__global__ void dense_lut_kernel(const uint32_t *__restrict__ counts, uint32_t M, uint32_t n_total,
                                 double *__restrict__ lut)
{
    const uint32_t mk = blockIdx.x * blockDim.x + threadIdx.x;
    if (mk >= M) return;
    const double n0 = counts[4 * size_t(mk)], n1 = counts[4 * size_t(mk) + 1], n2 = counts[4 * size_t(mk) + 2],
                 nm = counts[4 * size_t(mk) + 3];
    const double N = double(n_total);
    const double sum = n1 + 2.0 * n2;
    const double mean = sum / (N - nm);                     // :50
    const double mean2 = (sum + nm * mean) / N;             // snpData.mean(), :56
    const double c0 = 0.0 - mean2, c1 = 1.0 - mean2, c2 = 2.0 - mean2, cm = mean - mean2;
    const double sqn = n0 * c0 * c0 + n1 * c1 * c1 + n2 * c2 * c2 + nm * cm * cm;   // :57
    const double sigma = 1.0 / sqrt(sqn / (N - 1.0));       // :58
    lut[4 * size_t(mk) + 0] = c0 * sigma; lut[4 * size_t(mk) + 1] = c1 * sigma;
    lut[4 * size_t(mk) + 2] = c2 * sigma; lut[4 * size_t(mk) + 3] = cm * sigma;
}

// ---- marker-level parity hooks -------------------------------------------------------------------
// DenseRKernel::computeNum without the beta_old*(N-1) term (reference src/densebayesrkernel.cpp:10-15):
// one block per slice, deterministic block reduction, partial[cta] = sum_i lut[class_i] * eps_i.
__global__ void dense_dot_kernel(const uint8_t *__restrict__ planes, const double *__restrict__ lut, uint32_t marker,
                                 const double *__restrict__ eps, uint32_t n_local, uint32_t nCTA, uint32_t SL,
                                 uint32_t PB, double *__restrict__ partial)
{
    __shared__ double wsum[32];
    const uint32_t cta = blockIdx.x;
    const uint8_t *p1 = planes + (size_t(marker) * nCTA + cta) * 2 * PB, *p2 = p1 + PB;
    const double L0 = lut[4 * size_t(marker)], L1 = lut[4 * size_t(marker) + 1], L2 = lut[4 * size_t(marker) + 2],
                 L3 = lut[4 * size_t(marker) + 3];
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < SL; i += blockDim.x) {
        const uint32_t gi = cta * SL + i;
        if (gi < n_local) {
            const uint32_t b1 = (p1[i >> 3] >> (i & 7)) & 1u, b2 = (p2[i >> 3] >> (i & 7)) & 1u;
            const double x = b1 ? (b2 ? L3 : L1) : (b2 ? L2 : L0);
            acc += x * eps[gi];
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (uint32_t w = 0; w < blockDim.x / 32; ++w) s += wsum[w];
        partial[cta] = s;
    }
}

// eps += dB * x_marker (dB = beta_old - beta_new; reference src/densebayesrkernel.cpp:17-20 + BayesRBase.cpp:376)
__global__ void dense_axpy_kernel(const uint8_t *__restrict__ planes, const double *__restrict__ lut, uint32_t marker,
                                  double dB, double *__restrict__ eps, uint32_t n_local, uint32_t nCTA, uint32_t SL,
                                  uint32_t PB)
{
    const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= n_local) return;
    const uint32_t cta = gi / SL, i = gi % SL;
    const uint8_t *p1 = planes + (size_t(marker) * nCTA + cta) * 2 * PB, *p2 = p1 + PB;
    const uint32_t b1 = (p1[i >> 3] >> (i & 7)) & 1u, b2 = (p2[i >> 3] >> (i & 7)) & 1u;
    eps[gi] += dB * lut[4 * size_t(marker) + (b1 | (b2 << 1))];
}

// eps += c, then per-block partial sums of eps and eps^2 (deterministic); c = 0 for a plain reduce
__global__ void eps_add_reduce_kernel(double *__restrict__ eps, uint32_t n, double c, int do_add,
                                      double *__restrict__ partial /*[grid][2]*/)
{
    __shared__ double ws[32], wq[32];
    double s = 0.0, q = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double e = eps[i];
        if (do_add) { e += c; eps[i] = e; }
        s += e; q += e * e;
    }
    s = warp_sum(s); q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = s; wq[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (uint32_t w = 0; w < blockDim.x / 32; ++w) { a += ws[w]; b += wq[w]; }
        partial[2 * blockIdx.x] = a; partial[2 * blockIdx.x + 1] = b;
    }
}

} // namespace brr
