// Kernels around the device-resident genotype store: PLINK -> 2-bit codes, class counts, value
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

// Store layout (HBM): one ROW per (marker, CTA slice), PWT = PW + 8 words, at codes[(marker*nCTA + cta)*PWT]:
//   words [0, PW), PW = SL/16: word w holds the 2-bit codes of the 16 individuals cta*SL + 16w .. +15
//     (individual j of the word at bits 2j..2j+1).  Code = allele count 0,1,2; 3 = missing.  Individuals
//     past n_local get code 0.
//   words [PW, PW+8): the row's missing list as 16 x uint16: [0] = number of missing genotypes in the row
//     (the true count, may exceed 15), [1..15] = the first slice-local individual ids, ascending.
//   A row is fetched by ONE TMA bulk copy; the list lets the sweep correct for missing genotypes without
//   scanning the codes (rows with more than 15 are rescanned by the warp that owns them).
constexpr uint32_t kTailWords = 8;
// The PLINK code (b1 b0) maps as 00->2, 10->1, 11->0, 01->missing (reference src/preprocessgraph.cpp:50-51,
// src/densemarkerbuilder.cpp:30): new_hi = ~b1, new_lo = b1 ^ b0.

// PLINK payload -> codes. grid = (nCTA, markers in chunk), threads over words.
__global__ void bed_to_codes_kernel(const uint8_t *__restrict__ bed, size_t col_bytes, uint32_t first_marker,
                                    uint32_t n_local, uint32_t nCTA, uint32_t SL, uint32_t PW,
                                    uint32_t *__restrict__ codes)
{
    const uint32_t cta = blockIdx.x, mk = blockIdx.y;
    const uint8_t *col = bed + size_t(mk) * col_bytes;
    uint32_t *dst = codes + (size_t(first_marker + mk) * nCTA + cta) * (PW + kTailWords);
    for (uint32_t w = threadIdx.x; w < PW; w += blockDim.x) {
        const uint32_t i0 = cta * SL + 16 * w;
        uint32_t out = 0;
        if (i0 < n_local) {
            uint32_t raw = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const size_t byte = size_t(i0 >> 2) + k;
                if (byte < col_bytes) raw |= uint32_t(col[byte]) << (8 * k);
            }
            const uint32_t hi = raw & 0xAAAAAAAAu, lo = raw & 0x55555555u;
            out = (~hi & 0xAAAAAAAAu) | ((lo ^ (hi >> 1)) & 0x55555555u);
            const uint32_t valid = n_local - i0;                 // individuals of this word that exist
            if (valid < 16) out &= (1u << (2 * valid)) - 1u;
        }
        dst[w] = out;
    }
}

// synthetic genotypes straight into the store. grid = (nCTA, min(M, 65535)) with a marker loop.
__global__ void synth_codes_kernel(uint64_t seed, const uint32_t *__restrict__ thr, uint32_t M, uint32_t n_local,
                                   uint32_t ind_offset, uint32_t nCTA, uint32_t SL, uint32_t PW,
                                   uint32_t *__restrict__ codes)
{
    const uint32_t cta = blockIdx.x;
    for (uint32_t mk = blockIdx.y; mk < M; mk += gridDim.y) {
        uint32_t *dst = codes + (size_t(mk) * nCTA + cta) * (PW + kTailWords);
        const uint32_t t3[3] = {thr[3 * mk], thr[3 * mk + 1], thr[3 * mk + 2]};
        for (uint32_t w = threadIdx.x; w < PW; w += blockDim.x) {
            const uint32_t i0 = cta * SL + 16 * w;
            uint32_t out = 0;
#pragma unroll
            for (int t = 0; t < 16; ++t)
                if (i0 + t < n_local) out |= synth_class(seed, mk, ind_offset + i0 + t, t3) << (2 * t);
            dst[w] = out;
        }
    }
}

// ---- synthetic genotypes straight into the ragged / eigen store -----------------------------------------------------
// The same counter-based genotypes as synth_codes_kernel, as index lists.  Pass 1 counts the ones / twos / missing of
// every (CTA slice, marker); after a per-slice prefix sum on the host (seg), pass 2 writes the slice-local 16-bit ids in
// ascending order.  One warp per (slice, marker).
__global__ void synth_ragged_count_kernel(uint64_t seed, const uint32_t *__restrict__ thr, uint32_t M, uint32_t n_local,
                                          uint32_t ind_offset, uint32_t nCTA, uint32_t SL, uint32_t *__restrict__ cnt3)
{
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const uint32_t cta = blockIdx.x;
    for (uint32_t mk = blockIdx.y * wpb + wib; mk < M; mk += gridDim.y * wpb) {
        const uint32_t t3[3] = {thr[3 * mk], thr[3 * mk + 1], thr[3 * mk + 2]};
        uint32_t n1 = 0, n2 = 0, nm = 0;
        for (uint32_t i = lane; i < SL; i += 32) {
            const uint32_t gi = cta * SL + i;
            const uint32_t cls = gi < n_local ? synth_class(seed, mk, ind_offset + gi, t3) : 0u;
            n1 += cls == 1u; n2 += cls == 2u; nm += cls == 3u;
        }
        n1 = __reduce_add_sync(0xffffffffu, n1); n2 = __reduce_add_sync(0xffffffffu, n2); nm = __reduce_add_sync(0xffffffffu, nm);
        if (lane == 0) { uint32_t *o = cnt3 + (size_t(cta) * M + mk) * 3; o[0] = n1; o[1] = n2; o[2] = nm; }
    }
}

__global__ void synth_ragged_fill_kernel(uint64_t seed, const uint32_t *__restrict__ thr, uint32_t M, uint32_t n_local,
                                         uint32_t ind_offset, uint32_t nCTA, uint32_t SL, const uint32_t *__restrict__ seg,
                                         const uint64_t *__restrict__ cta_base, uint16_t *__restrict__ idx16)
{
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const uint32_t cta = blockIdx.x;
    uint16_t *dst = idx16 + cta_base[cta];
    for (uint32_t mk = blockIdx.y * wpb + wib; mk < M; mk += gridDim.y * wpb) {
        const uint32_t t3[3] = {thr[3 * mk], thr[3 * mk + 1], thr[3 * mk + 2]};
        const uint32_t *sg = seg + (size_t(cta) * M + mk) * 4;
        uint32_t pos[3] = {sg[0], sg[1], sg[2]};                 // write cursors of the ones / twos / missing lists
        for (uint32_t i0 = 0; i0 < SL; i0 += 32) {
            const uint32_t i = i0 + lane, gi = cta * SL + i;
            const uint32_t cls = (i < SL && gi < n_local) ? synth_class(seed, mk, ind_offset + gi, t3) : 0u;
#pragma unroll
            for (uint32_t l = 0; l < 3; ++l) {
                const unsigned m = __ballot_sync(0xffffffffu, cls == l + 1u);
                if (cls == l + 1u) dst[pos[l] + __popc(m & ((1u << lane) - 1u))] = uint16_t(i);
                pos[l] += __popc(m);
            }
        }
    }
}

// missing lists of the rows [first_row, first_row + n_rows): one warp per row, ordered compaction
__global__ void build_row_tails_kernel(uint32_t *__restrict__ codes, size_t first_row, size_t n_rows, uint32_t PW)
{
    const uint32_t lane = threadIdx.x & 31;
    const size_t r = size_t(blockIdx.x) * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    uint32_t *row = codes + (first_row + r) * (PW + kTailWords);
    uint16_t *tail = reinterpret_cast<uint16_t *>(row + PW);
    uint32_t cnt = 0;
    if (lane < 16) tail[lane] = 0;
    __syncwarp();
    for (uint32_t w0 = 0; w0 < PW; w0 += 32) {
        const uint32_t w = w0 + lane;
        const uint32_t x = (w < PW) ? row[w] : 0u;
        uint32_t mm = x & (x >> 1) & 0x55555555u;
        const uint32_t n = __popc(mm);
        unsigned bal = __ballot_sync(0xffffffffu, n != 0);
        while (bal) {
            const int l = __ffs(bal) - 1;
            bal &= bal - 1;
            const uint32_t nl = __shfl_sync(0xffffffffu, n, l);
            if (int(lane) == l) {
                uint32_t pos = cnt;
                while (mm) {
                    const int bit = __ffs(mm) - 1;
                    mm &= mm - 1;
                    if (pos < 15) tail[1 + pos] = uint16_t(16 * w + (bit >> 1));
                    ++pos;
                }
            }
            cnt += nl;
        }
    }
    if (lane == 0) tail[0] = uint16_t(cnt);
}

// class counts of this shard: one warp per marker. counts[marker] = {n0, n1, n2, nmiss}
__global__ void count_classes_kernel(const uint32_t *__restrict__ codes, uint32_t first, uint32_t count,
                                     uint32_t n_local, uint32_t nCTA, uint32_t PW, uint32_t *__restrict__ counts)
{
    const uint32_t warps_per_block = blockDim.x / 32, lane = threadIdx.x & 31;
    const uint32_t mk = first + blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (mk >= first + count) return;
    const uint32_t PWT = PW + kTailWords;
    const uint32_t *base = codes + size_t(mk) * nCTA * PWT;
    uint32_t n1 = 0, n2 = 0, nm = 0;
    for (uint32_t c = 0; c < nCTA; ++c)
        for (uint32_t w = lane; w < PW; w += 32) {
            const uint32_t x = base[size_t(c) * PWT + w];
            const uint32_t lo = x & 0x55555555u, hi = (x >> 1) & 0x55555555u;
            n1 += __popc(lo & ~hi); n2 += __popc(hi & ~lo); nm += __popc(lo & hi);
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

__device__ __forceinline__ uint32_t code_of(const uint32_t *__restrict__ codes, uint32_t marker, uint32_t nCTA,
                                            uint32_t SL, uint32_t PW, uint32_t gi)
{
    const uint32_t cta = gi / SL, i = gi - cta * SL;
    return (codes[(size_t(marker) * nCTA + cta) * (PW + kTailWords) + (i >> 4)] >> (2 * (i & 15u))) & 3u;
}

// ---- marker-level parity hooks -------------------------------------------------------------------
// DenseRKernel::computeNum without the beta_old*(N-1) term (reference src/densebayesrkernel.cpp:10-15):
// one block per slice, deterministic block reduction, partial[cta] = sum_i lut[code_i] * eps_i.
__global__ void dense_dot_kernel(const uint32_t *__restrict__ codes, const double *__restrict__ lut, uint32_t marker,
                                 const double *__restrict__ eps, uint32_t n_local, uint32_t nCTA, uint32_t SL,
                                 uint32_t PW, double *__restrict__ partial)
{
    __shared__ double wsum[32];
    const uint32_t cta = blockIdx.x;
    const uint32_t *row = codes + (size_t(marker) * nCTA + cta) * (PW + kTailWords);
    const double L[4] = {lut[4 * size_t(marker)], lut[4 * size_t(marker) + 1], lut[4 * size_t(marker) + 2],
                         lut[4 * size_t(marker) + 3]};
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < SL; i += blockDim.x) {
        const uint32_t gi = cta * SL + i;
        if (gi < n_local) {
            const uint32_t c = (row[i >> 4] >> (2 * (i & 15u))) & 3u;
            const double x = c == 0 ? L[0] : (c == 1 ? L[1] : (c == 2 ? L[2] : L[3]));
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
__global__ void dense_axpy_kernel(const uint32_t *__restrict__ codes, const double *__restrict__ lut, uint32_t marker,
                                  double dB, double *__restrict__ eps, uint32_t n_local, uint32_t nCTA, uint32_t SL,
                                  uint32_t PW)
{
    const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= n_local) return;
    eps[gi] += dB * lut[4 * size_t(marker) + code_of(codes, marker, nCTA, SL, PW, gi)];
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
