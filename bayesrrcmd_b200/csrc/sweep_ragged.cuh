// Ragged (index-list) layout: device store and sweep kernel.  (stage 1: store definition; the
// kernel follows in this file)
#pragma once
#include "common.cuh"
#include "sampler.cuh"

namespace brr {

struct RaggedStore {
    uint64_t *offsets = nullptr;   // [3M+1] into idx: ones_j, twos_j, missing_j
    int32_t *idx = nullptr;        // shard-local individual ids, ascending per list
    double *stats = nullptr;       // [M][4] mean, sd, sqrdZ, Zsum
    uint32_t *seg = nullptr;       // [M][3][nCTA+1] per-slice split points of every list
    size_t idx_bytes = 0;
    bool stats_given = false;
};

__host__ __device__ inline size_t ragged_smem_bytes(uint32_t SL) { return size_t(SL) * sizeof(double) + 65536; }

__global__ void sweep_ragged_kernel() {}

// RaggedSparseMarkerBuilder::endColumn (reference src/raggedsparsemarkerbuilder.cpp:43-55) on class counts
__global__ void ragged_stats_kernel(const uint32_t *__restrict__ counts, uint32_t M, uint32_t n_total,
                                    double *__restrict__ stats)
{
    const uint32_t mk = blockIdx.x * blockDim.x + threadIdx.x;
    if (mk >= M) return;
    const double n1 = counts[4 * size_t(mk) + 1], n2 = counts[4 * size_t(mk) + 2];
    const double N = double(n_total);
    const double Zsum = n1 + 2.0 * n2, sqrdZ = n1 + 4.0 * n2;
    const double mean = Zsum / N;                                                          // :49
    const double sd = sqrt((sqrdZ - 2.0 * mean * Zsum + N * mean * mean) / (N - 1.0));    // :53-54
    stats[4 * size_t(mk) + 0] = mean; stats[4 * size_t(mk) + 1] = sd;
    stats[4 * size_t(mk) + 2] = sqrdZ; stats[4 * size_t(mk) + 3] = Zsum;
}

} // namespace brr
