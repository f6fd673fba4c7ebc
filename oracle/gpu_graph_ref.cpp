// TEST INFRASTRUCTURE ONLY (compiled into oracle/_ref/libbrr_ref_gpu.so by oracle/Makefile, against the reference's own headers).
//
// The reference-side binding of INTEGRATION.md, compiled for real: a new AnalysisGraph (src/analysisgraph.hpp:9-26) whose exec()
// hands the whole sweep to the C ABI of include/brr_gpu.h.  With it, the UNMODIFIED reference driver BayesRBase::runGibbs
// (src/BayesRBase.cpp:130-285: its own init, its own RNG draws, updateMu, shuffle, sigmaE / sigmaG / pi updates) runs its Markov
// chain on the GPU engine; the only non-reference code between runGibbs and the kernel is this class.
//
// Upstream the class would be a `friend` of the analyses, exactly like LimitSequenceGraph / DenseParallelGraph
// (src/DenseBayesRRmz.hpp:15-16, src/SparseBayesRRG.hpp:10).  The reference tree is read-only here, so the protected chain state
// is reached through pointers to members of a derived class instead (legal C++, no header is touched).
#include "BayesRBase.hpp"
#include "analysisgraph.hpp"
#include "common.h"
#include "data.hpp"
#include "options.hpp"

#include "../include/brr_gpu.h"

#include <cstdint>
#include <iostream>
#include <vector>

namespace {
struct Access : BayesRBase {
    using BayesRBase::m_epsilon; using BayesRBase::m_beta; using BayesRBase::m_components; using BayesRBase::m_acum;
    using BayesRBase::m_pi; using BayesRBase::m_v; using BayesRBase::m_sigmaG; using BayesRBase::m_sigmaE;
    using BayesRBase::m_betasqnG; using BayesRBase::m_epsilonSum; using BayesRBase::m_randomNumbers;
};
#define FIELD(b, name) ((b)->*(&Access::name))
}

class GpuSweepGraph : public AnalysisGraph
{
public:
    // bedPayload: the PLINK SNP-major payload (no 3-byte magic) of Options::dataFile, numSnps columns of ceil(numInds/4) bytes
    GpuSweepGraph(const Data *data, const Options *opt, bool async, const uint8_t *bedPayload)
    {
        brr_gpu_config cfg{};
        cfg.abi_version = BRR_GPU_ABI_VERSION;
        cfg.layout = opt->preprocessDataType == PreprocessDataType::Dense ? BRR_GPU_LAYOUT_DENSE2BIT
                   : opt->preprocessDataType == PreprocessDataType::SparseEigen ? BRR_GPU_LAYOUT_EIGEN : BRR_GPU_LAYOUT_RAGGED;
        cfg.n_inds = cfg.n_inds_total = data->numInds;
        cfg.n_markers = data->numSnps;
        cfg.n_components = unsigned(opt->S.cols()) + 1;
        cfg.n_groups = data->numGroups;
        cfg.world = 1;
        if (brr_gpu_create(&cfg, &m_ctx)) { m_error = brr_gpu_last_error(); m_ctx = nullptr; return; }
        // the .bed payload IS the device layout: no .ppbed is written (replaces PreprocessGraph and MarkerCache::populate)
        const size_t stride = (size_t(data->numInds) + 3) / 4;
        int rc = cfg.layout == BRR_GPU_LAYOUT_DENSE2BIT ? brr_gpu_upload_dense2bit(m_ctx, bedPayload, stride, 0, data->numSnps)
                                                        : brr_gpu_upload_ragged_from_bed(m_ctx, bedPayload, stride);
        if (rc || brr_gpu_finalize_markers(m_ctx, nullptr)) { m_error = brr_gpu_last_error(); return; }
        std::vector<int32_t> groups(data->G.begin(), data->G.end());
        Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> cva = opt->S;      // nG x (K-1)
        if (brr_gpu_set_model(m_ctx, groups.data(), cva.data())) { m_error = brr_gpu_last_error(); return; }
        m_tokens = async ? unsigned(opt->analysisTokens) : 0u;
    }
    ~GpuSweepGraph() override { if (m_ctx) brr_gpu_destroy(m_ctx); }

    bool isAsynchronous() const override { return false; }     // the device reproduces the sync chain; staleness is a sweep argument
    bool ok() const { return m_ctx && m_error.empty(); }
    const std::string &error() const { return m_error; }
    unsigned sweeps() const { return m_sweeps; }

    void exec(Analysis *analysis, unsigned, unsigned numSnps, const std::vector<unsigned> &markerIndices) override
    {
        auto *b = dynamic_cast<BayesRBase *>(analysis);
        if (!b || !ok()) { std::cerr << "GpuSweepGraph: " << (b ? m_error : std::string("needs a BayesRBase analysis")) << std::endl; return; }
        // runGibbs touched epsilon on the host (updateMu, src/DenseBayesRRmz.cpp:36-42): push it
        brr_gpu_set_state(m_ctx, FIELD(b, m_epsilon).data(), m_stateOnDevice ? nullptr : FIELD(b, m_beta).data(),
                          m_stateOnDevice ? nullptr : FIELD(b, m_components).data(), nullptr);
        m_stateOnDevice = true;
        using RowMat = Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>;
        RowMat pi = FIELD(b, m_pi), v(FIELD(b, m_v).rows(), FIELD(b, m_v).cols());
        brr_gpu_sweep_out out{};
        static_assert(sizeof(FIELD((BayesRBase *)nullptr, m_randomNumbers)[0]) == 2 * sizeof(double), "(p, z) pairs");
        if (brr_gpu_sweep(m_ctx, markerIndices.data(), numSnps, &FIELD(b, m_randomNumbers)[0][0], pi.data(), FIELD(b, m_sigmaG).data(),
                          FIELD(b, m_sigmaE), FIELD(b, m_epsilonSum), m_tokens, v.data(), FIELD(b, m_betasqnG).data(), &out)) {
            m_error = brr_gpu_last_error();
            std::cerr << "brr_gpu_sweep: " << m_error << std::endl;           // no exceptions across exec()
            return;
        }
        FIELD(b, m_v) = v;
        FIELD(b, m_epsilonSum) = out.eps_sum;                                  // SparseBayesRRG's tracked sum
        brr_gpu_get_state(m_ctx, FIELD(b, m_epsilon).data(), FIELD(b, m_beta).data(), FIELD(b, m_components).data(), FIELD(b, m_acum).data());
        ++m_sweeps;
    }

private:
    brr_gpu_ctx *m_ctx = nullptr;
    unsigned m_tokens = 0, m_sweeps = 0;
    bool m_stateOnDevice = false;
    std::string m_error;
};

// factory for the harness (the reference's factory is AnalysisRunner::makeAnalysisGraph, src/analysisrunner.cpp:246-276)
AnalysisGraph *makeGpuSweepGraph(const Data *data, const Options *opt, bool async, const uint8_t *bedPayload, std::string *err)
{
    auto *g = new GpuSweepGraph(data, opt, async, bedPayload);
    if (!g->ok()) { if (err) *err = g->error(); delete g; return nullptr; }
    return g;
}
