// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// Harness around the UNMODIFIED reference sources (compiled where they lie under
// /root/reference by oracle/Makefile, outputs only into oracle/_ref/).  Nothing in this file is
// copied from the reference: it only #includes the reference's public headers, builds markers
// with the reference's own MarkerBuilders, calls the reference's own BayesRKernel virtuals, and
// drives the reference's own BayesRBase::runGibbs through the reference's AnalysisGraph plugin
// interface (src/analysisgraph.hpp:9-26) with a tap that records the chain state around every
// exec() call.  Exposed as a flat C API so tests / the golden generator / bench.py's
// `--impl reference` arm can reach it with ctypes.
//
// Used for: (1) validating oracle/bayesr_oracle.c (the C restatement), (2) generating
// tests/golden/*.npz (tools under oracle/make_golden.py), (3) the CPU baseline ("kind":"reference").

#include "BayesRBase.hpp"
#include "DenseBayesRRmz.hpp"
#include "SparseBayesRRG.hpp"
#include "analysisgraph.hpp"
#include "parallelgraph.h"
#include "markercache.h"
#include "bayesrkernel.h"
#include "sparsebayesrkernel.h"
#include "eigenbayesrkernel.h"
#include "densemarker.h"
#include "raggedsparsemarker.h"
#include "eigensparsemarker.h"
#include "markerbuilder.h"
#include "common.h"
#include "compression.h"

#include <chrono>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <memory>
#include <vector>

namespace {

PreprocessDataType layoutType(int layout)
{
    switch (layout) {
    case 0: return PreprocessDataType::Dense;
    case 1: return PreprocessDataType::SparseRagged;
    case 2: return PreprocessDataType::SparseEigen;
    default: return PreprocessDataType::None;
    }
}

// One PLINK .bed column -> reference Marker, through the reference's builder, feeding alleles the
// way the reference's preprocess step does (src/preprocessgraph.cpp:44-53).
ConstMarkerPtr buildMarker(MarkerBuilder *builder, unsigned snp, const uint8_t *bedCol, unsigned N)
{
    builder->initialise(snp, N);
    for (unsigned i = 0; i < N; ++i) {
        const unsigned byte = bedCol[i >> 2];
        const unsigned k = (i & 3u) * 2u;
        const unsigned allele1 = !((byte >> k) & 1u);
        const unsigned allele2 = !((byte >> (k + 1)) & 1u);
        builder->processAllele(i, allele1, allele2);
    }
    builder->endColumn();
    return ConstMarkerPtr(builder->build());
}

struct ChainDump {
    // per iteration, captured immediately before exec()
    double *mu;         // [iters]
    double *sigmaE_in;  // [iters]
    double *sigmaG_in;  // [iters*nG]
    double *pi_in;      // [iters*nG*K]
    double *rand_uz;    // [iters*M*2]
    uint32_t *order;    // [iters*M]
    double *eps_in;     // [iters*N]
    double *beta_in;    // [iters*M]
    double *epssum_in;  // [iters]
    // per iteration, captured immediately after exec()
    double *beta_out;   // [iters*M]
    double *comp_out;   // [iters*M]
    double *acum_out;   // [iters*M]
    double *eps_out;    // [iters*N]
    double *v_out;      // [iters*nG*K]
    double *betasq_out; // [iters*nG]
    double *epssum_out; // [iters]
    // per iteration, after the sweep epilogue (captured at the next exec, or at the end)
    double *sigmaE_post; // [iters]
    double *sigmaG_post; // [iters*nG]
    double *pi_post;     // [iters*nG*K]
    double *gamma_post;  // [iters*nF] fixed effects (src/BayesRBase.cpp:214-235), may be NULL
    double *eps_post;    // [iters*N] the residual after the fixed-effects loop, may be NULL
};

struct TapIface {
    virtual ~TapIface() = default;
    virtual void before(unsigned it, const std::vector<unsigned> &order) = 0;
    virtual void after(unsigned it) = 0;
    virtual void post(unsigned it) = 0;
};

// All chain state is protected in BayesRBase (src/BayesRBase.hpp:43-112): a subclass sees it.
template <class Base>
struct Tap : Base, TapIface {
    Tap(const Data *d, const Options *o, ChainDump *dump, unsigned M, unsigned N, unsigned nG, unsigned K)
        : Base(d, o), dump(dump), M(M), N(N), nG(nG), K(K)
    {
        // src/BayesRBase.cpp:218 reads m_sigmaF in the first iteration before anything has set it (:234 sets it to s02F afterwards):
        // the harness starts it at s02F, which is what every later iteration sees
        this->m_sigmaF = this->s02F;
    }

    void before(unsigned it, const std::vector<unsigned> &order) override
    {
        if (!dump) return;
        if (it > 0) post(it - 1);
        dump->mu[it] = this->m_mu;
        dump->sigmaE_in[it] = this->m_sigmaE;
        dump->epssum_in[it] = this->m_epsilonSum;
        for (unsigned g = 0; g < nG; ++g) {
            dump->sigmaG_in[it * nG + g] = this->m_sigmaG[g];
            for (unsigned k = 0; k < K; ++k)
                dump->pi_in[(it * nG + g) * K + k] = this->m_pi(g, k);
        }
        for (unsigned j = 0; j < M; ++j) {
            dump->rand_uz[(size_t(it) * M + j) * 2 + 0] = this->m_randomNumbers[j][0];
            dump->rand_uz[(size_t(it) * M + j) * 2 + 1] = this->m_randomNumbers[j][1];
            dump->order[size_t(it) * M + j] = order[j];
            dump->beta_in[size_t(it) * M + j] = this->m_beta[j];
        }
        std::memcpy(dump->eps_in + size_t(it) * N, this->m_epsilon.data(), sizeof(double) * N);
    }

    void after(unsigned it) override
    {
        if (!dump) return;
        for (unsigned j = 0; j < M; ++j) {
            dump->beta_out[size_t(it) * M + j] = this->m_beta[j];
            dump->comp_out[size_t(it) * M + j] = this->m_components[j];
            dump->acum_out[size_t(it) * M + j] = this->m_acum[j];
        }
        std::memcpy(dump->eps_out + size_t(it) * N, this->m_epsilon.data(), sizeof(double) * N);
        for (unsigned g = 0; g < nG; ++g) {
            dump->betasq_out[it * nG + g] = this->m_betasqnG[g];
            for (unsigned k = 0; k < K; ++k)
                dump->v_out[(it * nG + g) * K + k] = this->m_v(g, k);
        }
        dump->epssum_out[it] = this->m_epsilonSum;
    }

    void post(unsigned it) override
    {
        if (!dump) return;
        dump->sigmaE_post[it] = this->m_sigmaE;
        for (unsigned g = 0; g < nG; ++g) {
            dump->sigmaG_post[it * nG + g] = this->m_sigmaG[g];
            for (unsigned k = 0; k < K; ++k)
                dump->pi_post[(it * nG + g) * K + k] = this->m_pi(g, k);
        }
        const unsigned nF = unsigned(this->m_gamma.size());
        if (dump->gamma_post) for (unsigned f = 0; f < nF; ++f) dump->gamma_post[it * nF + f] = this->m_gamma[f];
        if (dump->eps_post) std::memcpy(dump->eps_post + size_t(it) * N, this->m_epsilon.data(), sizeof(double) * N);
    }

    ChainDump *dump;
    unsigned M, N, nG, K;
};

// The reference's sweep scheduler interface, implemented exactly like Sequential::exec
// (src/sequential.cpp:10-27) but over an in-memory marker list instead of the file-backed
// MarkerCache singleton.
class TapGraph : public AnalysisGraph
{
public:
    explicit TapGraph(const std::vector<ConstMarkerPtr> *markers) : m_markers(markers) {}
    bool isAsynchronous() const override { return false; }
    void exec(Analysis *analysis, unsigned, unsigned, const std::vector<unsigned> &order) override
    {
        auto *tap = dynamic_cast<TapIface *>(analysis);
        if (tap) tap->before(m_iteration, order);
        const auto t0 = std::chrono::steady_clock::now();
        for (unsigned i : order) {
            KernelPtr kernel = analysis->kernelForMarker((*m_markers)[i]);
            analysis->processColumn(kernel);
        }
        m_execSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (tap) tap->after(m_iteration);
        ++m_iteration;
    }
    double execSeconds() const { return m_execSeconds; }
    unsigned iterations() const { return m_iteration; }

private:
    const std::vector<ConstMarkerPtr> *m_markers;
    unsigned m_iteration = 0;
    double m_execSeconds = 0;
};

// The reference's asynchronous scheduler itself (src/parallelgraph.cpp: TBB flow graph, bounded staleness of `analysisTokens`
// markers, all host cores), unmodified, behind the same tap.  It reads its columns from the process-wide MarkerCache singleton
// (src/parallelgraph.cpp:19-24), which normally fills itself from a preprocessed file (src/markercache.cpp:19-45); here the
// markers built in memory are installed into it through the protected member (pointer-to-member of a derived class: legal).
struct CacheAccess : MarkerCache { using MarkerCache::m_markers; };
void installMarkerCache(const std::vector<ConstMarkerPtr> &markers)
{
    (markerCache().get()->*(&CacheAccess::m_markers)) = markers;
}

class TapParallelGraph : public AnalysisGraph
{
public:
    explicit TapParallelGraph(size_t analysisTokens) : m_inner(40, analysisTokens, true) {}      // defaults of src/options.hpp:32-33
    bool isAsynchronous() const override { return true; }
    void exec(Analysis *analysis, unsigned numInds, unsigned numSnps, const std::vector<unsigned> &order) override
    {
        auto *tap = dynamic_cast<TapIface *>(analysis);
        if (tap) tap->before(m_iteration, order);
        const auto t0 = std::chrono::steady_clock::now();
        m_inner.exec(analysis, numInds, numSnps, order);
        m_execSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (tap) tap->after(m_iteration);
        ++m_iteration;
    }
    double execSeconds() const { return m_execSeconds; }
    unsigned iterations() const { return m_iteration; }

private:
    ParallelGraph m_inner;
    unsigned m_iteration = 0;
    double m_execSeconds = 0;
};

// A graph that wraps any AnalysisGraph with the tap (used for the GPU graph of oracle/gpu_graph_ref.cpp)
class TapWrapGraph : public AnalysisGraph
{
public:
    explicit TapWrapGraph(AnalysisGraph *inner) : m_inner(inner) {}
    bool isAsynchronous() const override { return m_inner->isAsynchronous(); }
    void exec(Analysis *analysis, unsigned numInds, unsigned numSnps, const std::vector<unsigned> &order) override
    {
        auto *tap = dynamic_cast<TapIface *>(analysis);
        if (tap) tap->before(m_iteration, order);
        const auto t0 = std::chrono::steady_clock::now();
        m_inner->exec(analysis, numInds, numSnps, order);
        m_execSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (tap) tap->after(m_iteration);
        ++m_iteration;
    }
    double execSeconds() const { return m_execSeconds; }
    unsigned iterations() const { return m_iteration; }

private:
    std::unique_ptr<AnalysisGraph> m_inner;
    unsigned m_iteration = 0;
    double m_execSeconds = 0;
};

} // namespace

#ifdef BRR_REF_WITH_GPU
AnalysisGraph *makeGpuSweepGraph(const Data *data, const Options *opt, bool async, const uint8_t *bedPayload, std::string *err);
#endif

extern "C" {

// ---- tier 1: per-marker numbers straight from the reference's builders and kernels ------------

// Standardised dense column (src/densemarkerbuilder.cpp:23-60). out: N doubles.
int ref_dense_column(const uint8_t *bedCol, uint32_t N, double *out)
{
    std::unique_ptr<MarkerBuilder> b(builderForType(PreprocessDataType::Dense));
    auto m = std::dynamic_pointer_cast<const DenseMarker>(buildMarker(b.get(), 0, bedCol, N));
    if (!m) return 1;
    std::memcpy(out, m->Cx->data(), sizeof(double) * N);
    return 0;
}

// Ragged statistics and index lists (src/raggedsparsemarkerbuilder.cpp:25-55,
// src/sparsemarker.cpp:8-19). stats = {mean, sd, sqrdZ, Zsum}; lists need room for N ints each.
int ref_ragged_marker(const uint8_t *bedCol, uint32_t N, double *stats, int32_t *ones, uint32_t *nOnes,
                      int32_t *twos, uint32_t *nTwos, int32_t *missing, uint32_t *nMissing)
{
    std::unique_ptr<MarkerBuilder> b(builderForType(PreprocessDataType::SparseRagged));
    auto m = std::dynamic_pointer_cast<const RaggedSparseMarker>(buildMarker(b.get(), 0, bedCol, N));
    if (!m) return 1;
    stats[0] = m->mean; stats[1] = m->sd; stats[2] = m->sqrdZ; stats[3] = m->Zsum;
    *nOnes = m->Zones.size(); *nTwos = m->Ztwos.size(); *nMissing = m->Zmissing.size();
    std::copy(m->Zones.begin(), m->Zones.end(), ones);
    std::copy(m->Ztwos.begin(), m->Ztwos.end(), twos);
    std::copy(m->Zmissing.begin(), m->Zmissing.end(), missing);
    return 0;
}

// Eigen-sparse statistics (src/eigensparsemarkerbuilder.cpp:40-78). stats = {mean, sd, sqrdZ, Zsum}.
int ref_eigen_stats(const uint8_t *bedCol, uint32_t N, double *stats)
{
    std::unique_ptr<MarkerBuilder> b(builderForType(PreprocessDataType::SparseEigen));
    auto m = std::dynamic_pointer_cast<const EigenSparseMarker>(buildMarker(b.get(), 0, bedCol, N));
    if (!m) return 1;
    stats[0] = m->mean; stats[1] = m->sd; stats[2] = m->sqrdZ; stats[3] = m->Zsum;
    return 0;
}

// The reference's preprocessed files, written by the reference's own markers (Marker::write / Marker::compress, src/marker.cpp:10-37,
// src/densemarker.cpp, src/raggedsparsemarker.cpp:43-66, src/eigensparsemarker.cpp:38-57) and index writer
// (src/compression.cpp:128-154), as `brr --preprocess` leaves them: test input for the host side's readers.
// layout: 0 dense (.ppbed), 1 ragged, 2 eigen.  Returns 0 on success.
int ref_write_pp_files(const uint8_t *bed, size_t colStride, uint32_t N, uint32_t M, int layout, int compress,
                       const char *markerFile, const char *indexFile)
{
    const auto type = layoutType(layout);
    if (type == PreprocessDataType::None) return 1;
    std::ofstream out(markerFile, std::ios::binary), idx(indexFile, std::ios::binary);
    if (!out || !idx) return 2;
    std::unique_ptr<MarkerBuilder> b(builderForType(type));
    unsigned long pos = 0;
    for (uint32_t j = 0; j < M; ++j) {
        auto m = buildMarker(b.get(), j, bed + size_t(j) * colStride, N);
        if (!m) return 3;
        if (compress) {
            const auto c = m->compress();
            writeCompressedDataWithIndex(c.buffer.get(), c.index.compressedSize, c.index.originalSize, out, idx, pos);
        } else {
            std::ostringstream raw;
            m->write(&raw);
            const std::string bytes = raw.str();
            writeUncompressedDataWithIndex(reinterpret_cast<const unsigned char *>(bytes.data()), bytes.size(), out, idx, pos);
        }
    }
    return (out && idx) ? 0 : 4;
}

// computeNum and the epsilon change for one marker in one layout
// (src/densebayesrkernel.cpp:10-20, src/sparsebayesrkernel.cpp:10-34,
//  src/raggedbayesrkernel.cpp:10-35, src/eigenbayesrkernel.cpp:10-20).
// layout: 0 dense, 1 ragged, 2 eigen.  epsSumInOut: in = current sum(epsilon) (sparse layouts),
// out = the kernel's epsilonSum field after calculateEpsilonChange (= delta of the sum).
int ref_marker_ops(int layout, const uint8_t *bedCol, uint32_t N, const double *eps, double betaOld,
                   double betaNew, double *numOut, double *deltaOut, double *epsSumInOut)
{
    const auto type = layoutType(layout);
    if (type == PreprocessDataType::None) return 1;
    std::unique_ptr<MarkerBuilder> b(builderForType(type));
    ConstMarkerPtr marker = buildMarker(b.get(), 0, bedCol, N);
    if (!marker) return 2;

    Data data;
    data.numInds = N; data.numSnps = 1; data.numGroups = 1; data.G.assign(1, 0);
    data.y = VectorXf::Zero(N);
    Options opt;
    opt.preprocessDataType = type;

    std::unique_ptr<BayesRBase> analysis;
    if (layout == 0) analysis.reset(new DenseBayesRRmz(&data, &opt));
    else analysis.reset(new SparseBayesRRG(&data, &opt));

    std::unique_ptr<Kernel> k = analysis->kernelForMarker(marker);
    auto *bk = dynamic_cast<BayesRKernel *>(k.get());
    if (!bk) return 3;
    VectorXd ones = VectorXd::Ones(N);
    if (auto *sk = dynamic_cast<SparseBayesRKernel *>(bk)) sk->epsilonSum = *epsSumInOut;
    if (auto *ek = dynamic_cast<EigenBayesRKernel *>(bk)) ek->ones = &ones;

    Map<const VectorXd> e(eps, N);
    VectorXd ev = e;
    *numOut = bk->computeNum(ev, betaOld);
    VectorXdPtr d = bk->calculateEpsilonChange(betaOld, betaNew);
    if (d && deltaOut) std::memcpy(deltaOut, d->data(), sizeof(double) * N);
    if (auto *sk = dynamic_cast<SparseBayesRKernel *>(bk)) *epsSumInOut = sk->epsilonSum;
    return 0;
}

// ---- tier 2/3: the reference's own runGibbs with a tap ----------------------------------------

struct ref_chain_args {
    const uint8_t *bed;     // M columns of ceil(N/4) bytes (PLINK SNP-major payload, no magic)
    uint32_t N, M;
    const float *y;         // N phenotypes in the order Data::y would hold them (src/data.cpp:334-365)
    const int32_t *groups;  // M group ids or NULL (all 0, src/data.cpp:130)
    uint32_t nGroups;
    const double *S;        // nGroups x (K-1) row-major (Options::S)
    uint32_t K;
    uint32_t seed;
    uint32_t iterations;
    int32_t layout;         // 0 dense, 1 ragged, 2 eigen
    ChainDump *dump;        // may be NULL (timing only)
    double execSeconds;     // out: total seconds spent inside exec() (the marker loop)
    double totalSeconds;    // out: seconds inside runGibbs
    double buildSeconds;    // out: seconds building markers
    int32_t graphKind;      // 0 = sequential (TapGraph = Sequential::exec), 1 = the reference's ParallelGraph (asyncppbayes),
                            // 2 / 3 = GpuSweepGraph of oracle/gpu_graph_ref.cpp (ppbayes / asyncppbayes) -- libbrr_ref_gpu.so only
    uint32_t analysisTokens; // ParallelGraph: markers in flight between forced synchronisations (Options::analysisTokens, default 20)
    const float *X;          // fixed-effect covariates, N x nF column-major (Data::X, src/data.cpp:369-403), or NULL
    uint32_t nF;             // Options::fixedEffectNumber
};

int ref_run_chain(ref_chain_args *a)
{
    const auto type = layoutType(a->layout);
    if (type == PreprocessDataType::None) return 1;
    const size_t colBytes = (size_t(a->N) + 3) / 4;

    const auto tb0 = std::chrono::steady_clock::now();
    std::vector<ConstMarkerPtr> markers(a->M);
    {
        std::unique_ptr<MarkerBuilder> b(builderForType(type));
        for (uint32_t j = 0; j < a->M; ++j) {
            markers[j] = buildMarker(b.get(), j, a->bed + colBytes * j, a->N);
            if (!markers[j]) return 2;
        }
    }
    a->buildSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tb0).count();

    Data data;
    data.numInds = a->N; data.numSnps = a->M; data.numGroups = a->nGroups;
    data.G.assign(a->M, 0);
    if (a->groups) data.G.assign(a->groups, a->groups + a->M);
    data.y = Map<const VectorXf>(a->y, a->N);

    Options opt;
    opt.preprocessDataType = type;
    opt.chainLength = a->iterations;
    opt.burnin = a->iterations;   // never write a sample row: the CSV writer is not on the path
    opt.thin = 1;
    opt.seed = a->seed;
    opt.mcmcSampleFile = "/dev/null";
    opt.S = Map<const Matrix<double, Dynamic, Dynamic, RowMajor>>(a->S, a->nGroups, a->K - 1);
    opt.numGroups = a->nGroups;
    opt.fixedEffectNumber = a->X ? a->nF : 0;
    if (a->X && a->nF) { data.X = Map<const MatrixXf>(a->X, a->N, a->nF); data.numFixedEffects = a->nF; }
    opt.useMarkerCache = true;

    std::unique_ptr<BayesRBase> analysis;
    if (a->layout == 0)
        analysis.reset(new Tap<DenseBayesRRmz>(&data, &opt, a->dump, a->M, a->N, a->nGroups, a->K));
    else
        analysis.reset(new Tap<SparseBayesRRG>(&data, &opt, a->dump, a->M, a->N, a->nGroups, a->K));

    int rc;
    unsigned iters;
    const auto t0 = std::chrono::steady_clock::now();
    if (a->graphKind == 2 || a->graphKind == 3) {
#ifdef BRR_REF_WITH_GPU
        opt.analysisTokens = a->analysisTokens ? a->analysisTokens : 20;
        std::string err;
        AnalysisGraph *g = makeGpuSweepGraph(&data, &opt, a->graphKind == 3, a->bed, &err);
        if (!g) { std::cerr << "GpuSweepGraph: " << err << std::endl; return 10; }
        TapWrapGraph graph(g);
        rc = analysis->runGibbs(&graph);
        a->execSeconds = graph.execSeconds(); iters = graph.iterations();
#else
        return 11;
#endif
    } else if (a->graphKind == 1) {
        installMarkerCache(markers);
        TapParallelGraph graph(a->analysisTokens ? a->analysisTokens : 20);
        rc = analysis->runGibbs(&graph);
        a->execSeconds = graph.execSeconds(); iters = graph.iterations();
        markerCache()->clear();
    } else {
        TapGraph graph(&markers);
        rc = analysis->runGibbs(&graph);
        a->execSeconds = graph.execSeconds(); iters = graph.iterations();
    }
    a->totalSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (a->dump && iters > 0)
        dynamic_cast<TapIface *>(analysis.get())->post(iters - 1);
    return rc;
}

} // extern "C"
