// Host side of the B200 BayesR engine, above the C ABI (include/brr_gpu.h).
//
// It mirrors the reference's interfaces for the ppbayes path -- same names, argument meaning and error
// behaviour -- so that the GPU graph drops in where the reference's TBB graphs sit (reference paths are
// relative to the reference repository):
//
//   Options            src/options.hpp:20-112       (the brr command line: --data-file --pheno --analysis-type
//                                                     --chain-length --burn-in --thin --mcmc-samples --S --group
//                                                     --seed --sparse-data --analysis-tokens)
//   Data               src/data.hpp                  (PLINK bed/bim/fam, phenotype and group files)
//   Distributions      src/distributions_boost.hpp   (Boost.Random is replaced by the std:: distributions, exactly
//                                                     as oracle/shim does for the reference build used as the oracle)
//   Kernel/BayesRKernel src/kernel.h:8-14, src/bayesrkernel.h:6-16     -> GpuBayesRKernel
//   Analysis           src/analysis.h:25-52          (runGibbs + kernelForMarker)
//   AnalysisGraph      src/analysisgraph.hpp:9-26    (exec(Analysis*, numInds, numSnps, markerIndices))
//                                                     -> GpuSweepGraph (replaces Sequential / LimitSequenceGraph /
//                                                        ParallelGraph: src/sequential.cpp, src/limitsequencegraph.cpp,
//                                                        src/parallelgraph.cpp)
//   BayesRBase         src/BayesRBase.hpp, .cpp:39-285 (init, prepareForAnylsis, runGibbs)  -> GpuBayesR
//   SampleWriter       src/samplewriter.cpp:82-118   (openGroups / write)
//
// There is no CPU implementation of the marker loop here: GpuSweepGraph::exec hands the sweep to the device and
// fails loudly when libbrr_gpu.so cannot create a context.  Errors: no exceptions cross the interfaces; like the
// reference, failures are reported on std::cerr and by a non-zero return value (src/main.cpp:27-40).
#ifndef BRR_HOST_HPP
#define BRR_HOST_HPP

#include "brr_gpu.h"

#include <array>
#include <cstdint>
#include <fstream>
#include <memory>
#include <random>
#include <string>
#include <vector>

namespace brrhost {

using VectorXd = std::vector<double>;
using VectorXdPtr = std::unique_ptr<VectorXd>;

enum class AnalysisType { Unknown, PpBayes, AsyncPpBayes, Preprocess };
enum class PreprocessDataType { Dense, SparseEigen, SparseRagged };

struct Options {                                   // src/options.hpp:20-112 (the fields of the ppbayes path, same defaults)
    unsigned chainLength = 10000, burnin = 5000, thin = 5;
    unsigned seed = 0;                             // 0 = std::time(0), as the reference's default
    unsigned analysisTokens = 20;
    AnalysisType analysisType = AnalysisType::Unknown;
    PreprocessDataType preprocessDataType = PreprocessDataType::Dense;
    std::string dataFile, phenotypeFile, groupFile, mcmcSampleFile = "bayesOutput.csv";
    std::string fixedFile;                         // --fixed_effects <csv>: N rows of comma-separated covariates (src/options.cpp:162-165)
    unsigned fixedEffectNumber = 0;                // --fixedEffectNumber n   (src/options.cpp:167-170)
    std::vector<std::vector<double>> S{{0.01, 0.001, 0.0001}};     // [groups][K-1]
    double v0E = 0.0001, s02E = 0.0001, v0G = 0.0001, s02G = 0.0001;
    int device = 0;                                // --gpu n
    unsigned blockMarkers = 0;                     // --block-markers n (0 = engine default)
    bool writeEpsilon = true;                      // --no-epsilon drops the N residual columns from the sample file
    bool iterLog = false, colLog = false;          // --iterLog <file>: per-iteration log (src/logwriter.cpp); --colLog <file>: per-column timings (see brr_host.cpp)
    std::string iterLogFile, colLogFile, title, optionFile;
    bool readFile(const std::string &file);        // --inp-file: `key value` pairs (src/options.cpp:280-324)
    bool useMarkerCache = false;                   // --marker-cache: genotypes from the preprocessed files of --data-file (markers.hpp)
    bool compress = false;                         // --compress: those files hold deflated markers
    bool inputOptions(int argc, const char *const argv[]);   // false + message on std::cerr when the line is unusable
    static bool parseVarianceComponents(const std::string &arg, std::vector<std::vector<double>> &S);   // src/options.cpp:20-93
};

struct Data {                                      // src/data.hpp (what the path reads)
    unsigned numInds = 0, numSnps = 0, numGroups = 1;
    std::vector<float> y;                          // VectorXf y (src/data.hpp:92): phenotypes pass through float
    std::vector<int> G;                            // group of every marker (src/data.cpp:130, :424-450)
    std::vector<float> X;                          // MatrixXf X (src/data.hpp:89): fixed-effect covariates, column-major N x numFixedEffects
    unsigned numFixedEffects = 0;
    std::vector<uint8_t> bed;                      // PLINK SNP-major payload without the 3-byte magic
    size_t colStride = 0;
    bool readFamFile(const std::string &famFile);              // src/data.cpp:85-107
    bool readBimFile(const std::string &bimFile);              // src/data.cpp:109-133
    bool readBedFile(const std::string &bedFile);              // payload only: the .bed IS the device layout
    bool readPhenotypeFile(const std::string &phenFile);       // src/data.cpp:334-367
    bool readGroupFile(const std::string &groupFile);          // src/data.cpp:424-453
    bool readCSV(const std::string &filename, int cols);       // src/data.cpp:369-403 (fixed-effect covariates)
private:
    std::vector<std::string> m_ids;                // fid:pid in fam order
};

class Distributions {                              // src/distributions_boost.cpp:16-73 over oracle/shim's mapping
public:
    explicit Distributions(unsigned seed) : m_rng(seed) {}
    double unif_rng() { return std::uniform_real_distribution<double>(0, 1)(m_rng); }
    double norm_rng(double mean, double sigma2) { return std::normal_distribution<double>(mean, std::sqrt(sigma2))(m_rng); }
    double rgamma(double shape, double scale) { return std::gamma_distribution<double>(shape, scale)(m_rng); }
    double inv_gamma_rng(double shape, double scale) { return 1.0 / rgamma(shape, 1.0 / scale); }
    double inv_scaled_chisq_rng(double dof, double scale) { return inv_gamma_rng(0.5 * dof, 0.5 * dof * scale); }
    std::vector<double> dirichlet_rng(const std::vector<double> &alpha);
private:
    std::mt19937 m_rng;
};

// ---- the device, shared by the graph and the kernels ----------------------------------------------------------
class GpuDevice {
public:
    GpuDevice() = default;
    ~GpuDevice() { brr_gpu_destroy(m_ctx); }
    GpuDevice(const GpuDevice &) = delete;
    GpuDevice &operator=(const GpuDevice &) = delete;
    bool open(const Data &data, const Options &opt);      // context + genotype upload + model; false + std::cerr on failure
    brr_gpu_ctx *ctx() const { return m_ctx; }
    int layout() const { return m_layout; }
private:
    brr_gpu_ctx *m_ctx = nullptr;
    int m_layout = BRR_GPU_LAYOUT_DENSE2BIT;
};

// ---- src/kernel.h, src/bayesrkernel.h -------------------------------------------------------------------------
struct Kernel {
    explicit Kernel(unsigned marker) : marker(marker) {}
    virtual ~Kernel() = default;
    unsigned marker;
};

struct BayesRKernel : Kernel {
    using Kernel::Kernel;
    virtual double computeNum(const VectorXd &epsilon, double beta_old) = 0;
    virtual VectorXdPtr calculateEpsilonChange(double beta_old, double beta) = 0;
    double epsilonSum = 0;                         // in/out for the sparse layouts (src/sparsebayesrkernel.h:11)
};

// Marker-level kernel on the device (parity / compatibility: epsilon goes host -> device on every call).
struct GpuBayesRKernel : BayesRKernel {
    GpuBayesRKernel(GpuDevice *dev, unsigned marker, unsigned numInds) : BayesRKernel(marker), m_dev(dev), m_n(numInds) {}
    double computeNum(const VectorXd &epsilon, double beta_old) override;
    VectorXdPtr calculateEpsilonChange(double beta_old, double beta) override;     // nullptr on a device error
private:
    GpuDevice *m_dev;
    unsigned m_n;
};

// ---- src/analysis.h, src/analysisgraph.hpp --------------------------------------------------------------------
class AnalysisGraph;

class Analysis {
public:
    Analysis(const Data *data, const Options *opt) : m_data(data), m_opt(opt) {}
    virtual ~Analysis() = default;
    virtual std::unique_ptr<Kernel> kernelForMarker(unsigned marker) const = 0;
    virtual int runGibbs(AnalysisGraph *analysis) = 0;
protected:
    const Data *m_data;
    const Options *m_opt;
};

class AnalysisGraph {
public:
    virtual ~AnalysisGraph() = default;
    virtual bool isAsynchronous() const = 0;
    virtual void exec(Analysis *analysis, unsigned numInds, unsigned numSnps, const std::vector<unsigned> &markerIndices) = 0;
};

// ---- BayesRBase with its state on the device (src/BayesRBase.hpp:43-112) -----------------------------------------
class GpuSweepGraph;

class GpuBayesR : public Analysis {
    friend class GpuSweepGraph;                    // as LimitSequenceGraph / ParallelGraph are friends of the reference's analyses
public:
    GpuBayesR(const Data *data, const Options *opt, GpuDevice *dev);
    std::unique_ptr<Kernel> kernelForMarker(unsigned marker) const override;
    int runGibbs(AnalysisGraph *analysis) override;              // src/BayesRBase.cpp:130-285
    // chain state after runGibbs (what SampleWriter saw last)
    const VectorXd &beta() const { return m_beta; }
    const VectorXd &components() const { return m_components; }
    const VectorXd &acum() const { return m_acum; }
    double mu() const { return m_mu; }
    double sigmaE() const { return m_sigmaE; }
    const VectorXd &sigmaG() const { return m_sigmaG; }
    const VectorXd &gamma() const { return m_gamma; }
    double meanIterationSeconds() const { return m_meanIteration; }
    double meanSweepSeconds() const { return m_meanSweep; }
    bool failed() const { return m_failed; }
protected:
    void init(int K, unsigned markerCount, unsigned individualCount);   // :39-99
    void prepareForAnylsis();                                           // :101-110 (the reference's spelling)
    bool updateMu(double old_mu, double N);                             // src/DenseBayesRRmz.cpp:36-42, src/SparseBayesRRG.cpp:106-113
    bool updateFixedEffects(std::vector<int> &xI);                      // src/BayesRBase.cpp:214-235
    VectorXd m_gamma;                                                   // fixed effects
    const double s02F = 1.0;
    double m_sigmaF = 1.0;                                              // (the reference reads it uninitialised in iteration 0, :218; it is s02F from then on)
    GpuDevice *m_dev;
    Distributions m_dist;
    bool m_sparse = false, m_failed = false;
    std::vector<std::array<double, 2>> m_randomNumbers;                 // (u, z) per marker id
    std::vector<double> m_pi, m_v, m_betasqnG, m_sigmaG;               // [nG*K], [nG*K], [nG], [nG]
    double m_mu = 0, m_sigmaE = 0, m_epsilonSum = 0, m_epsilonSqNorm = 0, m_m0 = 0;
    VectorXd m_beta, m_components, m_acum, m_epsilon;                   // fetched from the device when a sample is written
    double m_meanIteration = 0, m_meanSweep = 0;
};

// One exec = one device sweep (brr_gpu_sweep): the marker loop of src/BayesRBase.cpp:287-392 (processColumn) or, with
// tokens > 0, the bounded-staleness variant of :394-523 under src/parallelgraph.cpp.
class GpuSweepGraph : public AnalysisGraph {
public:
    GpuSweepGraph(GpuDevice *dev, unsigned asyncTokens) : m_dev(dev), m_tokens(asyncTokens) {}
    bool isAsynchronous() const override { return m_tokens != 0; }
    void exec(Analysis *analysis, unsigned numInds, unsigned numSnps, const std::vector<unsigned> &markerIndices) override;
    const brr_gpu_sweep_out &last() const { return m_last; }
private:
    GpuDevice *m_dev;
    unsigned m_tokens;
    brr_gpu_sweep_out m_last{};
};

class SampleWriter {                               // src/samplewriter.cpp:82-118
public:
    void setFixedCount(unsigned n) { m_fixed = n; }                    // gamma[1..n] columns (src/samplewriter.cpp:82-118)
    bool openGroups(const std::string &file, unsigned markers, unsigned individuals, unsigned groups, bool epsilon);
    void write(unsigned iteration, double mu, const VectorXd &beta, double sigmaE, const VectorXd &sigmaG, const VectorXd &gamma,
               const VectorXd &components, const VectorXd &acum, const VectorXd &epsilon);
private:
    std::ofstream m_out;
    bool m_eps = true;
    unsigned m_fixed = 0;
};

// src/main.cpp + src/analysisrunner.cpp:123-152,246-276 for --analysis-type ppbayes / asyncppbayes
int brr_main(int argc, const char *const argv[]);

} // namespace brrhost

extern "C" int brr_host_main(int argc, const char *const argv[]);

#endif
