// Host side of the B200 BayesR engine: see brr_host.hpp for the map to the reference's sources.
#include "brr_host.hpp"
#include "markers.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iomanip>
#include <iostream>
#include <map>
#include <numeric>
#include <sstream>

namespace brrhost {

static int g_samplePrecision = 6;          // the reference writes with the stream's default precision (src/writer.cpp:9)

// ---- Options ----------------------------------------------------------------------------------------------------

static std::vector<std::string> split(const std::string &s, char sep)
{
    std::vector<std::string> out;
    std::string tok;
    std::istringstream in(s);
    while (std::getline(in, tok, sep))
        if (!tok.empty()) out.push_back(tok);
    return out;
}

bool Options::parseVarianceComponents(const std::string &arg, std::vector<std::vector<double>> &S)
{
    std::string text = arg;
    {   // a file name or the components themselves (src/options.cpp:84-93)
        std::ifstream in(arg);
        if (in.is_open()) text.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
    }
    text.erase(std::remove_if(text.begin(), text.end(), [](char c) { return c == '\n' || c == '\r' || c == ' ' || c == '\t'; }), text.end());
    const auto groups = split(text, ';');
    if (groups.empty()) { std::cout << "Failed to parse variance components: " << arg << std::endl; return false; }
    std::vector<std::vector<double>> out;
    for (const auto &g : groups) {
        std::vector<double> row;
        for (const auto &c : split(g, ',')) {
            try { row.push_back(std::stod(c)); }
            catch (const std::exception &) { std::cerr << "Could not parse variance component: " << c << std::endl; return false; }
        }
        if (row.empty() || (!out.empty() && row.size() != out[0].size())) {
            std::cout << "Incorrect number of variance components! Got: " << row.size() << ", expected: " << (out.empty() ? 0 : out[0].size()) << std::endl;
            return false;
        }
        out.push_back(row);
    }
    S = out;
    return true;
}

bool Options::inputOptions(int argc, const char *const argv[])
{
    auto need = [&](int &i) -> const char * {
        if (i + 1 >= argc) { std::cerr << "Error: option " << argv[i] << " needs a value" << std::endl; return nullptr; }
        return argv[++i];
    };
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        const char *v = nullptr;
        if (a == "--inp-file") {                                      // the options come from a file; the rest of the line is ignored (src/options.cpp:98-101)
            if (!(v = need(i))) return false;
            if (!readFile(v)) return false;
            break;
        }
        if (a == "--analysis-type") {
            if (!(v = need(i))) return false;
            const std::string t = v;
            analysisType = t == "ppbayes" ? AnalysisType::PpBayes : t == "asyncppbayes" ? AnalysisType::AsyncPpBayes
                         : t == "preprocess" ? AnalysisType::Preprocess : AnalysisType::Unknown;      // src/options.cpp:7-21
        } else if (a == "--data-file") { if (!(v = need(i))) return false; dataFile = v; }
        else if (a == "--pheno") { if (!(v = need(i))) return false; phenotypeFile = v; }
        else if (a == "--mcmc-samples") { if (!(v = need(i))) return false; mcmcSampleFile = v; }
        else if (a == "--chain-length") { if (!(v = need(i))) return false; chainLength = unsigned(std::atoi(v)); }
        else if (a == "--burn-in") { if (!(v = need(i))) return false; burnin = unsigned(std::atoi(v)); }
        else if (a == "--seed") { if (!(v = need(i))) return false; seed = unsigned(std::atoi(v)); }
        else if (a == "--thin") { if (!(v = need(i))) return false; thin = unsigned(std::atoi(v)); }
        else if (a == "--S") { if (!(v = need(i))) return false; if (!parseVarianceComponents(v, S)) return false; }
        else if (a == "--group") { if (!(v = need(i))) return false; groupFile = v; }
        else if (a == "--fixed_effects") { if (!(v = need(i))) return false; fixedFile = v; }
        else if (a == "--fixedEffectNumber") { if (!(v = need(i))) return false; fixedEffectNumber = unsigned(std::atoi(v)); }
        else if (a == "--analysis-tokens") { if (!(v = need(i))) return false; analysisTokens = unsigned(std::atoi(v)); }
        else if (a == "--sparse-data") {
            if (!(v = need(i))) return false;
            const std::string t = v;
            if (t == "eigen") preprocessDataType = PreprocessDataType::SparseEigen;
            else if (t == "ragged") preprocessDataType = PreprocessDataType::SparseRagged;
            else { std::cerr << "Error: --sparse-data must be eigen or ragged" << std::endl; return false; }
        }
        else if (a == "--gpu") { if (!(v = need(i))) return false; device = std::atoi(v); }
        else if (a == "--block-markers") { if (!(v = need(i))) return false; blockMarkers = unsigned(std::atoi(v)); }
        else if (a == "--sample-precision") { if (!(v = need(i))) return false; g_samplePrecision = std::atoi(v); }
        else if (a == "--no-epsilon") writeEpsilon = false;
        else if (a == "--out") { if (!(v = need(i))) return false; title = v; }
        else if (a == "--iterLog") { if (!(v = need(i))) return false; iterLog = true; iterLogFile = v; }
        // --colLog: the reference logs, per marker, the microseconds of its three steps (src/BayesRBase.cpp:381-389).  A sweep is ONE kernel
        // here, there are no per-marker host timings to log: the file gets its header and one line per sweep (markers, device ms).
        else if (a == "--colLog") { if (!(v = need(i))) return false; colLog = true; colLogFile = v; }
        // (the last group is accepted for command-line compatibility: it configures the reference's TBB graphs, which do not exist here)
        else if (a == "--marker-cache") useMarkerCache = true;       // genotypes from the reference's preprocessed files (markers.hpp)
        else if (a == "--compress") compress = true;
        else if (a == "--preprocess") analysisType = AnalysisType::Preprocess;      // src/options.cpp:111-114
        else if (a == "--thread" || a == "--thread-spawned" || a == "--decompression-concurrency" || a == "--decompression-tokens" ||
                 a == "--analysis-concurrency" || a == "--preprocess-chunks") { if (!need(i)) return false; }
        else { std::cerr << "Error: invalid option \"" << a << "\"." << std::endl; return false; }
    }
    if (seed == 0) seed = static_cast<unsigned>(std::time(nullptr));
    if (thin == 0) thin = 1;
    return true;
}

bool Options::readFile(const std::string &file)
{
    optionFile = file;
    title = file.substr(0, file.rfind('.'));                          // makeTitle, src/options.cpp:326-332
    std::ifstream in(file);
    if (!in) { std::cerr << "Error: can not open the file [" << file << "] to read." << std::endl; return false; }
    std::string key, value;
    while (in >> key >> value) {
        if (key == "phenotypeFile") phenotypeFile = value;
        else if (key == "analysisType")
            analysisType = value == "ppbayes" ? AnalysisType::PpBayes : value == "asyncppbayes" ? AnalysisType::AsyncPpBayes
                         : value == "preprocess" ? AnalysisType::Preprocess : AnalysisType::Unknown;
        else if (key == "mcmcSampleFile") mcmcSampleFile = value;
        else if (key == "chainLength") chainLength = unsigned(std::atoi(value.c_str()));
        else if (key == "burnin") burnin = unsigned(std::atoi(value.c_str()));
        else if (key == "seed") seed = unsigned(std::atoi(value.c_str()));
        else if (key == "thin") thin = unsigned(std::atoi(value.c_str()));
        else if (key == "S") { if (!parseVarianceComponents(value, S)) return false; }
        else if (key == "numThread") { }
        else if (key.substr(0, 2) == "//" || key.substr(0, 1) == "#") continue;
        else { std::cerr << "Error: invalid option " << key << " " << value << std::endl; return false; }
    }
    return true;
}

// ---- Data -------------------------------------------------------------------------------------------------------

bool Data::readFamFile(const std::string &famFile)
{
    std::ifstream in(famFile);
    if (!in) { std::cerr << "Error: can not open the file [" << famFile << "] to read." << std::endl; return false; }
    std::cout << "Reading PLINK FAM file from [" << famFile << "]." << std::endl;
    m_ids.clear();
    std::map<std::string, unsigned> seen;
    std::string fid, pid, dad, mom, sex, phen;
    while (in >> fid >> pid >> dad >> mom >> sex >> phen) {
        const std::string id = fid + ":" + pid;
        if (!seen.emplace(id, unsigned(m_ids.size())).second) {
            std::cerr << "Error: Duplicate individual ID found: \"" << fid << "\t" << pid << "\"." << std::endl;
            return false;
        }
        m_ids.push_back(id);
    }
    numInds = unsigned(m_ids.size());
    std::cout << numInds << " individuals to be included from [" << famFile << "]." << std::endl;
    return numInds > 0;
}

bool Data::readBimFile(const std::string &bimFile)
{
    std::ifstream in(bimFile);
    if (!in) { std::cerr << "Error: can not open the file [" << bimFile << "] to read." << std::endl; return false; }
    std::cout << "Reading PLINK BIM file from [" << bimFile << "]." << std::endl;
    std::string chr, id, genPos, physPos, a1, a2;
    unsigned n = 0;
    while (in >> chr >> id >> genPos >> physPos >> a1 >> a2) ++n;
    numSnps = n;
    G.assign(numSnps, 0);
    numGroups = 1;
    std::cout << numSnps << " SNPs to be included from [" << bimFile << "]." << std::endl;
    return numSnps > 0;
}

bool Data::readBedFile(const std::string &bedFile)
{
    std::ifstream in(bedFile, std::ios::binary);
    if (!in) { std::cerr << "Error: can not open the file [" << bedFile << "] to read." << std::endl; return false; }
    unsigned char magic[3] = {0, 0, 0};
    in.read(reinterpret_cast<char *>(magic), 3);
    if (!in || magic[0] != 0x6c || magic[1] != 0x1b || magic[2] != 0x01) {
        std::cerr << "Error: [" << bedFile << "] is not a SNP-major PLINK .bed file." << std::endl;
        return false;
    }
    colStride = (size_t(numInds) + 3) / 4;
    bed.resize(colStride * numSnps);
    in.read(reinterpret_cast<char *>(bed.data()), std::streamsize(bed.size()));
    if (size_t(in.gcount()) != bed.size()) {
        std::cerr << "Error: [" << bedFile << "] is shorter than " << numSnps << " x " << numInds << " genotypes." << std::endl;
        return false;
    }
    return true;
}

bool Data::readPhenotypeFile(const std::string &phenFile)
{
    std::ifstream in(phenFile);
    if (!in) { std::cerr << "Error: can not open the phenotype file [" << phenFile << "] to read." << std::endl; return false; }
    std::cout << "Reading phenotypes from [" << phenFile << "]." << std::endl;
    std::map<std::string, unsigned> index;
    for (unsigned i = 0; i < m_ids.size(); ++i) index[m_ids[i]] = i;
    y.assign(numInds, 0.0f);
    std::string line;
    unsigned row = 0;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string fid, pid, val;
        if (!(ls >> fid >> pid >> val)) continue;
        // as the reference: the value of the n-th matching, non-NA line goes to y[n] (src/data.cpp:352-363)
        if (index.count(fid + ":" + pid) && val != "NA" && row < numInds) y[row++] = float(std::atof(val.c_str()));
    }
    return row > 0;
}

bool Data::readGroupFile(const std::string &groupFile)
{
    std::ifstream in(groupFile);
    if (!in.is_open()) { std::cout << "Error opening the file" << std::endl; return false; }
    G.clear();
    std::string name;
    int g;
    while (in >> name >> g) G.push_back(g);
    auto tmp = G;
    std::sort(tmp.begin(), tmp.end());
    tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    numGroups = unsigned(tmp.size());
    std::cout << "Groups read from file: " << numGroups << std::endl;
    if (G.size() != numSnps) { std::cerr << "Error: the group file lists " << G.size() << " markers, the .bim " << numSnps << std::endl; return false; }
    for (int v : G)
        if (v < 0 || unsigned(v) >= numGroups) { std::cerr << "Error: group labels must be 0.." << numGroups - 1 << std::endl; return false; }
    return true;
}

bool Data::readCSV(const std::string &filename, int cols)
{
    // src/data.cpp:369-403: one line per individual, comma-separated, atof per cell, values kept as float
    std::ifstream in(filename);
    if (!in) { std::cerr << "Error: can not open the file [" << filename << "] to read." << std::endl; return false; }
    X.assign(size_t(numInds) * size_t(cols), 0.0f);
    std::string line;
    unsigned row = 0;
    while (row < numInds && std::getline(in, line)) {
        int col = 0;
        size_t start = 0;
        for (size_t i = 0; i <= line.size() && col < cols; ++i)
            if (i == line.size() || line[i] == ',') { X[size_t(col) * numInds + row] = float(std::atof(line.c_str() + start)); ++col; start = i + 1; }
        ++row;
    }
    if (row != numInds) { std::cerr << "Error: [" << filename << "] has " << row << " lines for " << numInds << " individuals" << std::endl; return false; }
    numFixedEffects = unsigned(cols);
    return true;
}

std::vector<double> Distributions::dirichlet_rng(const std::vector<double> &alpha)
{
    std::vector<double> r(alpha.size());
    double s = 0;
    for (size_t i = 0; i < alpha.size(); ++i) { r[i] = rgamma(alpha[i], 1.0); s += r[i]; }
    for (auto &x : r) x /= s;
    return r;
}

// ---- device -----------------------------------------------------------------------------------------------------

bool GpuDevice::open(const Data &data, const Options &opt)
{
    brr_gpu_config cfg{};
    cfg.abi_version = BRR_GPU_ABI_VERSION;
    cfg.device = opt.device;
    m_layout = opt.preprocessDataType == PreprocessDataType::Dense ? BRR_GPU_LAYOUT_DENSE2BIT
             : opt.preprocessDataType == PreprocessDataType::SparseEigen ? BRR_GPU_LAYOUT_EIGEN : BRR_GPU_LAYOUT_RAGGED;
    cfg.layout = m_layout;
    cfg.n_inds = cfg.n_inds_total = data.numInds;
    cfg.n_markers = data.numSnps;
    cfg.n_components = unsigned(opt.S[0].size()) + 1;
    cfg.n_groups = data.numGroups;
    cfg.block_markers = opt.blockMarkers;
    cfg.rank = 0; cfg.world = 1;
    auto fail = [&](const char *what) { std::cerr << "Error: " << what << ": " << brr_gpu_last_error() << std::endl; return false; };
    if (brr_gpu_create(&cfg, &m_ctx)) return fail("cannot create the GPU context (there is no CPU fallback)");
    if (opt.useMarkerCache) {
        // --marker-cache: the reference's preprocessed files (.ppbed / .ragged.sparsebed / .eigen.sparsebed + index, --compress for
        // deflated markers) through MarkerCache::populate, then the cache becomes the device store (src/markercache.cpp:19-45)
        if (!markerCache()->populate(opt.dataFile, opt.preprocessDataType, data.numSnps, data.numInds, opt.compress)) return false;
        if (!markerCache()->upload(m_ctx)) return false;
        markerCache()->clear();
    } else {
        // the .bed payload IS the device layout: no .ppbed is written (replaces PreprocessGraph, src/preprocessgraph.cpp:10-73)
        const int rc = m_layout == BRR_GPU_LAYOUT_DENSE2BIT ? brr_gpu_upload_dense2bit(m_ctx, data.bed.data(), data.colStride, 0, data.numSnps)
                                                             : brr_gpu_upload_ragged_from_bed(m_ctx, data.bed.data(), data.colStride);
        if (rc) return fail("genotype upload failed");
        if (brr_gpu_finalize_markers(m_ctx, nullptr)) return fail("marker statistics failed");
    }
    std::vector<int32_t> groups(data.G.begin(), data.G.end());
    std::vector<double> cva;
    for (const auto &row : opt.S) cva.insert(cva.end(), row.begin(), row.end());
    if (brr_gpu_set_model(m_ctx, groups.data(), cva.data())) return fail("model setup failed");
    return true;
}

double GpuBayesRKernel::computeNum(const VectorXd &epsilon, double beta_old)
{
    double num = std::nan("");
    if (epsilon.size() != m_n || brr_gpu_set_state(m_dev->ctx(), epsilon.data(), nullptr, nullptr, nullptr) ||
        brr_gpu_compute_num(m_dev->ctx(), marker, beta_old, epsilonSum, &num))
        std::cerr << "GpuBayesRKernel::computeNum failed: " << brr_gpu_last_error() << std::endl;
    return num;
}

VectorXdPtr GpuBayesRKernel::calculateEpsilonChange(double beta_old, double beta)
{
    // the device applies the change to the resident residual (the one computeNum installed); hand back the difference
    // and restore the residual, so that the call has the reference's value semantics
    auto before = std::make_unique<VectorXd>(m_n), after = std::make_unique<VectorXd>(m_n);
    double dsum = 0;
    if (brr_gpu_get_state(m_dev->ctx(), before->data(), nullptr, nullptr, nullptr) ||
        brr_gpu_apply_delta(m_dev->ctx(), marker, beta_old, beta, &dsum) ||
        brr_gpu_get_state(m_dev->ctx(), after->data(), nullptr, nullptr, nullptr) ||
        brr_gpu_set_state(m_dev->ctx(), before->data(), nullptr, nullptr, nullptr)) {
        std::cerr << "GpuBayesRKernel::calculateEpsilonChange failed: " << brr_gpu_last_error() << std::endl;
        return nullptr;
    }
    epsilonSum = dsum;                                                  // src/sparsebayesrkernel.cpp:29-34
    for (unsigned i = 0; i < m_n; ++i) (*after)[i] -= (*before)[i];
    return after;
}

// ---- the analysis -----------------------------------------------------------------------------------------------

GpuBayesR::GpuBayesR(const Data *data, const Options *opt, GpuDevice *dev)
    : Analysis(data, opt), m_dev(dev), m_dist(opt->seed), m_sparse(opt->preprocessDataType != PreprocessDataType::Dense)
{
}

std::unique_ptr<Kernel> GpuBayesR::kernelForMarker(unsigned marker) const
{
    return std::make_unique<GpuBayesRKernel>(m_dev, marker, m_data->numInds);
}

void GpuBayesR::init(int K, unsigned markerCount, unsigned individualCount)
{
    const unsigned nG = m_data->numGroups;
    m_pi.assign(size_t(nG) * K, 0.5 / K);                               // :43-50
    for (unsigned g = 0; g < nG; ++g) m_pi[size_t(g) * K] = 0.5;
    m_v.assign(size_t(nG) * K, 0.0);
    m_mu = 0.0;
    m_sigmaG.resize(nG);
    for (auto &s : m_sigmaG) s = m_dist.unif_rng();                     // :61-64
    m_betasqnG.assign(nG, 0.0);
    m_beta.assign(markerCount, 0.0);
    m_components.assign(markerCount, 0.0);
    m_acum.assign(markerCount, 0.0);
    // y: float phenotypes, centred and scaled in double (:85-86)
    const unsigned N = individualCount;
    m_epsilon.resize(N);
    double mean = 0;
    for (unsigned i = 0; i < N; ++i) mean += double(m_data->y[i]);
    mean /= N;
    double sq = 0;
    for (unsigned i = 0; i < N; ++i) { m_epsilon[i] = double(m_data->y[i]) - mean; sq += m_epsilon[i] * m_epsilon[i]; }
    const double sc = std::sqrt(sq / double(N - 1));
    sq = 0; m_epsilonSum = 0;
    for (unsigned i = 0; i < N; ++i) { m_epsilon[i] = m_epsilon[i] / sc - m_mu; sq += m_epsilon[i] * m_epsilon[i]; m_epsilonSum += m_epsilon[i]; }
    m_sigmaE = sq / N * 0.5;                                            // :89
    m_randomNumbers.resize(markerCount);
    if (brr_gpu_set_state(m_dev->ctx(), m_epsilon.data(), m_beta.data(), m_components.data(), m_acum.data())) {
        std::cerr << "Error: cannot install the chain state on the device: " << brr_gpu_last_error() << std::endl;
        m_failed = true;
    }
}

void GpuBayesR::prepareForAnylsis()
{
    // the random numbers of an iteration are drawn up front, on this thread, in marker-id order (:101-110)
    for (auto &r : m_randomNumbers) { const double u = m_dist.unif_rng(); r = {u, m_dist.norm_rng(0, 1)}; }
}

bool GpuBayesR::updateMu(double old_mu, double N)
{
    double sum = 0;
    if (brr_gpu_eps_add_scalar(m_dev->ctx(), m_mu, &sum)) return false;                 // epsilon += mu
    if (m_sparse) {
        m_epsilonSum += old_mu * N;                                                     // src/SparseBayesRRG.cpp:109
        m_mu = m_dist.norm_rng(m_epsilonSum / N, m_sigmaE / N);
    } else {
        m_mu = m_dist.norm_rng(sum / N, m_sigmaE / N);                                  // src/DenseBayesRRmz.cpp:39
    }
    if (brr_gpu_eps_add_scalar(m_dev->ctx(), -m_mu, &sum)) return false;                // epsilon -= mu
    if (m_sparse) m_epsilonSum -= m_mu * N;
    return true;
}

int GpuBayesR::runGibbs(AnalysisGraph *analysis)
{
    if (!analysis) {
        std::cout << "Cannot run Gibbs analysis without a flow graph!" << std::endl;
        return 1;
    }
    const unsigned M = m_data->numSnps, N = m_data->numInds, nGroups = m_data->numGroups;
    const int K = int(m_opt->S[0].size()) + 1;
    init(K, M, N);
    if (m_failed) return 1;

    const unsigned nF = m_opt->fixedEffectNumber;
    SampleWriter writer;
    writer.setFixedCount(nF);
    if (!writer.openGroups(m_opt->mcmcSampleFile, M, N, nGroups, m_opt->writeEpsilon)) return 1;
    m_gamma.assign(nF, 0.0);
    std::vector<int> xI(nF);
    std::iota(xI.begin(), xI.end(), 0);

    std::vector<unsigned> markerI(M);
    std::iota(markerI.begin(), markerI.end(), 0);
    std::cout << "Number of groups: " << nGroups << std::endl << "Running Gibbs sampling" << std::endl;
    std::ofstream iterLogger, colLogger;
    if (m_opt->iterLog) {                                             // LogWriter::open, src/logwriter.cpp:6-24
        std::cout << "Opening iteration log file " << m_opt->iterLogFile << std::endl;
        iterLogger.open(m_opt->iterLogFile);
        if (!iterLogger) { std::cerr << "Error: can not open the file [" << m_opt->iterLogFile << "] to write." << std::endl; return 1; }
        iterLogger << "iter,m_0,sigmaG,sigmaE,mu,mu_upd_microS,sigmaG_upd_microS,sigmaE_upd_microS,flowgraph_milliS,iter_d_milliS" << std::endl;
    }
    if (m_opt->colLog) {
        colLogger.open(m_opt->colLogFile);
        if (!colLogger) { std::cerr << "Error: can not open the file [" << m_opt->colLogFile << "] to write." << std::endl; return 1; }
        colLogger << "iter,markers,markers_changing_the_residual,sweep_device_milliS" << std::endl;
    }
    const auto t1 = std::chrono::high_resolution_clock::now();
    double iterSeconds = 0, sweepSeconds = 0;

    for (unsigned iteration = 0; iteration < m_opt->chainLength; ++iteration) {
        const auto startTime = std::chrono::high_resolution_clock::now();
        std::cout << "iteration " << iteration << ": ";
        const double old_mu = m_mu;
        if (!updateMu(old_mu, double(N))) { std::cerr << "Error: " << brr_gpu_last_error() << std::endl; return 1; }
        const auto muTime = std::chrono::high_resolution_clock::now();
        prepareForAnylsis();
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wdeprecated-declarations"
        std::random_shuffle(markerI.begin(), markerI.end());            // as the reference (:195): glibc rand(), never seeded
#pragma GCC diagnostic pop
        m_m0 = 0;
        std::fill(m_v.begin(), m_v.end(), 0.0);
        std::fill(m_betasqnG.begin(), m_betasqnG.end(), 0.0);

        const auto sweepStart = std::chrono::high_resolution_clock::now();
        analysis->exec(this, N, M, markerI);
        const auto sweepEnd = std::chrono::high_resolution_clock::now();
        if (m_failed) return 1;
        if (nF > 0 && !updateFixedEffects(xI)) { std::cerr << "Error: " << brr_gpu_last_error() << std::endl; return 1; }

        const auto sEstartTime = std::chrono::high_resolution_clock::now();
        m_sigmaE = m_dist.inv_scaled_chisq_rng(m_opt->v0E + N, (m_epsilonSqNorm + m_opt->v0E * m_opt->s02E) / (m_opt->v0E + N));   // :238-239
        const auto sEendTime = std::chrono::high_resolution_clock::now();
        const auto sGstartTime = std::chrono::high_resolution_clock::now();
        for (unsigned g = 0; g < nGroups; ++g) {                                                                                    // :243-247
            double rowSum = 0;
            for (int k = 0; k < K; ++k) rowSum += m_v[size_t(g) * K + k];
            m_m0 = rowSum - m_v[size_t(g) * K];
            m_sigmaG[g] = m_dist.inv_scaled_chisq_rng(m_opt->v0G + m_m0, (m_betasqnG[g] * m_m0 + m_opt->v0G * m_opt->s02G) / (m_opt->v0G + m_m0));
            std::vector<double> alpha(K);
            for (int k = 0; k < K; ++k) alpha[k] = m_v[size_t(g) * K + k] + 1.0;
            const auto pi = m_dist.dirichlet_rng(alpha);
            std::copy(pi.begin(), pi.end(), m_pi.begin() + size_t(g) * K);
        }
        const auto sGendTime = std::chrono::high_resolution_clock::now();

        if (iteration >= m_opt->burnin && iteration % m_opt->thin == 0) {
            if (brr_gpu_get_state(m_dev->ctx(), m_opt->writeEpsilon ? m_epsilon.data() : nullptr, m_beta.data(), m_components.data(), m_acum.data())) {
                std::cerr << "Error: " << brr_gpu_last_error() << std::endl;
                return 1;
            }
            writer.write(iteration, m_mu, m_beta, m_sigmaE, m_sigmaG, m_gamma, m_components, m_acum, m_epsilon);
        }
        const auto endTime = std::chrono::high_resolution_clock::now();
        const double it = std::chrono::duration<double>(endTime - startTime).count(), sw = std::chrono::duration<double>(sweepEnd - sweepStart).count();
        std::cout << it << "s (" << sw << "s)" << std::endl;
        iterSeconds += it; sweepSeconds += sw;
        if (m_opt->iterLog) {       // src/BayesRBase.cpp:261-273 (the reference's 10-column line holds ONE sigmaG: with several groups, the last group's m_0 / sigmaG)
            auto us = [](auto a, auto b) { return double(std::chrono::duration_cast<std::chrono::microseconds>(b - a).count()); };
            iterLogger << iteration << "," << m_m0 << "," << m_sigmaG[nGroups - 1] << "," << m_sigmaE << "," << m_mu << "," << us(startTime, muTime) << ","
                       << us(sGstartTime, sGendTime) << "," << us(sEstartTime, sEendTime) << "," << std::floor(sw * 1e3) << "," << std::floor(it * 1e3) << std::endl;
        }
        if (m_opt->colLog) {
            const auto *g = dynamic_cast<GpuSweepGraph *>(analysis);
            colLogger << iteration << "," << M << "," << (g ? g->last().n_nonskipped : 0u) << "," << (g ? g->last().kernel_ms : 0.0f) << std::endl;
        }
    }
    if (brr_gpu_get_state(m_dev->ctx(), m_epsilon.data(), m_beta.data(), m_components.data(), m_acum.data())) {
        std::cerr << "Error: " << brr_gpu_last_error() << std::endl;
        return 1;
    }
    const auto t2 = std::chrono::high_resolution_clock::now();
    std::cout << "duration: " << std::chrono::duration_cast<std::chrono::seconds>(t2 - t1).count() << "s" << std::endl;
    m_meanIteration = iterSeconds / std::max(1u, m_opt->chainLength);
    m_meanSweep = sweepSeconds / std::max(1u, m_opt->chainLength);
    std::cout << "mean iteration duration: " << m_meanIteration << "s" << std::endl
              << "mean flowgraph duration: " << m_meanSweep << "s" << std::endl;
    return 0;
}

// Fixed effects (src/BayesRBase.cpp:214-235): per covariate, in shuffled order, gamma ~ N(num / denom, sigmaE / denom) with
// num = x . (eps + gamma_old x), denom = (N - 1) + sigmaE / sigmaF, then eps += (gamma_old - gamma) x -- the reference's scalar loops
// over the float covariates.  Once per sweep and O(N nF): the residual makes one round trip over PCIe (8 N bytes each way) and the
// loop runs here; ||eps||^2 (and the dense sum) are taken again on the device.
bool GpuBayesR::updateFixedEffects(std::vector<int> &xI)
{
    const unsigned N = m_data->numInds, nF = m_opt->fixedEffectNumber;
    if (m_data->X.size() != size_t(N) * nF) { std::cerr << "Error: --fixedEffectNumber " << nF << " without matching --fixed_effects data" << std::endl; return false; }
    if (brr_gpu_get_state(m_dev->ctx(), m_epsilon.data(), nullptr, nullptr, nullptr)) return false;
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wdeprecated-declarations"
    std::random_shuffle(xI.begin(), xI.end());
#pragma GCC diagnostic pop
    const double dNm1 = double(N - 1), sigE_sigF = m_sigmaE / m_sigmaF;
    for (unsigned i = 0; i < nF; ++i) {
        const float *x = m_data->X.data() + size_t(xI[i]) * N;
        const double gamma_old = m_gamma[xI[i]];
        double num_f = 0.0;
        for (unsigned k = 0; k < N; ++k) num_f += double(x[k]) * (m_epsilon[k] + gamma_old * double(x[k]));
        const double denom_f = dNm1 + sigE_sigF;
        m_gamma[xI[i]] = m_dist.norm_rng(num_f / denom_f, m_sigmaE / denom_f);
        const double d = gamma_old - m_gamma[xI[i]];
        for (unsigned k = 0; k < N; ++k) m_epsilon[k] = m_epsilon[k] + d * double(x[k]);
    }
    m_sigmaF = s02F;
    if (brr_gpu_set_state(m_dev->ctx(), m_epsilon.data(), nullptr, nullptr, nullptr)) return false;
    double sum = 0, sq = 0;
    if (brr_gpu_eps_reduce(m_dev->ctx(), &sum, &sq)) return false;
    m_epsilonSqNorm = sq;                                   // (m_epsilonSum of the sparse layouts is NOT touched: neither does the reference)
    return true;
}

void GpuSweepGraph::exec(Analysis *analysis, unsigned numInds, unsigned numSnps, const std::vector<unsigned> &markerIndices)
{
    auto *a = dynamic_cast<GpuBayesR *>(analysis);
    if (!a || numInds != a->m_data->numInds || numSnps != a->m_data->numSnps || markerIndices.size() > numSnps) {
        std::cerr << "GpuSweepGraph::exec: the analysis is not a GpuBayesR over this data set" << std::endl;
        if (a) a->m_failed = true;
        return;
    }
    static_assert(sizeof(std::array<double, 2>) == 2 * sizeof(double), "(u, z) pairs must be contiguous");
    const int rc = brr_gpu_sweep(m_dev->ctx(), markerIndices.data(), uint32_t(markerIndices.size()), a->m_randomNumbers[0].data(),
                                 a->m_pi.data(), a->m_sigmaG.data(), a->m_sigmaE, a->m_epsilonSum, m_tokens, a->m_v.data(),
                                 a->m_betasqnG.data(), &m_last);
    if (rc) {
        std::cerr << "GpuSweepGraph::exec: " << brr_gpu_last_error() << std::endl;
        a->m_failed = true;
        return;
    }
    a->m_epsilonSqNorm = m_last.eps_sqnorm;
    if (a->m_sparse) a->m_epsilonSum = m_last.eps_sum;                 // the analytically tracked sum (src/SparseBayesRRG.cpp:84-103)
}

// ---- SampleWriter -----------------------------------------------------------------------------------------------

bool SampleWriter::openGroups(const std::string &file, unsigned markers, unsigned individuals, unsigned groups, bool epsilon)
{
    std::cout << "Opening results file " << file << std::endl;
    m_out.open(file);
    if (!m_out) { std::cerr << "Error: can not open the file [" << file << "] to write." << std::endl; return false; }
    m_eps = epsilon;
    m_out << "iteration,mu,";
    for (unsigned i = 0; i < markers; ++i) m_out << "beta[" << (i + 1) << "],";
    m_out << "sigmaE,";
    for (unsigned i = 0; i < groups; ++i) m_out << "sigmaG[" << (i + 1) << "],";
    for (unsigned i = 0; i < m_fixed; ++i) m_out << "gamma[" << (i + 1) << "],";
    for (unsigned i = 0; i < markers; ++i) m_out << "comp[" << (i + 1) << "],";
    for (unsigned i = 0; i < markers; ++i) m_out << "acum[" << (i + 1) << "]" << ((i + 1 < markers || epsilon) ? "," : "");
    if (epsilon)
        for (unsigned i = 0; i < individuals; ++i) m_out << "epsilon[" << (i + 1) << "]" << (i + 1 < individuals ? "," : "");
    m_out << std::endl;
    m_out << std::setprecision(g_samplePrecision);
    return true;
}

void SampleWriter::write(unsigned iteration, double mu, const VectorXd &beta, double sigmaE, const VectorXd &sigmaG, const VectorXd &gamma,
                         const VectorXd &components, const VectorXd &acum, const VectorXd &epsilon)
{
    m_out << iteration << ", " << mu;
    for (double b : beta) m_out << ", " << b;
    m_out << ", " << sigmaE;
    for (double s : sigmaG) m_out << ", " << s;
    for (double g : gamma) m_out << ", " << g;
    for (double c : components) m_out << ", " << c;
    for (double a : acum) m_out << ", " << a;
    if (m_eps)
        for (double e : epsilon) m_out << ", " << e;
    m_out << std::endl;
    m_out.flush();
}

// ---- main -------------------------------------------------------------------------------------------------------

static std::string replaceExtension(const std::string &file, const std::string &ext)
{
    const auto dot = file.rfind('.');
    return (dot == std::string::npos ? file : file.substr(0, dot)) + ext;
}

int brr_main(int argc, const char *const argv[])
{
    Options opt;
    if (!opt.inputOptions(argc, argv)) return 1;
    if (opt.analysisType == AnalysisType::Preprocess) {
        // `brr --preprocess` (src/analysisrunner.cpp:47-78, src/preprocessgraph.cpp:36-66): the .bed through the marker builders into the
        // reference's preprocessed files, for users who keep that workflow (`--marker-cache` reads them back).  Host work only, in memory:
        // not meant for data sets whose dense columns (8 N M bytes) do not fit -- the GPU path itself needs no preprocessing.
        if (opt.dataFile.empty()) { std::cerr << "Error: --data-file <x.bed> is required" << std::endl; return 1; }
        Data data;
        if (!data.readFamFile(replaceExtension(opt.dataFile, ".fam")) || !data.readBimFile(replaceExtension(opt.dataFile, ".bim")) ||
            !data.readBedFile(opt.dataFile))
            return 1;
        std::cout << "Start preprocessing " << opt.dataFile << std::endl;
        MarkerCache cache;
        cache.populateFromBed(data.bed.data(), data.colStride, opt.preprocessDataType, data.numSnps, data.numInds);
        if (!cache.writeFiles(opt.dataFile, opt.preprocessDataType, opt.compress)) return 1;
        std::cout << "Finished preprocessing the bed file: " << ppFileForType(opt.preprocessDataType, opt.dataFile) << std::endl;
        return 0;
    }
    if (opt.analysisType != AnalysisType::PpBayes && opt.analysisType != AnalysisType::AsyncPpBayes) {
        std::cerr << "Error: --analysis-type must be ppbayes, asyncppbayes or preprocess" << std::endl;
        return 1;
    }
    if (opt.dataFile.empty() || opt.phenotypeFile.empty()) {
        std::cerr << "Error: --data-file <x.bed> and --pheno <x.phen> are required" << std::endl;
        return 1;
    }
    Data data;
    if (!data.readFamFile(replaceExtension(opt.dataFile, ".fam")) || !data.readBimFile(replaceExtension(opt.dataFile, ".bim")) ||
        (!opt.useMarkerCache && !data.readBedFile(opt.dataFile)) ||      // (--marker-cache: the preprocessed files replace the .bed)
        !data.readPhenotypeFile(opt.phenotypeFile))
        return 1;
    if (!opt.groupFile.empty() && !data.readGroupFile(opt.groupFile)) return 1;
    if (!opt.fixedFile.empty() && !data.readCSV(opt.fixedFile, int(opt.fixedEffectNumber))) return 1;      // src/analysisrunner.cpp:125-127
    if (opt.S.size() != data.numGroups) {
        std::cerr << "Error: --S has " << opt.S.size() << " group(s) of variance components, the data " << data.numGroups << std::endl;
        return 1;
    }
    GpuDevice dev;
    if (!dev.open(data, opt)) return 1;
    GpuBayesR analysis(&data, &opt, &dev);
    GpuSweepGraph graph(&dev, opt.analysisType == AnalysisType::AsyncPpBayes ? opt.analysisTokens : 0u);
    const int rc = analysis.runGibbs(&graph);
    if (rc) return rc;
    std::cout << "last sweep: " << graph.last().n_nonskipped << " markers changed the residual, " << graph.last().n_blocks << " blocks, "
              << graph.last().kernel_ms << " ms on the device" << std::endl;
    return 0;
}

} // namespace brrhost

extern "C" int brr_host_main(int argc, const char *const argv[])
{
    return brrhost::brr_main(argc, argv);
}
