// Marker / MarkerBuilder / MarkerCache of the host side and the reference's preprocessed file formats (see markers.hpp).
#include "markers.hpp"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

namespace brrhost {

// ---- file names, index, compression ---------------------------------------------------------------------------

static std::string fileWithoutExtension(const std::string &file)
{
    const auto dot = file.find_last_of('.');
    const auto slash = file.find_last_of('/');
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return file;
    return file.substr(0, dot);
}

std::string ppFileForType(PreprocessDataType type, const std::string &dataFile)
{
    const std::string base = fileWithoutExtension(dataFile);
    switch (type) {
    case PreprocessDataType::Dense: return base + ".ppbed";
    case PreprocessDataType::SparseEigen: return base + ".eigen.sparsebed";
    case PreprocessDataType::SparseRagged: return base + ".ragged.sparsebed";
    }
    return {};
}

std::string ppIndexFileForType(PreprocessDataType type, const std::string &dataFile)
{
    return ppFileForType(type, dataFile) + "index";
}

bool readIndexFile(const std::string &file, unsigned numSnps, std::vector<IndexEntry> &index)
{
    std::ifstream in(file, std::ios::binary);
    if (!in) { std::cerr << "Error: cannot open the index file [" << file << "]" << std::endl; return false; }
    index.assign(numSnps, IndexEntry{});
    static_assert(sizeof(IndexEntry) == 3 * sizeof(unsigned long), "three unsigned longs per marker");
    in.read(reinterpret_cast<char *>(index.data()), std::streamsize(sizeof(IndexEntry)) * numSnps);
    if (in.gcount() != std::streamsize(sizeof(IndexEntry)) * numSnps) {
        std::cerr << "Error: the index file [" << file << "] holds fewer than " << numSnps << " entries" << std::endl;
        return false;
    }
    return true;
}

bool extractData(const unsigned char *src, unsigned long srcSize, unsigned char *dst, unsigned long dstSize)
{
    z_stream strm{};
    if (inflateInit(&strm) != Z_OK) { std::cerr << "Error: inflateInit failed" << std::endl; return false; }
    strm.next_in = const_cast<unsigned char *>(src); strm.avail_in = static_cast<unsigned>(srcSize);
    strm.next_out = dst; strm.avail_out = static_cast<unsigned>(dstSize);
    const int ret = inflate(&strm, Z_FINISH);
    inflateEnd(&strm);
    if (ret != Z_STREAM_END) { std::cerr << "Error: failed to inflate a compressed marker" << std::endl; return false; }
    return true;
}

unsigned long compressData(const unsigned char *src, unsigned long srcSize, std::vector<unsigned char> &dst)
{
    uLongf cap = compressBound(srcSize);
    dst.resize(cap);
    if (compress2(dst.data(), &cap, src, srcSize, Z_DEFAULT_COMPRESSION) != Z_OK) { dst.clear(); return 0; }
    dst.resize(cap);
    return cap;
}

// ---- markers --------------------------------------------------------------------------------------------------------

bool Marker::decompress(const unsigned char *data, const IndexEntry &index)
{
    std::string raw(index.originalSize, '\0');
    if (!extractData(data + index.pos, index.compressedSize, reinterpret_cast<unsigned char *>(&raw[0]), index.originalSize)) return false;
    std::istringstream in(raw, std::ios::binary);
    read(&in);
    return isValid();
}

std::streamsize DenseMarker::size() const { return std::streamsize(sizeof(double)) * numInds; }
void DenseMarker::read(std::istream *in)
{
    if (in->fail()) { std::cerr << "Error: unable to read DenseMarker!" << std::endl; return; }
    Cx.assign(numInds, 0.0);
    in->read(reinterpret_cast<char *>(Cx.data()), size());
    if (in->gcount() != size()) Cx.clear();
}
void DenseMarker::write(std::ostream *out) const
{
    if (out->fail()) { std::cerr << "Error: unable to write DenseMarker!" << std::endl; return; }
    out->write(reinterpret_cast<const char *>(Cx.data()), size());
}

void SparseMarker::updateStatistics(unsigned allele1, unsigned allele2)
{
    if (allele1 == 0 && allele2 == 1) return;                  // missing genotype
    if (allele1 == 1 || allele2 == 1) {
        const double value = allele1 + allele2;
        mean += value; sqrdZ += value * value; Zsum += value;
    }
}
std::streamsize SparseMarker::size() const { return std::streamsize(sizeof(double)) * 4; }
void SparseMarker::read(std::istream *in)
{
    if (in->fail()) { std::cerr << "Error: unable to read SparseMarker statistics!" << std::endl; return; }
    double d[4] = {0, 0, 0, 0};
    in->read(reinterpret_cast<char *>(d), sizeof(d));
    mean = d[0]; sd = d[1]; sqrdZ = d[2]; Zsum = d[3];
}
void SparseMarker::write(std::ostream *out) const
{
    if (out->fail()) { std::cerr << "Error: unable to write SparseMarker statistics!" << std::endl; return; }
    const double d[4] = {mean, sd, sqrdZ, Zsum};
    out->write(reinterpret_cast<const char *>(d), sizeof(d));
}

std::streamsize RaggedSparseMarker::size() const
{
    return SparseMarker::size() + std::streamsize(3 * sizeof(size_t) + sizeof(int) * (Zones.size() + Ztwos.size() + Zmissing.size()));
}
void RaggedSparseMarker::read(std::istream *in)
{
    SparseMarker::read(in);
    for (IndexVector *v : {&Zones, &Ztwos, &Zmissing}) {
        size_t n = 0;
        in->read(reinterpret_cast<char *>(&n), sizeof(n));
        if (in->fail() || n > size_t(numInds)) { std::cerr << "Error: unable to read RaggedSparseMarker!" << std::endl; v->clear(); in->setstate(std::ios::failbit); return; }
        v->assign(n, 0);
        if (n) in->read(reinterpret_cast<char *>(v->data()), std::streamsize(n * sizeof(int)));
    }
}
void RaggedSparseMarker::write(std::ostream *out) const
{
    SparseMarker::write(out);
    for (const IndexVector *v : {&Zones, &Ztwos, &Zmissing}) {
        const size_t n = v->size();
        out->write(reinterpret_cast<const char *>(&n), sizeof(n));
        if (n) out->write(reinterpret_cast<const char *>(v->data()), std::streamsize(n * sizeof(int)));
    }
}
bool RaggedSparseMarker::isValid() const { return !(Zones.empty() && Ztwos.empty()); }      // src/raggedsparsemarker.cpp:68-75

std::streamsize EigenSparseMarker::size() const
{
    return SparseMarker::size() + std::streamsize(sizeof(long) + (sizeof(double) + sizeof(int)) * values.size());
}
void EigenSparseMarker::read(std::istream *in)
{
    SparseMarker::read(in);
    long nnz = 0;                                              // Eigen::Index
    in->read(reinterpret_cast<char *>(&nnz), sizeof(nnz));
    if (in->fail() || nnz < 0 || nnz > long(numInds)) { std::cerr << "Error: unable to read EigenSparseMarker!" << std::endl; in->setstate(std::ios::failbit); return; }
    values.assign(size_t(nnz), 0.0); indices.assign(size_t(nnz), 0);
    if (nnz) {
        in->read(reinterpret_cast<char *>(values.data()), std::streamsize(sizeof(double)) * nnz);
        in->read(reinterpret_cast<char *>(indices.data()), std::streamsize(sizeof(int)) * nnz);
    }
}
void EigenSparseMarker::write(std::ostream *out) const
{
    SparseMarker::write(out);
    const long nnz = long(values.size());
    out->write(reinterpret_cast<const char *>(&nnz), sizeof(nnz));
    if (nnz) {
        out->write(reinterpret_cast<const char *>(values.data()), std::streamsize(sizeof(double)) * nnz);
        out->write(reinterpret_cast<const char *>(indices.data()), std::streamsize(sizeof(int)) * nnz);
    }
}
bool EigenSparseMarker::isValid() const { return !values.empty() && values.size() == indices.size(); }

// ---- builders -------------------------------------------------------------------------------------------------------

void MarkerBuilder::initialise(unsigned snp, unsigned numInds)
{
    m_snp = snp; m_numInds = numInds; m_missingIndices.clear(); m_sum = 0;
    m_marker = makeMarker();
    m_marker->i = snp; m_marker->numInds = numInds;
}
bool MarkerBuilder::read(const std::string &file, const IndexEntry &index)
{
    std::ifstream in(file, std::ios::binary);
    if (!in) { std::cerr << "Error: cannot open [" << file << "]" << std::endl; return false; }
    in.seekg(std::streamoff(index.pos));
    m_marker->read(&in);
    return !in.fail() && m_marker->isValid();
}
bool MarkerBuilder::decompress(const unsigned char *data, const IndexEntry &index) { return m_marker->decompress(data, index); }
std::unique_ptr<Marker> MarkerBuilder::build() { return std::move(m_marker); }

void DenseMarkerBuilder::processAllele(unsigned individual, unsigned allele1, unsigned allele2)
{
    auto *m = static_cast<DenseMarker *>(m_marker.get());
    if (m->Cx.size() != size_t(m_numInds)) m->Cx.assign(size_t(m_numInds), 0.0);
    if (allele1 == 0 && allele2 == 1) m_missingIndices.push_back(individual);        // filled with the mean in endColumn
    else { const double value = allele1 + allele2; m->Cx[individual] = value; m_sum += value; }
}
void DenseMarkerBuilder::endColumn()
{
    // src/densemarkerbuilder.cpp:43-60: impute the mean of the observed genotypes, centre by the column mean, scale to
    // squared norm N - 1
    auto *m = static_cast<DenseMarker *>(m_marker.get());
    if (m->Cx.size() != size_t(m_numInds)) m->Cx.assign(size_t(m_numInds), 0.0);
    const double mean = m_sum / (m_numInds - double(m_missingIndices.size()));
    for (unsigned idx : m_missingIndices) m->Cx[idx] = mean;
    double s = 0;
    for (double x : m->Cx) s += x;
    const double colMean = s / m_numInds;
    double sqn = 0;
    for (double &x : m->Cx) { x -= colMean; sqn += x * x; }
    const double sigma = 1.0 / std::sqrt(sqn / (m_numInds - 1.0));
    for (double &x : m->Cx) x *= sigma;
}

void RaggedSparseMarkerBuilder::processAllele(unsigned individual, unsigned allele1, unsigned allele2)
{
    auto *m = static_cast<RaggedSparseMarker *>(m_marker.get());
    m->updateStatistics(allele1, allele2);
    if (allele1 == 0 && allele2 == 1) m->Zmissing.push_back(int(individual));
    else if (allele1 == 1 && allele2 == 0) m->Zones.push_back(int(individual));
    else if (allele1 == 1 && allele2 == 1) m->Ztwos.push_back(int(individual));
}
void RaggedSparseMarkerBuilder::endColumn()
{
    auto *m = static_cast<RaggedSparseMarker *>(m_marker.get());       // src/raggedsparsemarkerbuilder.cpp:43-55
    m->mean /= m_numInds;
    const double mean = m->mean;
    m->sd = std::sqrt((m->sqrdZ - 2.0 * mean * m->Zsum + m_numInds * mean * mean) / (m_numInds - 1.0));
}

void EigenSparseMarkerBuilder::processAllele(unsigned individual, unsigned allele1, unsigned allele2)
{
    auto *m = static_cast<EigenSparseMarker *>(m_marker.get());
    if (m->values.empty() && m_entries.size() && m_entries.front().first >= int(individual)) m_entries.clear();   // (a new column)
    m->updateStatistics(allele1, allele2);
    if (allele1 == 0 && allele2 == 1) m_entries.emplace_back(int(individual), -1.0);
    else if (allele1 == 1 || allele2 == 1) m_entries.emplace_back(int(individual), double(allele1 + allele2));
}
void EigenSparseMarkerBuilder::endColumn()
{
    auto *m = static_cast<EigenSparseMarker *>(m_marker.get());        // src/eigensparsemarkerbuilder.cpp:40-78
    double nMissing = 0;
    for (const auto &e : m_entries) if (e.second <= 0) nMissing += 1;
    m->mean /= m_numInds;
    const double mean = m->mean;
    m->sqrdZ += (mean * mean) * nMissing;
    m->Zsum += mean * nMissing;
    m->sd = std::sqrt((m->sqrdZ - 2.0 * mean * m->Zsum + m_numInds * mean * mean) / (m_numInds - 1.0));
    m->values.clear(); m->indices.clear();
    for (const auto &e : m_entries) { m->indices.push_back(e.first); m->values.push_back(e.second <= 0 ? mean : e.second); }
    m_entries.clear();
}

MarkerBuilder *builderForType(PreprocessDataType type)
{
    switch (type) {
    case PreprocessDataType::Dense: return new DenseMarkerBuilder;
    case PreprocessDataType::SparseEigen: return new EigenSparseMarkerBuilder;
    case PreprocessDataType::SparseRagged: return new RaggedSparseMarkerBuilder;
    }
    return nullptr;
}

// ---- the cache ----------------------------------------------------------------------------------------------------------

const std::shared_ptr<MarkerCache> markerCache()
{
    static const auto cache = std::make_shared<MarkerCache>();
    return cache;
}

bool MarkerCache::populate(const std::string &dataFile, PreprocessDataType type, unsigned numSnps, unsigned numInds, bool compress)
{
    clear();
    std::vector<IndexEntry> index;
    if (!readIndexFile(ppIndexFileForType(type, dataFile), numSnps, index)) return false;
    const std::string ppFile = ppFileForType(type, dataFile);
    std::ifstream in(ppFile, std::ios::binary);
    if (!in) { std::cerr << "Error: cannot open the preprocessed file [" << ppFile << "]" << std::endl; return false; }
    std::unique_ptr<MarkerBuilder> builder{builderForType(type)};
    m_markers.resize(numSnps);
    std::vector<unsigned char> buf;
    for (unsigned snp = 0; snp < numSnps; ++snp) {
        builder->initialise(snp, numInds);
        const IndexEntry &ie = index[snp];
        bool ok;
        if (compress) {                                         // the reference maps the file; here one marker's bytes at a time
            buf.resize(ie.compressedSize);
            in.seekg(std::streamoff(ie.pos));
            in.read(reinterpret_cast<char *>(buf.data()), std::streamsize(ie.compressedSize));
            IndexEntry local = ie; local.pos = 0;
            ok = in.gcount() == std::streamsize(ie.compressedSize) && builder->decompress(buf.data(), local);
        } else {
            in.seekg(std::streamoff(ie.pos));
            std::unique_ptr<Marker> m = builder->build();
            m->read(&in);
            ok = !in.fail() && m->isValid();
            m_markers[snp] = std::move(m);
            if (!ok) { std::cerr << "Error: marker " << snp << " of [" << ppFile << "] cannot be read" << std::endl; clear(); return false; }
            continue;
        }
        if (!ok) { std::cerr << "Error: marker " << snp << " of [" << ppFile << "] cannot be read" << std::endl; clear(); return false; }
        m_markers[snp] = builder->build();
    }
    std::cout << "Cached " << numSnps << " markers" << std::endl;
    return true;
}

void MarkerCache::populateFromBed(const uint8_t *bed, size_t colStride, PreprocessDataType type, unsigned numSnps, unsigned numInds)
{
    clear();
    std::unique_ptr<MarkerBuilder> builder{builderForType(type)};
    m_markers.resize(numSnps);
    for (unsigned snp = 0; snp < numSnps; ++snp) {
        builder->initialise(snp, numInds);
        const uint8_t *col = bed + size_t(snp) * colStride;
        for (unsigned i = 0; i < numInds; ++i) {                // bit pairs LSB first; allele = !bit (src/preprocessgraph.cpp:44-53)
            const unsigned two = (col[i >> 2] >> (2 * (i & 3))) & 3u;
            builder->processAllele(i, !(two & 1u), !((two >> 1) & 1u));
        }
        builder->endColumn();
        m_markers[snp] = builder->build();
    }
}

bool MarkerCache::writeFiles(const std::string &dataFile, PreprocessDataType type, bool compress) const
{
    std::ofstream out(ppFileForType(type, dataFile), std::ios::binary), idx(ppIndexFileForType(type, dataFile), std::ios::binary);
    if (!out || !idx) return false;
    unsigned long pos = 0;
    std::vector<unsigned char> packed;
    for (const auto &m : m_markers) {
        std::ostringstream raw(std::ios::binary);
        m->write(&raw);
        const std::string bytes = raw.str();
        IndexEntry ie; ie.pos = pos; ie.originalSize = bytes.size(); ie.compressedSize = bytes.size();
        if (compress) {
            ie.compressedSize = compressData(reinterpret_cast<const unsigned char *>(bytes.data()), bytes.size(), packed);
            if (!ie.compressedSize) return false;
            out.write(reinterpret_cast<const char *>(packed.data()), std::streamsize(ie.compressedSize));
        } else {
            out.write(bytes.data(), std::streamsize(bytes.size()));
        }
        idx.write(reinterpret_cast<const char *>(&ie), sizeof(ie));
        pos += ie.compressedSize;
    }
    return bool(out) && bool(idx);
}

ConstMarkerPtr MarkerCache::marker(unsigned i) const
{
    if (i >= m_markers.size()) {
        std::cerr << "Requesting marker out of bounds: " << i << ". " << m_markers.size() << " cached markers." << std::endl;
        return {};
    }
    return m_markers[i];
}

bool denseColumnToCodes(const std::vector<double> &Cx, uint8_t *bedColumn, double lut[4])
{
    // distinct values of the column, with a tolerance far below the spacing 1 / sd of the genotype values
    const size_t N = Cx.size();
    double lo = Cx[0], hi = Cx[0];
    for (double x : Cx) { lo = std::min(lo, x); hi = std::max(hi, x); }
    const double tol = 1e-9 * std::max(1.0, hi - lo);
    std::vector<double> vals;
    for (double x : Cx) {
        bool seen = false;
        for (double v : vals) if (std::fabs(v - x) <= tol) { seen = true; break; }
        if (!seen) { vals.push_back(x); if (vals.size() > 4) return false; }
    }
    std::sort(vals.begin(), vals.end());
    // Genotype values are x_g = (g - mean) / sd: spacing d = 1 / sd, and sum of squares = N - 1 fixes sd once the classes are known.
    // Try every assignment of the distinct values to (a subset of) the classes 0, 1, 2, missing that is increasing in g with equal
    // spacing; the imputed missing value is the observed mean, i.e. it lies strictly inside (x_0, x_2) and is not a genotype value.
    auto count = [&](double v) { size_t c = 0; for (double x : Cx) if (std::fabs(x - v) <= tol) ++c; return c; };
    const int nv = int(vals.size());
    int best[4] = {-1, -1, -1, -1};                             // index into vals for class 0, 1, 2, missing
    bool found = false;
    for (int miss = -1; miss < nv && !found; ++miss) {          // which value (if any) is the imputed missing one
        std::vector<int> g;
        for (int k = 0; k < nv; ++k) if (k != miss) g.push_back(k);
        if (g.empty() || g.size() > 3) continue;
        // the genotype values present, in increasing order, are classes c0 < c1 < ...: try the possible class sets
        for (int c0 = 0; c0 <= 2 && !found; ++c0)
            for (int step1 = 1; step1 <= 2 && !found; ++step1)
                for (int step2 = 1; step2 <= 2 && !found; ++step2) {
                    int cls[3] = {c0, c0 + step1, c0 + step1 + step2};
                    bool ok = true;
                    for (size_t k = 0; k < g.size(); ++k) ok = ok && cls[k] <= 2;
                    if (!ok) continue;
                    if (g.size() >= 2) {
                        // spacing per class step must be constant and must match the scale: sum over all x^2 = N - 1 holds by
                        // construction, so check the mean instead: sum_g n_g x_g + n_miss x_miss = 0 is also by construction.
                        // What identifies the classes is x_g = (g - mean) / sd with mean = sum_g g n_g / (N - n_miss):
                        const double d = (vals[g[1]] - vals[g[0]]) / double(cls[1] - cls[0]);
                        if (g.size() == 3 && std::fabs((vals[g[2]] - vals[g[1]]) - d * double(cls[2] - cls[1])) > 1e-6 * d) continue;
                        double sumg = 0, nobs = 0;
                        for (size_t k = 0; k < g.size(); ++k) { const double n = double(count(vals[g[k]])); sumg += cls[k] * n; nobs += n; }
                        const double mean = sumg / nobs;
                        // value of class c: x_c = (c - meanAll) * d where meanAll is the column mean after imputation = mean
                        const double x0 = (double(cls[0]) - mean) * d;
                        if (std::fabs(x0 - vals[g[0]]) > 1e-6 * std::max(1.0, std::fabs(d))) continue;
                        if (miss >= 0 && std::fabs(vals[miss]) > 1e-6 * std::max(1.0, std::fabs(d))) continue;      // imputed mean -> 0 after centring
                    } else {
                        if (miss >= 0) continue;                // one genotype class only: a constant column cannot be standardised anyway
                    }
                    for (int &b : best) b = -1;
                    for (size_t k = 0; k < g.size(); ++k) best[cls[k]] = g[k];
                    best[3] = miss;
                    found = true;
                }
    }
    if (!found) return false;
    for (int c = 0; c < 4; ++c) lut[c] = best[c] >= 0 ? vals[best[c]] : 0.0;
    // absent classes get the value their genotype would have (keeps the table affine; never addressed by a code)
    {
        int a = -1, b = -1;
        for (int c = 0; c < 3; ++c) if (best[c] >= 0) { if (a < 0) a = c; else if (b < 0) b = c; }
        if (a >= 0 && b >= 0) {
            const double d = (lut[b] - lut[a]) / double(b - a);
            for (int c = 0; c < 3; ++c) if (best[c] < 0) lut[c] = lut[a] + d * double(c - a);
        }
    }
    static const uint8_t kBedCode[4] = {3u, 2u, 0u, 1u};         // allele count 0, 1, 2, missing -> PLINK bit pair (11, 10, 00, 01)
    std::memset(bedColumn, 0, (N + 3) / 4);
    for (size_t i = 0; i < N; ++i) {
        int cls = -1;
        for (int c = 0; c < 4; ++c) if (best[c] >= 0 && std::fabs(Cx[i] - vals[best[c]]) <= tol) { cls = c; break; }
        if (cls < 0) return false;
        bedColumn[i >> 2] |= uint8_t(kBedCode[cls] << (2 * (i & 3)));
    }
    return true;
}

bool MarkerCache::upload(brr_gpu_ctx *ctx) const
{
    if (m_markers.empty() || !ctx) { std::cerr << "Error: MarkerCache::upload without markers / context" << std::endl; return false; }
    const unsigned M = unsigned(m_markers.size()), N = m_markers[0]->numInds;
    auto fail = [&](const char *what) { std::cerr << "Error: " << what << ": " << brr_gpu_last_error() << std::endl; return false; };
    if (dynamic_cast<const DenseMarker *>(m_markers[0].get())) {
        const size_t stride = (size_t(N) + 3) / 4;
        std::vector<uint8_t> bed(stride * M);
        std::vector<double> lut(size_t(M) * 4);
        for (unsigned j = 0; j < M; ++j) {
            const auto *m = dynamic_cast<const DenseMarker *>(m_markers[j].get());
            if (!m || !denseColumnToCodes(m->Cx, bed.data() + stride * j, &lut[size_t(j) * 4])) {
                std::cerr << "Error: dense marker " << j << " is not a standardised biallelic column" << std::endl;
                return false;
            }
        }
        if (brr_gpu_upload_dense2bit(ctx, bed.data(), stride, 0, M)) return fail("dense upload failed");
        if (brr_gpu_finalize_markers(ctx, nullptr)) return fail("marker statistics failed");
        if (brr_gpu_set_dense_lut(ctx, lut.data())) return fail("value table upload failed");    // the file's own values, bit for bit
        return true;
    }
    std::vector<uint64_t> offsets(size_t(3) * M + 1, 0);
    std::vector<int32_t> idx;
    std::vector<double> stats(size_t(M) * 4);
    for (unsigned j = 0; j < M; ++j) {
        const auto *s = dynamic_cast<const SparseMarker *>(m_markers[j].get());
        if (!s) { std::cerr << "Error: marker " << j << " is not a sparse marker" << std::endl; return false; }
        stats[size_t(j) * 4 + 0] = s->mean; stats[size_t(j) * 4 + 1] = s->sd; stats[size_t(j) * 4 + 2] = s->sqrdZ; stats[size_t(j) * 4 + 3] = s->Zsum;
        std::vector<int> ones, twos, missing;
        if (const auto *r = dynamic_cast<const RaggedSparseMarker *>(s)) { ones = r->Zones; twos = r->Ztwos; missing = r->Zmissing; }
        else if (const auto *e = dynamic_cast<const EigenSparseMarker *>(s)) {
            for (size_t k = 0; k < e->values.size(); ++k) {      // values are exactly 1, 2 or the column mean (a missing genotype)
                if (e->values[k] == 1.0 && e->mean != 1.0) ones.push_back(e->indices[k]);
                else if (e->values[k] == 2.0 && e->mean != 2.0) twos.push_back(e->indices[k]);
                else missing.push_back(e->indices[k]);
            }
        }
        for (const auto *v : {&ones, &twos, &missing}) {
            idx.insert(idx.end(), v->begin(), v->end());
            offsets[size_t(3) * j + size_t(v == &ones ? 1 : v == &twos ? 2 : 3)] = idx.size();
        }
    }
    if (brr_gpu_upload_ragged(ctx, offsets.data(), idx.data(), stats.data())) return fail("sparse upload failed");
    if (brr_gpu_finalize_markers(ctx, nullptr)) return fail("marker statistics failed");
    return true;
}

} // namespace brrhost

// ---- C entry points for the bindings / tests (ctypes): the process-wide cache ----------------------------------------
using namespace brrhost;

static PreprocessDataType typeOf(int layout)
{
    return layout == 0 ? PreprocessDataType::Dense : layout == 2 ? PreprocessDataType::SparseEigen : PreprocessDataType::SparseRagged;
}

extern "C" int brr_host_cache_from_files(const char *dataFile, int layout, unsigned numSnps, unsigned numInds, int compress)
{
    return markerCache()->populate(dataFile, typeOf(layout), numSnps, numInds, compress != 0) ? 0 : 1;
}

extern "C" int brr_host_cache_from_bed(const uint8_t *bed, size_t colStride, int layout, unsigned numSnps, unsigned numInds)
{
    markerCache()->populateFromBed(bed, colStride, typeOf(layout), numSnps, numInds);
    return 0;
}

extern "C" int brr_host_cache_write(const char *dataFile, int layout, int compress)
{
    return markerCache()->writeFiles(dataFile, typeOf(layout), compress != 0) ? 0 : 1;
}

// One cached marker: dense -> `dense` gets numInds doubles; sparse -> stats4 = {mean, sd, sqrdZ, Zsum} and the three index lists
// (ragged) or their reconstruction from the sparse vector (eigen; `values` then also gets the nnz values in file order).
extern "C" int brr_host_cache_marker(unsigned i, double *dense, double *stats4, int *ones, unsigned *nOnes, int *twos, unsigned *nTwos,
                                     int *missing, unsigned *nMissing, double *values, int *indices, unsigned *nnz)
{
    const ConstMarkerPtr m = markerCache()->marker(i);
    if (!m) return 1;
    if (const auto *d = dynamic_cast<const DenseMarker *>(m.get())) {
        if (dense) std::copy(d->Cx.begin(), d->Cx.end(), dense);
        return 0;
    }
    const auto *s = dynamic_cast<const SparseMarker *>(m.get());
    if (!s) return 2;
    if (stats4) { stats4[0] = s->mean; stats4[1] = s->sd; stats4[2] = s->sqrdZ; stats4[3] = s->Zsum; }
    if (const auto *r = dynamic_cast<const RaggedSparseMarker *>(s)) {
        if (ones) std::copy(r->Zones.begin(), r->Zones.end(), ones);
        if (twos) std::copy(r->Ztwos.begin(), r->Ztwos.end(), twos);
        if (missing) std::copy(r->Zmissing.begin(), r->Zmissing.end(), missing);
        if (nOnes) *nOnes = unsigned(r->Zones.size());
        if (nTwos) *nTwos = unsigned(r->Ztwos.size());
        if (nMissing) *nMissing = unsigned(r->Zmissing.size());
    } else if (const auto *e = dynamic_cast<const EigenSparseMarker *>(s)) {
        if (values) std::copy(e->values.begin(), e->values.end(), values);
        if (indices) std::copy(e->indices.begin(), e->indices.end(), indices);
        if (nnz) *nnz = unsigned(e->values.size());
    }
    return 0;
}

extern "C" int brr_host_cache_upload(brr_gpu_ctx *ctx) { return markerCache()->upload(ctx) ? 0 : 1; }

extern "C" int brr_host_dense_column_to_codes(const double *Cx, unsigned numInds, uint8_t *bedColumn, double *lut4)
{
    return denseColumnToCodes(std::vector<double>(Cx, Cx + numInds), bedColumn, lut4) ? 0 : 1;
}
