// Column supply of the host side: the reference's Marker / MarkerBuilder / MarkerCache interfaces and its preprocessed
// file formats, so that data prepared with `brr --preprocess` (.ppbed / .ragged.sparsebed / .eigen.sparsebed + their index
// files, plain or zlib-compressed) feeds the device store through the same calls the reference makes.
//
//   IndexEntry                 src/common.h:57-61           {pos, compressedSize, originalSize}, 3 x unsigned long per marker
//   Marker                     src/marker.h:19-35           i, numInds, size / read / write / isValid, decompress
//   DenseMarker                src/densemarker.h:7-20       N standardised doubles               (.ppbed)
//   SparseMarker               src/sparsemarker.h:6-18      mean, sd, sqrdZ, Zsum + updateStatistics (src/sparsemarker.cpp:8-19)
//   RaggedSparseMarker         src/raggedsparsemarker.h:7-23   Zones / Ztwos / Zmissing          (.ragged.sparsebed)
//   EigenSparseMarker          src/eigensparsemarker.h:6-16    sparse vector Zg (values {1,2,mean}) (.eigen.sparsebed)
//   MarkerBuilder (+ Dense / RaggedSparse / EigenSparse)   src/markerbuilder.h:11-42, src/densemarkerbuilder.cpp:23-60,
//                              src/raggedsparsemarkerbuilder.cpp:25-55, src/eigensparsemarkerbuilder.cpp:40-78
//   MarkerCache                src/markercache.cpp:19-57    populate(data, options), marker(i)
//
// The cache is the hand-over point to the device ("MarkerCache::populate-time device upload"): MarkerCache::upload
// turns the cached markers into the device store -- ragged / eigen markers as index lists + statistics, dense markers as 2-bit
// codes + the per-marker value table recovered from the standardised column (it has at most four distinct values).
#ifndef BRR_HOST_MARKERS_HPP
#define BRR_HOST_MARKERS_HPP

#include "brr_host.hpp"

#include <iosfwd>
#include <memory>
#include <string>
#include <vector>

namespace brrhost {

struct IndexEntry {                                // src/common.h:57-61
    unsigned long pos = 0, compressedSize = 0, originalSize = 0;
};

std::string ppFileForType(PreprocessDataType type, const std::string &dataFile);        // src/common.cpp: .ppbed / .eigen.sparsebed / .ragged.sparsebed
std::string ppIndexFileForType(PreprocessDataType type, const std::string &dataFile);   // the same + "index"
bool readIndexFile(const std::string &file, unsigned numSnps, std::vector<IndexEntry> &index);    // src/data.cpp:56-61
// zlib inflate of one marker's bytes (src/compression.cpp:100-126); false + std::cerr on failure
bool extractData(const unsigned char *src, unsigned long srcSize, unsigned char *dst, unsigned long dstSize);
unsigned long compressData(const unsigned char *src, unsigned long srcSize, std::vector<unsigned char> &dst);   // deflate, level -1

struct Marker {
    virtual ~Marker() = default;
    unsigned i = 0;
    unsigned numInds = 0;
    virtual std::streamsize size() const = 0;
    virtual void read(std::istream *in) = 0;
    virtual void write(std::ostream *out) const = 0;
    virtual bool isValid() const = 0;
    virtual bool decompress(const unsigned char *data, const IndexEntry &index);   // inflate + read
};
using MarkerPtr = std::shared_ptr<Marker>;
using ConstMarkerPtr = std::shared_ptr<const Marker>;

struct DenseMarker : Marker {
    std::vector<double> Cx;                        // the standardised column (the reference maps it over a raw buffer)
    std::streamsize size() const override;
    void read(std::istream *in) override;
    void write(std::ostream *out) const override;
    bool isValid() const override { return Cx.size() == numInds && numInds > 0; }
};

struct SparseMarker : Marker {
    double mean = 0, sd = 0, sqrdZ = 0, Zsum = 0;
    virtual void updateStatistics(unsigned allele1, unsigned allele2);
    std::streamsize size() const override;
    void read(std::istream *in) override;
    void write(std::ostream *out) const override;
};

struct RaggedSparseMarker : SparseMarker {
    using IndexVector = std::vector<int>;
    IndexVector Zones, Ztwos, Zmissing;
    std::streamsize size() const override;
    void read(std::istream *in) override;
    void write(std::ostream *out) const override;
    bool isValid() const override;
};

struct EigenSparseMarker : SparseMarker {
    std::vector<double> values;                    // Zg: 1, 2 or the column mean (missing genotypes)
    std::vector<int> indices;                      // ascending individual ids
    std::streamsize size() const override;
    void read(std::istream *in) override;
    void write(std::ostream *out) const override;
    bool isValid() const override;
};

class MarkerBuilder {
public:
    virtual ~MarkerBuilder() = default;
    virtual void initialise(unsigned snp, unsigned numInds);
    virtual void processAllele(unsigned individual, unsigned allele1, unsigned allele2) = 0;
    virtual void endColumn() = 0;
    virtual bool read(const std::string &file, const IndexEntry &index);            // the marker's bytes straight from the file
    virtual bool decompress(const unsigned char *data, const IndexEntry &index);    // from a memory-mapped compressed file
    virtual std::unique_ptr<Marker> build();       // hands the marker over (the builder is empty afterwards)
protected:
    virtual std::unique_ptr<Marker> makeMarker() const = 0;
    std::unique_ptr<Marker> m_marker;
    unsigned m_snp = 0;
    double m_numInds = 0;
    std::vector<unsigned> m_missingIndices;
    double m_sum = 0;
};

class DenseMarkerBuilder : public MarkerBuilder {
public:
    void processAllele(unsigned individual, unsigned allele1, unsigned allele2) override;
    void endColumn() override;
protected:
    std::unique_ptr<Marker> makeMarker() const override { return std::make_unique<DenseMarker>(); }
};

class RaggedSparseMarkerBuilder : public MarkerBuilder {
public:
    void processAllele(unsigned individual, unsigned allele1, unsigned allele2) override;
    void endColumn() override;
protected:
    std::unique_ptr<Marker> makeMarker() const override { return std::make_unique<RaggedSparseMarker>(); }
};

class EigenSparseMarkerBuilder : public MarkerBuilder {
public:
    void processAllele(unsigned individual, unsigned allele1, unsigned allele2) override;
    void endColumn() override;
protected:
    std::unique_ptr<Marker> makeMarker() const override { return std::make_unique<EigenSparseMarker>(); }
private:
    std::vector<std::pair<int, double>> m_entries; // (individual, value), missing entries valued at endColumn
};

MarkerBuilder *builderForType(PreprocessDataType type);       // caller owns (src/common.cpp)

class MarkerCache {
public:
    void clear() { m_markers.clear(); }
    // From the preprocessed files of `dataFile` (index file + marker file; `compress` = the file holds deflated markers):
    // MarkerCache::populate of src/markercache.cpp:19-45.  false + std::cerr when a file is missing or short.
    bool populate(const std::string &dataFile, PreprocessDataType type, unsigned numSnps, unsigned numInds, bool compress);
    // From a PLINK payload through the builders (what `brr --preprocess` computes, src/preprocessgraph.cpp:36-66), without files.
    void populateFromBed(const uint8_t *bed, size_t colStride, PreprocessDataType type, unsigned numSnps, unsigned numInds);
    // Writes the cached markers in the reference's file formats (marker file + index file); returns false on an I/O error.
    bool writeFiles(const std::string &dataFile, PreprocessDataType type, bool compress) const;
    ConstMarkerPtr marker(unsigned i) const;       // {} + std::cerr when out of bounds
    size_t size() const { return m_markers.size(); }
    // The cached markers become the device store of `ctx` (layout taken from the marker type).  Dense markers: the column's
    // distinct values give the 2-bit codes and the value table (brr_gpu_upload_dense2bit + brr_gpu_set_dense_lut); sparse
    // markers: index lists + (mean, sd, sqrdZ, Zsum) (brr_gpu_upload_ragged).  false + std::cerr on failure.
    bool upload(brr_gpu_ctx *ctx) const;
private:
    std::vector<ConstMarkerPtr> m_markers;
};

const std::shared_ptr<MarkerCache> markerCache();   // the process-wide cache (src/markercache.cpp:8-12)

// Dense column -> PLINK-style 2-bit codes (00 -> 2, 10 -> 1, 11 -> 0, 01 -> missing) + the four values (class 0, 1, 2, missing).
// The standardised column of a biallelic marker takes the values (g - mean) / sd for g = 0, 1, 2 -- equally spaced, increasing --
// and (about) 0 for imputed missing genotypes; false when the column does not have that shape.
bool denseColumnToCodes(const std::vector<double> &Cx, uint8_t *bedColumn, double lut[4]);

} // namespace brrhost

#endif
