// mpas_io.hpp -- MPAS-Ocean file ingestion for the drop-in API: a YAML-subset "stream" description
// (the schema of the reference's mpas.yaml / tutorial/test.yaml, which the reference parses with the
// third-party ndarray `ftk::stream`) and a dependency-free NetCDF-3 reader (classic CDF-1, 64-bit
// offset CDF-2 and CDF-5 headers).  netCDF-4 / HDF5 files are NOT supported (no HDF5 in this image):
// opening one fails with a message saying so.
//
// Reference being replaced: src/IO/MPASOReader.cpp:128-245 (readGridData / readSolData) and the
// ftk::stream calls it makes (parse_yaml, read_static, read(index), substreams[i]->filenames,
// first_timestep_per_file).  Wire format: SURVEY.md Appendix C.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace MOPS {
namespace io {

// ---- YAML subset --------------------------------------------------------------------------------
struct YamlNode {
    enum Kind { Null, Scalar, Map, List } kind = Null;
    std::string scalar;
    std::vector<std::pair<std::string, YamlNode>> map; // insertion order kept
    std::vector<YamlNode> list;
    const YamlNode* get(const std::string& key) const;
    std::string str(const std::string& key, const std::string& dflt = "") const;
    bool boolean(const std::string& key, bool dflt) const;
};
// block-style maps and lists, scalars (plain / single- / double-quoted), '#' comments.  Throws
// std::runtime_error on anything else (flow collections, anchors, multi-line scalars).
YamlNode parse_yaml_subset(const std::string& text);

// ---- NetCDF-3 -------------------------------------------------------------------------------------
struct NcVar {
    std::string name;
    std::vector<int> dimids;
    int type = 0;        // 1 byte, 2 char, 3 short, 4 int, 5 float, 6 double, 7..11 CDF-5 unsigned / int64
    uint64_t vsize = 0;  // bytes per record (record vars) or of the whole variable
    uint64_t begin = 0;
    bool is_record = false;
    uint64_t elems_per_record = 0; // product of the non-record dimensions
};

class NcFile {
public:
    explicit NcFile(const std::string& path); // throws std::runtime_error (message names the reason)
    ~NcFile();
    const NcVar* var(const std::string& name) const;
    uint64_t num_records() const { return numrecs_; }
    uint64_t dim_len(const std::string& name) const; // 0 if absent (record dim: num_records)
    // whole variable (record == -1; record variables: all records) or one record of a record variable
    bool read_double(const std::string& name, int64_t record, std::vector<double>& out) const;
    bool read_int(const std::string& name, int64_t record, std::vector<int64_t>& out) const;
    bool read_char(const std::string& name, int64_t record, std::vector<char>& out) const;
    const std::string& path() const { return path_; }

private:
    template <class T> bool read_as(const std::string& name, int64_t record, std::vector<T>& out) const;
    std::string path_;
    void* fp_ = nullptr;
    uint64_t file_size_ = 0;
    int version_ = 1;
    uint64_t numrecs_ = 0, recsize_ = 0;
    std::vector<std::pair<std::string, uint64_t>> dims_;
    int recdim_ = -1;
    std::vector<NcVar> vars_;
};

// ---- stream (mesh substream + time-varying data substream) --------------------------------------
struct StreamVar {
    std::string name;                        // the name the reader asks for
    std::vector<std::string> possible_names; // aliases tried in order in the file
    bool optional = false;
};
struct Substream {
    std::string name, format;
    bool is_static = false;
    std::vector<std::string> filenames;      // glob-expanded, sorted, relative to path_prefix as written
    std::vector<std::string> paths;          // absolute
    std::vector<int> first_timestep_per_file;
    std::vector<StreamVar> vars;
};
class Stream {
public:
    void parse_yaml(const std::string& yaml_path); // throws std::runtime_error
    std::string path_prefix;
    std::vector<std::shared_ptr<Substream>> substreams;
    // name -> file variable resolution through possible_names; nullptr if absent
    std::shared_ptr<NcFile> open_static() const;
    // global record index -> (file, local record)
    std::shared_ptr<NcFile> open_record(int index, int64_t& local_record) const;
    std::string resolve(const Substream& sub, const NcFile& f, const std::string& wanted) const;
    int total_timesteps() const;
};

} // namespace io
} // namespace MOPS
