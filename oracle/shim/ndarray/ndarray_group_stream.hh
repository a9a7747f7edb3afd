// TEST INFRASTRUCTURE ONLY (oracle build shim) -- minimal stand-in for the
// un-vendored `ndarray` dependency (hguo/ndarray @ 7eda716c, script/download.lua:158).
// Only the type surface that src/IO/MPASOReader.*, src/Core/MPASOGrid.* and
// src/Core/MPASOSolution.* name is provided; no file is ever read through it.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>
namespace ftk {
struct ndarray_base {
    virtual ~ndarray_base() = default;
    virtual int type() const { return 0; }
    static std::string dtype2str(int) { return "stub"; }
};
template <class T>
struct ndarray : public ndarray_base {
    std::vector<T> data;
    const std::vector<T>& std_vector() const { return data; }
    std::vector<T>& std_vector() { return data; }
};
struct ndarray_group {
    std::map<std::string, std::shared_ptr<ndarray_base>> arrays;
    bool has(const std::string& k) const { return arrays.count(k) != 0; }
    std::shared_ptr<ndarray_base> get(const std::string& k) const
    {
        auto it = arrays.find(k);
        return it == arrays.end() ? nullptr : it->second;
    }
};
struct substream {
    std::vector<std::string> filenames;
    std::vector<int> first_timestep_per_file;
};
struct stream {
    std::string path_prefix;
    std::vector<std::shared_ptr<substream>> substreams;
    void parse_yaml(const std::string&) {}
    void set_path_prefix(const std::string& p) { path_prefix = p; }
    std::shared_ptr<ndarray_group> read_static() { return std::make_shared<ndarray_group>(); }
    std::shared_ptr<ndarray_group> read(int) { return std::make_shared<ndarray_group>(); }
    int total_timesteps() const { return 0; }
};
inline void ndarray_finalize() {}
} // namespace ftk
