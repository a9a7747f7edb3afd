// mpas_io.cpp -- see mpas_io.hpp.
#include "mpas_io.hpp"

#include <glob.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace MOPS {
namespace io {

// =================================================================================================
// YAML subset
// =================================================================================================
const YamlNode* YamlNode::get(const std::string& key) const
{
    for (auto& kv : map)
        if (kv.first == key) return &kv.second;
    return nullptr;
}
std::string YamlNode::str(const std::string& key, const std::string& dflt) const
{
    const YamlNode* n = get(key);
    return (n && n->kind == Scalar) ? n->scalar : dflt;
}
bool YamlNode::boolean(const std::string& key, bool dflt) const
{
    const YamlNode* n = get(key);
    if (!n || n->kind != Scalar) return dflt;
    std::string s = n->scalar;
    std::transform(s.begin(), s.end(), s.begin(), ::tolower);
    if (s == "true" || s == "yes" || s == "on" || s == "1") return true;
    if (s == "false" || s == "no" || s == "off" || s == "0") return false;
    return dflt;
}

namespace {
struct Line {
    int indent;
    std::string text; // without indentation / comment / trailing blanks
    int number;
};

std::string strip_comment(const std::string& s)
{
    bool sq = false, dq = false;
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[i];
        if (c == '\'' && !dq) sq = !sq;
        else if (c == '"' && !sq) dq = !dq;
        else if (c == '#' && !sq && !dq && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return s.substr(0, i);
    }
    return s;
}
std::string trim(const std::string& s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
std::string unquote(const std::string& s)
{
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
    return s;
}
// position of the ':' that separates key and value (not inside quotes; followed by blank or end)
size_t key_colon(const std::string& s)
{
    bool sq = false, dq = false;
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[i];
        if (c == '\'' && !dq) sq = !sq;
        else if (c == '"' && !sq) dq = !dq;
        else if (c == ':' && !sq && !dq && (i + 1 == s.size() || s[i + 1] == ' ' || s[i + 1] == '\t')) return i;
    }
    return std::string::npos;
}

struct Parser {
    std::vector<Line> lines;
    size_t pos = 0;

    [[noreturn]] void fail(const Line& l, const char* why) const
    {
        throw std::runtime_error("yaml line " + std::to_string(l.number) + ": " + why + " ('" + l.text + "')");
    }

    YamlNode parse_block(int indent)
    {
        YamlNode node;
        if (pos >= lines.size()) return node;
        if (lines[pos].text.rfind("- ", 0) == 0 || lines[pos].text == "-") {
            node.kind = YamlNode::List;
            while (pos < lines.size() && lines[pos].indent == indent && (lines[pos].text.rfind("- ", 0) == 0 || lines[pos].text == "-")) {
                Line l = lines[pos];
                const std::string rest = trim(l.text.substr(1));
                if (rest.empty()) {
                    ++pos;
                    if (pos < lines.size() && lines[pos].indent > indent) node.list.push_back(parse_block(lines[pos].indent));
                    else node.list.emplace_back();
                } else if (key_colon(rest) != std::string::npos) {
                    // "- key: value": a map whose first entry sits on the dash line; its other entries are
                    // indented to the column of that key
                    const int child_indent = indent + static_cast<int>(l.text.size() - trim(l.text.substr(1)).size());
                    lines[pos].indent = child_indent;
                    lines[pos].text = rest;
                    node.list.push_back(parse_block(child_indent));
                } else {
                    YamlNode s;
                    s.kind = YamlNode::Scalar;
                    s.scalar = unquote(rest);
                    node.list.push_back(s);
                    ++pos;
                }
            }
            return node;
        }
        node.kind = YamlNode::Map;
        while (pos < lines.size() && lines[pos].indent == indent) {
            const Line l = lines[pos];
            if (l.text.rfind("- ", 0) == 0) break;
            const size_t c = key_colon(l.text);
            if (c == std::string::npos) fail(l, "expected 'key: value'");
            const std::string key = unquote(trim(l.text.substr(0, c)));
            const std::string val = trim(l.text.substr(c + 1));
            ++pos;
            YamlNode child;
            if (!val.empty()) {
                if (val[0] == '[' || val[0] == '{' || val[0] == '|' || val[0] == '>' || val[0] == '&' || val[0] == '*')
                    fail(l, "flow collections / block scalars / anchors are outside the supported subset");
                child.kind = YamlNode::Scalar;
                child.scalar = unquote(val);
            } else if (pos < lines.size() && (lines[pos].indent > indent ||
                                              (lines[pos].indent == indent && lines[pos].text.rfind("- ", 0) == 0))) {
                child = parse_block(lines[pos].indent); // nested map, or a list (which YAML allows at the same indent)
            }
            node.map.emplace_back(key, child);
        }
        return node;
    }
};
} // namespace

YamlNode parse_yaml_subset(const std::string& text)
{
    Parser p;
    std::istringstream in(text);
    std::string raw;
    int n = 0;
    while (std::getline(in, raw)) {
        ++n;
        if (raw.find('\t') != std::string::npos && trim(raw).size() && raw.find_first_not_of(" ") != std::string::npos &&
            raw[raw.find_first_not_of(" ")] == '\t')
            throw std::runtime_error("yaml line " + std::to_string(n) + ": tab indentation");
        std::string s = strip_comment(raw);
        const std::string t = trim(s);
        if (t.empty() || t == "---") continue;
        const int indent = static_cast<int>(s.find_first_not_of(' '));
        p.lines.push_back({indent, t, n});
    }
    if (p.lines.empty()) return YamlNode{};
    return p.parse_block(p.lines[0].indent);
}

// =================================================================================================
// NetCDF-3
// =================================================================================================
namespace {
uint64_t be(const unsigned char* p, int n)
{
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) v = (v << 8) | p[i];
    return v;
}
int type_size(int t)
{
    switch (t) {
    case 1: case 2: case 7: return 1;
    case 3: case 8: return 2;
    case 4: case 5: case 9: return 4;
    case 6: case 10: case 11: return 8;
    default: return 0;
    }
}
struct Cursor {
    FILE* f;
    int version;
    uint64_t file_size; // every length read from the header is checked against it: a malformed file cannot make us allocate or seek past it
    void need(void* dst, size_t n)
    {
        if (std::fread(dst, 1, n, f) != n) throw std::runtime_error("netcdf: truncated header");
    }
    void skip(uint64_t n)
    {
        if (n > file_size || fseeko(f, static_cast<off_t>(n), SEEK_CUR) != 0) throw std::runtime_error("netcdf: truncated header");
    }
    uint32_t u32() { unsigned char b[4]; need(b, 4); return static_cast<uint32_t>(be(b, 4)); }
    uint64_t u64() { unsigned char b[8]; need(b, 8); return be(b, 8); }
    uint64_t nonneg() { return version == 5 ? u64() : u32(); } // NON_NEG: 64-bit in CDF-5
    uint64_t offset() { return version == 1 ? u32() : u64(); }
    std::string name()
    {
        const uint64_t n = nonneg();
        if (n > file_size || n > (1u << 20)) throw std::runtime_error("netcdf: implausible name length in the header");
        std::string s(n, '\0');
        if (n) need(&s[0], n);
        const uint64_t pad = (4 - n % 4) % 4;
        if (pad) skip(pad);
        return s;
    }
    void skip_att_list()
    {
        const uint32_t tag = u32();
        const uint64_t n = nonneg();
        if (tag == 0 && n == 0) return;
        if (tag != 0x0C) throw std::runtime_error("netcdf: bad attribute list tag");
        for (uint64_t i = 0; i < n; ++i) {
            (void)name();
            const uint32_t t = u32();
            const uint64_t ne = nonneg();
            if (ne > file_size) throw std::runtime_error("netcdf: implausible attribute length in the header");
            uint64_t bytes = ne * static_cast<uint64_t>(type_size(static_cast<int>(t)));
            bytes += (4 - bytes % 4) % 4;
            skip(bytes);
        }
    }
};
template <class T>
T convert_be(const unsigned char* p, int type)
{
    switch (type) {
    case 1: return static_cast<T>(static_cast<int8_t>(p[0]));
    case 2: case 7: return static_cast<T>(p[0]);
    case 3: return static_cast<T>(static_cast<int16_t>(be(p, 2)));
    case 8: return static_cast<T>(static_cast<uint16_t>(be(p, 2)));
    case 4: return static_cast<T>(static_cast<int32_t>(be(p, 4)));
    case 9: return static_cast<T>(static_cast<uint32_t>(be(p, 4)));
    case 5: { uint32_t u = static_cast<uint32_t>(be(p, 4)); float f; std::memcpy(&f, &u, 4); return static_cast<T>(f); }
    case 6: { uint64_t u = be(p, 8); double d; std::memcpy(&d, &u, 8); return static_cast<T>(d); }
    case 10: return static_cast<T>(static_cast<int64_t>(be(p, 8)));
    case 11: return static_cast<T>(be(p, 8));
    default: return T();
    }
}
} // namespace

NcFile::NcFile(const std::string& path) : path_(path)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("netcdf: cannot open " + path);
    // the destructor does not run when this constructor throws: close the file on every exit path but the normal one
    struct Guard {
        FILE* f;
        ~Guard() { if (f) std::fclose(f); }
    } guard{f};
    if (fseeko(f, 0, SEEK_END) != 0) throw std::runtime_error("netcdf: cannot seek in " + path);
    const uint64_t file_size = static_cast<uint64_t>(ftello(f));
    if (fseeko(f, 0, SEEK_SET) != 0) throw std::runtime_error("netcdf: cannot seek in " + path);
    unsigned char magic[4];
    if (std::fread(magic, 1, 4, f) != 4) throw std::runtime_error("netcdf: empty file " + path);
    if (magic[0] == 0x89 && magic[1] == 'H' && magic[2] == 'D' && magic[3] == 'F')
        throw std::runtime_error("netcdf: " + path + " is a netCDF-4/HDF5 file; only NetCDF-3 (classic, 64-bit offset, CDF-5) is supported -- "
                                 "convert with `nccopy -k cdf5`");
    if (magic[0] != 'C' || magic[1] != 'D' || magic[2] != 'F' || (magic[3] != 1 && magic[3] != 2 && magic[3] != 5))
        throw std::runtime_error("netcdf: " + path + " is not a NetCDF-3 file");
    version_ = magic[3];
    Cursor c{f, version_, file_size};
    numrecs_ = c.nonneg();
    // dimensions
    {
        const uint32_t tag = c.u32();
        const uint64_t n = c.nonneg();
        if (!(tag == 0 && n == 0)) {
            if (tag != 0x0A) throw std::runtime_error("netcdf: bad dimension list tag");
            if (n > file_size) throw std::runtime_error("netcdf: implausible dimension count");
            for (uint64_t i = 0; i < n; ++i) {
                std::string nm = c.name();
                const uint64_t len = c.nonneg();
                if (len == 0) recdim_ = static_cast<int>(i);
                dims_.emplace_back(nm, len);
            }
        }
    }
    c.skip_att_list();
    // variables
    {
        const uint32_t tag = c.u32();
        const uint64_t n = c.nonneg();
        if (!(tag == 0 && n == 0)) {
            if (tag != 0x0B) throw std::runtime_error("netcdf: bad variable list tag");
            if (n > file_size) throw std::runtime_error("netcdf: implausible variable count");
            for (uint64_t i = 0; i < n; ++i) {
                NcVar v;
                v.name = c.name();
                const uint64_t nd = c.nonneg();
                if (nd > 1024) throw std::runtime_error("netcdf: implausible rank of variable " + v.name);
                for (uint64_t d = 0; d < nd; ++d) {
                    const uint64_t id = c.nonneg();
                    if (id >= dims_.size()) throw std::runtime_error("netcdf: variable " + v.name + " names a dimension that does not exist");
                    v.dimids.push_back(static_cast<int>(id));
                }
                c.skip_att_list();
                v.type = static_cast<int>(c.u32());
                v.vsize = c.nonneg();
                v.begin = c.offset();
                v.is_record = !v.dimids.empty() && v.dimids[0] == recdim_;
                v.elems_per_record = 1;
                for (size_t d = v.is_record ? 1 : 0; d < v.dimids.size(); ++d) v.elems_per_record *= dims_[v.dimids[d]].second;
                vars_.push_back(v);
            }
        }
    }
    int nrec_vars = 0;
    for (auto& v : vars_)
        if (v.is_record) { recsize_ += v.vsize; ++nrec_vars; }
    // a single record variable is stored without padding between records
    if (nrec_vars == 1)
        for (auto& v : vars_)
            if (v.is_record) recsize_ = v.elems_per_record * static_cast<uint64_t>(type_size(v.type));
    if (numrecs_ == 0xFFFFFFFFull && version_ != 5 && recsize_ > 0) { // STREAMING: derive from the file size
        const uint64_t size = file_size;
        uint64_t first = UINT64_MAX;
        for (auto& v : vars_)
            if (v.is_record) first = std::min(first, v.begin);
        numrecs_ = first < size ? (size - first) / recsize_ : 0;
    }
    fp_ = f;
    file_size_ = file_size;
    guard.f = nullptr; // ownership passes to the object
}

NcFile::~NcFile()
{
    if (fp_) std::fclose(static_cast<FILE*>(fp_));
}

const NcVar* NcFile::var(const std::string& name) const
{
    for (auto& v : vars_)
        if (v.name == name) return &v;
    return nullptr;
}

uint64_t NcFile::dim_len(const std::string& name) const
{
    for (size_t i = 0; i < dims_.size(); ++i)
        if (dims_[i].first == name) return static_cast<int>(i) == recdim_ ? numrecs_ : dims_[i].second;
    return 0;
}

template <class T>
bool NcFile::read_as(const std::string& name, int64_t record, std::vector<T>& out) const
{
    const NcVar* v = var(name);
    out.clear();
    if (!v) return false;
    const int ts = type_size(v->type);
    if (ts == 0) return false;
    FILE* f = static_cast<FILE*>(fp_);
    const uint64_t per = v->elems_per_record;
    auto read_block = [&](uint64_t offset, uint64_t count) {
        // sizes come from the header: never allocate or read beyond what the file can hold
        if (offset > file_size_ || count > (file_size_ - offset) / static_cast<uint64_t>(ts))
            throw std::runtime_error("netcdf: variable " + name + " extends past the end of " + path_);
        std::vector<unsigned char> buf(count * ts);
        if (fseeko(f, static_cast<off_t>(offset), SEEK_SET) != 0 || std::fread(buf.data(), 1, buf.size(), f) != buf.size())
            throw std::runtime_error("netcdf: short read of " + name + " in " + path_);
        const size_t base = out.size();
        out.resize(base + count);
        for (uint64_t i = 0; i < count; ++i) out[base + i] = convert_be<T>(buf.data() + i * ts, v->type);
    };
    if (!v->is_record) {
        read_block(v->begin, per);
    } else if (record >= 0) {
        if (static_cast<uint64_t>(record) >= numrecs_) return false;
        read_block(v->begin + static_cast<uint64_t>(record) * recsize_, per);
    } else {
        for (uint64_t r = 0; r < numrecs_; ++r) read_block(v->begin + r * recsize_, per);
    }
    return true;
}
bool NcFile::read_double(const std::string& name, int64_t record, std::vector<double>& out) const { return read_as<double>(name, record, out); }
bool NcFile::read_int(const std::string& name, int64_t record, std::vector<int64_t>& out) const { return read_as<int64_t>(name, record, out); }
bool NcFile::read_char(const std::string& name, int64_t record, std::vector<char>& out) const { return read_as<char>(name, record, out); }

// =================================================================================================
// stream
// =================================================================================================
void Stream::parse_yaml(const std::string& yaml_path)
{
    std::ifstream in(yaml_path);
    if (!in) throw std::runtime_error("stream: cannot open " + yaml_path);
    std::stringstream ss;
    ss << in.rdbuf();
    const YamlNode root = parse_yaml_subset(ss.str());
    const YamlNode* st = root.get("stream");
    if (!st || st->kind != YamlNode::Map) throw std::runtime_error("stream: top-level 'stream:' map missing in " + yaml_path);
    path_prefix = st->str("path_prefix");
    const YamlNode* subs = st->get("substreams");
    if (!subs || subs->kind != YamlNode::List) throw std::runtime_error("stream: 'substreams:' list missing");
    substreams.clear();
    for (const YamlNode& sn : subs->list) {
        auto sub = std::make_shared<Substream>();
        sub->name = sn.str("name");
        sub->format = sn.str("format", "netcdf");
        sub->is_static = sn.boolean("static", false);
        std::vector<std::string> patterns;
        if (const YamlNode* fn = sn.get("filenames")) {
            if (fn->kind == YamlNode::Scalar) patterns.push_back(fn->scalar);
            else if (fn->kind == YamlNode::List)
                for (auto& e : fn->list) patterns.push_back(e.scalar);
        }
        for (const std::string& pat : patterns) {
            const std::string full = (path_prefix.empty() || (!pat.empty() && pat[0] == '/')) ? pat : path_prefix + "/" + pat;
            glob_t g;
            std::memset(&g, 0, sizeof(g));
            if (glob(full.c_str(), 0, nullptr, &g) == 0) {
                for (size_t i = 0; i < g.gl_pathc; ++i) {
                    const std::string p = g.gl_pathv[i];
                    sub->paths.push_back(p);
                    sub->filenames.push_back(path_prefix.empty() ? p : (p.rfind(path_prefix + "/", 0) == 0 ? p.substr(path_prefix.size() + 1) : p));
                }
            }
            globfree(&g);
        }
        if (const YamlNode* vs = sn.get("vars"))
            for (const YamlNode& vn : vs->list) {
                StreamVar sv;
                sv.name = vn.str("name");
                sv.optional = vn.boolean("optional", false);
                if (const YamlNode* pn = vn.get("possible_names"))
                    for (auto& e : pn->list) sv.possible_names.push_back(e.scalar);
                sub->vars.push_back(sv);
            }
        if (!sub->is_static) {
            int first = 0;
            for (const std::string& p : sub->paths) {
                sub->first_timestep_per_file.push_back(first);
                NcFile f(p);
                first += static_cast<int>(std::max<uint64_t>(f.num_records(), 1));
            }
        }
        substreams.push_back(sub);
    }
}

std::shared_ptr<NcFile> Stream::open_static() const
{
    for (auto& s : substreams)
        if (s->is_static) {
            if (s->paths.empty()) throw std::runtime_error("stream: mesh file not found (" + s->name + ")");
            return std::make_shared<NcFile>(s->paths[0]);
        }
    throw std::runtime_error("stream: no static substream");
}

int Stream::total_timesteps() const
{
    for (auto& s : substreams)
        if (!s->is_static && !s->paths.empty()) {
            NcFile f(s->paths.back());
            return s->first_timestep_per_file.back() + static_cast<int>(std::max<uint64_t>(f.num_records(), 1));
        }
    return 0;
}

std::shared_ptr<NcFile> Stream::open_record(int index, int64_t& local_record) const
{
    for (auto& s : substreams)
        if (!s->is_static) {
            for (size_t i = s->paths.size(); i-- > 0;)
                if (index >= s->first_timestep_per_file[i]) {
                    local_record = index - s->first_timestep_per_file[i];
                    return std::make_shared<NcFile>(s->paths[i]);
                }
        }
    throw std::runtime_error("stream: no data substream / bad timestep index");
}

std::string Stream::resolve(const Substream& sub, const NcFile& f, const std::string& wanted) const
{
    if (f.var(wanted)) return wanted;
    for (auto& v : sub.vars) {
        const bool mine = v.name == wanted || std::find(v.possible_names.begin(), v.possible_names.end(), wanted) != v.possible_names.end();
        if (!mine) continue;
        if (f.var(v.name)) return v.name;
        for (auto& a : v.possible_names)
            if (f.var(a)) return a;
    }
    return "";
}

} // namespace io
} // namespace MOPS
