// mops_reader.cpp -- MPASOReader / initGrid / initSolution / addAttribute of the drop-in API over
// mpas_io (YAML-subset stream + NetCDF-3).  Variable names and fall-backs follow the reference's
// reader (src/IO/MPASOReader.cpp:141-154, 215-235).
#include "api/MOPS.h"
#include "mpas_io.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <iostream>

namespace MOPS {
namespace {
struct OpenRecord {
    std::shared_ptr<io::Stream> stream;
    std::shared_ptr<io::NcFile> file;
    std::shared_ptr<io::Substream> sub;
    int64_t record = 0;
};

std::string strip_ext(const std::string& f)
{
    const size_t slash = f.find_last_of('/');
    const std::string base = slash == std::string::npos ? f : f.substr(slash + 1);
    const size_t dot = base.find_last_of('.');
    return dot == std::string::npos ? base : base.substr(0, dot);
}
bool read_d(const io::Stream& st, const io::Substream& sub, const io::NcFile& f, const std::string& name, int64_t rec, std::vector<double>& out)
{
    const std::string actual = st.resolve(sub, f, name);
    out.clear();
    if (actual.empty()) return false;
    const io::NcVar* v = f.var(actual);
    return f.read_double(actual, v->is_record ? rec : -1, out);
}
void read_sz(const io::NcFile& f, const std::string& name, std::vector<size_t>& out)
{
    std::vector<int64_t> tmp;
    out.clear();
    if (!f.read_int(name, -1, tmp)) return;
    out.assign(tmp.begin(), tmp.end());
}
void read_vec3(const io::NcFile& f, const char* x, const char* y, const char* z, std::vector<vec3>& out)
{
    std::vector<double> a, b, c;
    out.clear();
    if (!f.read_double(x, -1, a) || !f.read_double(y, -1, b) || !f.read_double(z, -1, c)) return;
    out.resize(a.size());
    for (size_t i = 0; i < a.size(); ++i) out[i] = vec3(a[i], b[i], c[i]);
}
} // namespace

MPASOReader::Ptr MPASOReader::readGridData(const std::string& yaml_path)
{
    auto reader = std::make_shared<MPASOReader>();
    reader->path = yaml_path;
    try {
        io::Stream st;
        st.parse_yaml(yaml_path);
        auto f = st.open_static();
        for (auto& s : st.substreams)
            if (s->is_static && !s->filenames.empty()) reader->mMeshName = strip_ext(s->filenames[0]);
        reader->mFolderName = st.path_prefix;
        read_vec3(*f, "xCell", "yCell", "zCell", reader->cellCoord_vec);
        read_vec3(*f, "xVertex", "yVertex", "zVertex", reader->vertexCoord_vec);
        read_vec3(*f, "xEdge", "yEdge", "zEdge", reader->edgeCoord_vec);
        std::vector<double> lat, lon;
        if (f->read_double("latVertex", -1, lat) && f->read_double("lonVertex", -1, lon)) {
            reader->vertexLatLon_vec.resize(lat.size());
            for (size_t i = 0; i < lat.size(); ++i) reader->vertexLatLon_vec[i] = vec2(lat[i], lon[i]);
        }
        read_sz(*f, "verticesOnCell", reader->verticesOnCell_vec);
        read_sz(*f, "verticesOnEdge", reader->verticesOnEdge_vec);
        read_sz(*f, "cellsOnVertex", reader->cellsOnVertex_vec);
        read_sz(*f, "cellsOnCell", reader->cellsOnCell_vec);
        read_sz(*f, "nEdgesOnCell", reader->numberVertexOnCell_vec);
        read_sz(*f, "cellsOnEdge", reader->cellsOnEdge_vec);
        read_sz(*f, "edgesOnCell", reader->edgesOnCell_vec);
        f->read_double("refBottomDepth", -1, reader->cellRefBottomDepth_vec);
        reader->mCellsSize = static_cast<int>(reader->cellCoord_vec.size());
        reader->mEdgesSize = static_cast<int>(reader->edgeCoord_vec.size());
        reader->mVertexSize = static_cast<int>(reader->vertexCoord_vec.size());
        if (reader->mCellsSize > 0) {
            // the reference derives maxEdges from edgesOnCell (:159); meshes written without edge arrays
            // fall back to the width of verticesOnCell
            const size_t rows = !reader->edgesOnCell_vec.empty() ? reader->edgesOnCell_vec.size() : reader->verticesOnCell_vec.size();
            reader->mMaxEdgesSize = static_cast<int>(rows / static_cast<size_t>(reader->mCellsSize));
        }
    } catch (const std::exception& e) {
        std::cerr << "[MPASOReader]::Error: " << e.what() << std::endl;
        std::exit(-1);
    }
    return reader;
}

MPASOReader::Ptr MPASOReader::readSolData(const std::string& yaml_path, const std::string& data_name, const int& timestep)
{
    auto reader = std::make_shared<MPASOReader>();
    reader->path = yaml_path;
    try {
        auto st = std::make_shared<io::Stream>();
        st->parse_yaml(yaml_path);
        std::shared_ptr<io::Substream> sub;
        for (auto& s : st->substreams)
            if (!s->is_static) { sub = s; break; }
        if (!sub) throw std::runtime_error("no data substream in " + yaml_path);
        auto it = std::find_if(sub->filenames.begin(), sub->filenames.end(),
                               [&](const std::string& fn) { return fn.find(data_name) != std::string::npos; });
        if (it == sub->filenames.end()) {
            std::cerr << "[MPASOReader]::Error: Data file with name containing '" << data_name << "' not found in YAML." << std::endl;
            std::exit(-1);
        }
        if (timestep < 0) {
            std::fprintf(stderr, "[MPASOReader]::Error: Invalid timestep index %d\n", timestep);
            std::exit(-1);
        }
        const int fi = static_cast<int>(std::distance(sub->filenames.begin(), it));
        const int index = sub->first_timestep_per_file[fi] + timestep;
        int64_t local = 0;
        auto f = st->open_record(index, local);
        auto rec = std::make_shared<OpenRecord>();
        rec->stream = st; rec->file = f; rec->sub = sub; rec->record = local;
        reader->mGroupT = rec;
        reader->mTimesteps = timestep;
        reader->mDataName = strip_ext(*it);
        reader->mFolderName = st->path_prefix;
        read_d(*st, *sub, *f, "bottomDepth", local, reader->cellBottomDepth_vec);
        read_d(*st, *sub, *f, "seaSurfaceHeight", local, reader->cellSurfaceHeight_vec);
        read_d(*st, *sub, *f, "velocityZonal", local, reader->cellZonalVelocity_vec);
        read_d(*st, *sub, *f, "velocityMeridional", local, reader->cellMeridionalVelocity_vec);
        read_d(*st, *sub, *f, "layerThickness", local, reader->cellLayerThickness_vec);
        read_d(*st, *sub, *f, "zTop", local, reader->cellZTop_vec);
        read_d(*st, *sub, *f, "normalVelocity", local, reader->cellNormalVelocity_vec);
        read_d(*st, *sub, *f, "vertVelocityTop", local, reader->cellVertVelocity_vec);
        const std::string xt = st->resolve(*sub, *f, "xtime");
        if (!xt.empty()) {
            std::vector<char> chars;
            f->read_char(xt, f->var(xt)->is_record ? local : -1, chars);
            reader->mTimeStamp.assign(chars.begin(), chars.end());
        }
        if (!reader->cellSurfaceHeight_vec.empty())
            reader->mVertLevels = static_cast<int>(reader->cellLayerThickness_vec.size() / reader->cellSurfaceHeight_vec.size());
        else if (!reader->cellBottomDepth_vec.empty())
            reader->mVertLevels = static_cast<int>(reader->cellLayerThickness_vec.size() / reader->cellBottomDepth_vec.size());
        reader->mVertLevelsP1 = reader->mVertLevels + 1;
    } catch (const std::exception& e) {
        std::cerr << "[MPASOReader]::Error: " << e.what() << std::endl;
        std::exit(-1);
    }
    return reader;
}

void MPASOGrid::initGrid_DemoLoading(const char* yaml_path)
{
    auto r = MPASOReader::readGridData(yaml_path);
    initGrid(r.get());
}

void MPASOSolution::initSolution_DemoLoading(const char* yaml_path, int timestep)
{
    // the global record index is resolved to (file, local record); the file is then addressed by name,
    // which is what readSolData expects
    try {
        io::Stream st;
        st.parse_yaml(yaml_path);
        for (auto& sub : st.substreams) {
            if (sub->is_static) continue;
            for (size_t i = sub->filenames.size(); i-- > 0;) {
                if (timestep >= sub->first_timestep_per_file[i]) {
                    auto r = MPASOReader::readSolData(yaml_path, sub->filenames[i], timestep - sub->first_timestep_per_file[i]);
                    initSolution(r.get());
                    mTimesteps = timestep;
                    return;
                }
            }
        }
        throw std::runtime_error("no data substream");
    } catch (const std::exception& e) {
        std::cerr << "[MPASOSolution]::initSolution_DemoLoading: " << e.what() << std::endl;
        std::exit(-1);
    }
}

void MPASOGrid::initGrid(MPASOReader* r)
{
    mCellsSize = r->mCellsSize; mEdgesSize = r->mEdgesSize; mMaxEdgesSize = r->mMaxEdgesSize; mVertexSize = r->mVertexSize;
    mVertLevels = r->mVertLevels; mVertLevelsP1 = r->mVertLevelsP1;
    vertexCoord_vec = std::move(r->vertexCoord_vec);
    cellCoord_vec = std::move(r->cellCoord_vec);
    edgeCoord_vec = std::move(r->edgeCoord_vec);
    vertexLatLon_vec = std::move(r->vertexLatLon_vec);
    verticesOnCell_vec = std::move(r->verticesOnCell_vec);
    verticesOnEdge_vec = std::move(r->verticesOnEdge_vec);
    cellsOnVertex_vec = std::move(r->cellsOnVertex_vec);
    cellsOnCell_vec = std::move(r->cellsOnCell_vec);
    numberVertexOnCell_vec = std::move(r->numberVertexOnCell_vec);
    cellsOnEdge_vec = std::move(r->cellsOnEdge_vec);
    edgesOnCell_vec = std::move(r->edgesOnCell_vec);
    cellRefBottomDepth_vec = std::move(r->cellRefBottomDepth_vec);
    mMeshName = r->mMeshName;
    mFolderPath = r->mFolderName;
}

void MPASOSolution::initSolution(MPASOReader* r)
{
    mTimeStamp = std::move(r->mTimeStamp);
    mVertLevels = r->mVertLevels; mVertLevelsP1 = r->mVertLevelsP1; mTimesteps = r->mTimesteps;
    mDataName = std::move(r->mDataName);
    gt = std::move(r->mGroupT);
    cellBottomDepth_vec = std::move(r->cellBottomDepth_vec);
    cellSurfaceHeight_vec = std::move(r->cellSurfaceHeight_vec);
    cellZonalVelocity_vec = std::move(r->cellZonalVelocity_vec);
    cellMeridionalVelocity_vec = std::move(r->cellMeridionalVelocity_vec);
    cellLayerThickness_vec = std::move(r->cellLayerThickness_vec);
    cellZTop_vec = std::move(r->cellZTop_vec);
    cellNormalVelocity_vec = std::move(r->cellNormalVelocity_vec);
    cellVertVelocity_vec = std::move(r->cellVertVelocity_vec);
}

void MPASOSolution::addAttribute(std::string name, AttributeFormat type)
{
    if (!gt) {
        std::cerr << "[MPASOSolution]::gt is not initialized" << std::endl;
        return;
    }
    if (type != AttributeFormat::kDouble && type != AttributeFormat::kFloat) return; // char / vec3: not on the hot path
    auto rec = std::static_pointer_cast<OpenRecord>(gt);
    std::vector<double> vec;
    try {
        read_d(*rec->stream, *rec->sub, *rec->file, name, rec->record, vec);
    } catch (const std::exception& e) {
        std::cerr << "[MPASOSolution]::addAttribute: " << e.what() << std::endl;
    }
    mDoubleAttributes[name] = vec; // an absent variable leaves an empty entry, as the reference does
}

} // namespace MOPS
