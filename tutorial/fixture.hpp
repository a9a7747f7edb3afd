// fixture.hpp -- loads a synthetic MPAS-format fixture (written by mops_b200.synthetic.dump_fixture)
// into MPASOGrid / MPASOSolution objects through their public setters, i.e. the route pyMOPS uses to
// feed the reference without netCDF (tools/pyMOPS/bindings.cpp:103-223).  The reference's tutorials
// read NERSC netCDF files through MPASOReader instead; file ingestion is a "next" row.
#pragma once
#include "api/MOPS.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace fixture {

struct Loaded {
    std::shared_ptr<MOPS::MPASOGrid> grid;
    std::vector<std::shared_ptr<MOPS::MPASOSolution>> sols;
};

template <class T>
inline std::vector<T> read_vec(FILE* f, size_t n)
{
    std::vector<T> v(n);
    if (n && std::fread(v.data(), sizeof(T), n, f) != n) {
        std::fprintf(stderr, "fixture: short read\n");
        std::exit(2);
    }
    return v;
}

inline Loaded load(const std::string& path)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) {
        std::fprintf(stderr, "fixture: cannot open %s\n", path.c_str());
        std::exit(2);
    }
    char magic[8];
    if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "MOPSFIX1", 8) != 0) {
        std::fprintf(stderr, "fixture: bad magic\n");
        std::exit(2);
    }
    const auto hdr = read_vec<int32_t>(f, 6);
    const size_t nC = hdr[0], nV = hdr[1], E = hdr[2], L = hdr[3], nS = hdr[4], nA = hdr[5];
    Loaded out;
    out.grid = std::make_shared<MOPS::MPASOGrid>();
    auto& g = *out.grid;
    g.setGridAttribute(MOPS::GridAttributeType::kCellSize, (int)nC);
    g.setGridAttribute(MOPS::GridAttributeType::kVertexSize, (int)nV);
    g.setGridAttribute(MOPS::GridAttributeType::kMaxEdgesSize, (int)E);
    g.setGridAttribute(MOPS::GridAttributeType::kVertLevels, (int)L);
    g.setGridAttribute(MOPS::GridAttributeType::kVertLevelsP1, (int)L + 1);
    auto to_vec3 = [](const std::vector<double>& a) {
        std::vector<vec3> v(a.size() / 3);
        for (size_t i = 0; i < v.size(); ++i) v[i] = vec3(a[3 * i], a[3 * i + 1], a[3 * i + 2]);
        return v;
    };
    auto to_sz = [](const std::vector<int32_t>& a) { return std::vector<size_t>(a.begin(), a.end()); };
    g.setGridAttributesVec3(MOPS::GridAttributeType::kCellCoord, to_vec3(read_vec<double>(f, nC * 3)));
    g.setGridAttributesVec3(MOPS::GridAttributeType::kVertexCoord, to_vec3(read_vec<double>(f, nV * 3)));
    g.setGridAttributesInt(MOPS::GridAttributeType::kVerticesOnCell, to_sz(read_vec<int32_t>(f, nC * E)));
    g.setGridAttributesInt(MOPS::GridAttributeType::kCellsOnCell, to_sz(read_vec<int32_t>(f, nC * E)));
    g.setGridAttributesInt(MOPS::GridAttributeType::kCellsOnVertex, to_sz(read_vec<int32_t>(f, nV * 3)));
    g.setGridAttributesInt(MOPS::GridAttributeType::kNumberVertexOnCell, to_sz(read_vec<int32_t>(f, nC)));
    g.mMeshName = "synthetic";
    for (size_t s = 0; s < nS; ++s) {
        auto sol = std::make_shared<MOPS::MPASOSolution>();
        sol->mVertLevels = (int)L;
        sol->mVertLevelsP1 = (int)L + 1;
        sol->mTimesteps = (int)s;
        sol->mTimeStamp = "0001-01-" + std::to_string(s + 1);
        sol->setAttributesDouble(MOPS::AttributeType::kZonalVelocity, read_vec<double>(f, nC * L));
        sol->setAttributesDouble(MOPS::AttributeType::kMeridionalVelocity, read_vec<double>(f, nC * L));
        sol->setAttributesDouble(MOPS::AttributeType::kLayerThickness, read_vec<double>(f, nC * L));
        sol->setAttributesDouble(MOPS::AttributeType::kBottomDepth, read_vec<double>(f, nC));
        sol->cellVertVelocity_vec = read_vec<double>(f, nC * (L + 1));
        for (size_t a = 0; a < nA; ++a) {
            char name[33] = {0};
            if (std::fread(name, 1, 32, f) != 32) std::exit(2);
            sol->mDoubleAttributes[name] = read_vec<double>(f, nC * L);
        }
        out.sols.push_back(sol);
    }
    std::fclose(f);
    return out;
}

// flat dump of a set of lines: int64 n, int64 per, points[n][per][3], velocity[n][per][3], last[n][3]
inline void dump_lines(const std::string& path, const std::vector<MOPS::TrajectoryLine>& lines)
{
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) std::exit(2);
    const int64_t n = (int64_t)lines.size(), per = n ? (int64_t)lines[0].points.size() : 0;
    std::fwrite(&n, 8, 1, f);
    std::fwrite(&per, 8, 1, f);
    for (auto& l : lines) std::fwrite(l.points.data(), 24, (size_t)per, f);
    for (auto& l : lines) std::fwrite(l.velocity.data(), 24, (size_t)per, f);
    for (auto& l : lines) std::fwrite(&l.lastPoint, 24, 1, f);
    std::fclose(f);
}

} // namespace fixture
