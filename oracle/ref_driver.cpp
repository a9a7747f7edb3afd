// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product.
//
// ref_driver: a flat C ABI over the *unmodified* reference (YosefQiu/MOPS, TBB/CPU
// backend) so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can
// feed it synthetic MPAS-format arrays through ctypes and read back flat results.
// It is compiled by oracle/build_ref.sh together with the reference's own translation
// units taken where they lie under /root/reference (nothing is copied into this repo);
// the outputs go to the git-ignored oracle/_ref/.
//
// It follows the pyMOPS route for feeding data without netCDF
// (tools/pyMOPS/bindings.cpp:103-223): public setters on MPASOGrid / MPASOSolution,
// then MOPS_Init / Begin / AddGridMesh / AddAttribute / End / ActiveAttribute / Run*
// (include/api/MOPS.h:20-102, call order of tutorial/streamLine.cpp:72-103).
#include "api/MOPS.h"
#include "Core/MPASOVisualizer.h"
#include "Core/CPUContext.h"
#include "Core/RuntimeContext.h"
#include <tbb/parallel_for.h>

#include <cstdint>
#include <cstring>
#include <filesystem>
#include <memory>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace MOPS;

namespace {
std::shared_ptr<MPASOGrid> g_grid;
std::string g_cache_dir;
int g_levels = 0;
bool g_inited = false;
std::map<int, std::shared_ptr<MPASOSolution>> g_sols;
// MOPSApp::addSol silently ignores a solID it has already seen (src/Core/MOPSApp.cpp:82-87)
// and its map is private, so each refo_init() session maps caller ids into a fresh range.
int g_session = 0;
inline int internal_id(int sol_id) { return g_session * 4096 + sol_id; }

std::vector<vec3> to_vec3(const double* p, size_t n)
{
    std::vector<vec3> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    return v;
}
std::vector<size_t> to_sizet(const int32_t* p, size_t n)
{
    std::vector<size_t> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = static_cast<size_t>(p[i]);
    return v;
}
} // namespace

extern "C" {

int refo_sizeof_vec3() { return static_cast<int>(sizeof(vec3)); }

int refo_max_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void refo_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// cache_dir must be a fresh, existing, writable directory: the reference caches every
// preprocessing product there keyed only by the snapshot's timestep index
// (src/Core/MPASOSolution.cpp:22-27) and throws if KDTree.bin cannot be written
// (src/Core/MPASOGrid.cpp:247-274).
int refo_init(const char* cache_dir)
{
    g_cache_dir = cache_dir;
    std::filesystem::create_directories(g_cache_dir);
    if (!g_inited) {
        MOPS_Init("cpu");
        g_inited = true;
    }
    g_sols.clear();
    ++g_session;
    MOPS_Begin();
    return 0;
}

// Connectivity arrays are 1-based and 0-padded, exactly as in an MPAS file
// (src/IO/MPASOReader.cpp:147-153 keeps them 1-based; kernels subtract 1).
int refo_set_mesh(int n_cells, int n_vertices, int max_edges, int n_levels,
                  const double* cell_xyz, const double* vertex_xyz,
                  const int32_t* vertices_on_cell, const int32_t* cells_on_cell,
                  const int32_t* cells_on_vertex, const int32_t* n_edges_on_cell,
                  const double* ref_bottom_depth)
{
    g_grid = std::make_shared<MPASOGrid>();
    g_levels = n_levels;
    g_grid->setGridAttribute(GridAttributeType::kCellSize, n_cells);
    g_grid->setGridAttribute(GridAttributeType::kVertexSize, n_vertices);
    g_grid->setGridAttribute(GridAttributeType::kMaxEdgesSize, max_edges);
    g_grid->setGridAttribute(GridAttributeType::kEdgeSize, 0);
    g_grid->setGridAttribute(GridAttributeType::kVertLevels, n_levels);
    g_grid->setGridAttribute(GridAttributeType::kVertLevelsP1, n_levels + 1);
    g_grid->setGridAttributesVec3(GridAttributeType::kCellCoord, to_vec3(cell_xyz, n_cells));
    g_grid->setGridAttributesVec3(GridAttributeType::kVertexCoord, to_vec3(vertex_xyz, n_vertices));
    g_grid->setGridAttributesInt(GridAttributeType::kVerticesOnCell,
                                 to_sizet(vertices_on_cell, static_cast<size_t>(n_cells) * max_edges));
    g_grid->setGridAttributesInt(GridAttributeType::kCellsOnCell,
                                 to_sizet(cells_on_cell, static_cast<size_t>(n_cells) * max_edges));
    g_grid->setGridAttributesInt(GridAttributeType::kCellsOnVertex,
                                 to_sizet(cells_on_vertex, static_cast<size_t>(n_vertices) * 3));
    g_grid->setGridAttributesInt(GridAttributeType::kNumberVertexOnCell, to_sizet(n_edges_on_cell, n_cells));
    if (ref_bottom_depth != nullptr) {
        g_grid->cellRefBottomDepth_vec.assign(ref_bottom_depth, ref_bottom_depth + n_levels);
    }
    g_grid->mMeshName = "oracle_mesh";
    g_grid->mCachedDataDir = g_cache_dir;
    MOPS_AddGridMesh(g_grid);
    return 0;
}

// Cell-major [n_cells * n_levels] inputs; vert_vel_top is [n_cells * (n_levels+1)] or
// NULL (then zeros: the reference indexes it unconditionally,
// src/CPU/TBB/MPASOSolutionTBB.cpp:348-349, and has no setter for it).
// `timestep_tag` must be distinct per snapshot (the disk-cache key).
int refo_add_snapshot(int sol_id, int timestep_tag,
                      const double* zonal, const double* meridional,
                      const double* layer_thickness, const double* bottom_depth,
                      const double* vert_vel_top,
                      int n_attr, const char** attr_names, const double** attrs)
{
    if (!g_grid) return -1;
    const size_t nc = static_cast<size_t>(g_grid->mCellsSize);
    const size_t L = static_cast<size_t>(g_levels);
    auto sol = std::make_shared<MPASOSolution>();
    sol->mCellsSize = g_grid->mCellsSize;
    sol->mEdgesSize = 0;
    sol->mMaxEdgesSize = g_grid->mMaxEdgesSize;
    sol->mVertexSize = g_grid->mVertexSize;
    sol->mVertLevels = g_levels;
    sol->mVertLevelsP1 = g_levels + 1;
    sol->mTimesteps = timestep_tag;
    sol->mTotalZTopLayer = 0;
    sol->mTotalZTopLayerP1 = 0;
    sol->mTimeStamp = "0000-01-01_00:00:00";
    sol->setAttributesDouble(AttributeType::kZonalVelocity, std::vector<double>(zonal, zonal + nc * L));
    sol->setAttributesDouble(AttributeType::kMeridionalVelocity, std::vector<double>(meridional, meridional + nc * L));
    sol->setAttributesDouble(AttributeType::kLayerThickness, std::vector<double>(layer_thickness, layer_thickness + nc * L));
    sol->setAttributesDouble(AttributeType::kBottomDepth, std::vector<double>(bottom_depth, bottom_depth + nc));
    if (vert_vel_top != nullptr) {
        sol->cellVertVelocity_vec.assign(vert_vel_top, vert_vel_top + nc * (L + 1));
    } else {
        sol->cellVertVelocity_vec.assign(nc * (L + 1), 0.0);
    }
    for (int a = 0; a < n_attr; ++a) {
        sol->mDoubleAttributes[attr_names[a]] = std::vector<double>(attrs[a], attrs[a] + nc * L);
    }
    MOPS_AddAttribute(internal_id(sol_id), sol);
    g_sols[sol_id] = sol;
    return 0;
}

int refo_end()
{
    MOPS_End();
    return 0;
}

int refo_activate(int id1, int id2)
{
    if (id2 >= 0) {
        MOPS_ActiveAttribute(internal_id(id1), internal_id(id2));
    } else {
        MOPS_ActiveAttribute(internal_id(id1));
    }
    return 0;
}

// Products of the reference's preprocessing chain (MOPSApp::addSol,
// src/Core/MOPSApp.cpp:100-130), vertex-major.  Any pointer may be NULL.
int refo_get_prepared(int sol_id, double* ztop_vertex, double* vel_vertex, double* vertvel_vertex,
                      double* ztop_cell, double* vel_cell)
{
    auto it = g_sols.find(sol_id);
    if (it == g_sols.end()) return -1;
    auto& s = *it->second;
    if (ztop_vertex) std::memcpy(ztop_vertex, s.cellVertexZTop_vec.data(), s.cellVertexZTop_vec.size() * sizeof(double));
    if (vel_vertex) std::memcpy(vel_vertex, s.cellVertexVelocity_vec.data(), s.cellVertexVelocity_vec.size() * sizeof(vec3));
    if (vertvel_vertex) std::memcpy(vertvel_vertex, s.cellVertexVertVelocity_vec.data(), s.cellVertexVertVelocity_vec.size() * sizeof(double));
    if (ztop_cell) std::memcpy(ztop_cell, s.cellZTop_vec.data(), s.cellZTop_vec.size() * sizeof(double));
    if (vel_cell) std::memcpy(vel_cell, s.cellCenterVelocity_vec.data(), s.cellCenterVelocity_vec.size() * sizeof(vec3));
    return 0;
}

int refo_get_prepared_attr(int sol_id, const char* name, double* attr_vertex)
{
    auto it = g_sols.find(sol_id);
    if (it == g_sols.end()) return -1;
    auto jt = it->second->mDoubleAttributes_CtoV.find(name);
    if (jt == it->second->mDoubleAttributes_CtoV.end()) return -2;
    std::memcpy(attr_vertex, jt->second.data(), jt->second.size() * sizeof(double));
    return 0;
}

// Exact nearest cell centre through the reference's own nanoflann tree
// (MPASOField::calcInWhichCells -> MPASOGrid::searchKDT, src/Core/MPASOGrid.cpp:287-313).
int refo_locate(int64_t n, const double* xyz, int32_t* cells)
{
    if (!g_grid) return -1;
    tbb::parallel_for(int64_t(0), n, [&](int64_t i) {
        int c = -1;
        g_grid->searchKDT(vec3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), c);
        cells[i] = c;
    });
    return 0;
}

// Seed grid of MPASOVisualizer::GenerateSamplePoint (src/Core/MPASOVisualizer.cpp:120-149).
// Returns the number of points; writes at most `cap` of them.
int64_t refo_generate_seeds(int nx, int ny, double lat_min, double lat_max, double lon_min, double lon_max,
                            double depth, double* xyz_out, int64_t cap)
{
    SamplingSettings s;
    s.setSampleRange(vec2i(nx, ny));
    s.setGeoBox(vec2(lat_min, lat_max), vec2(lon_min, lon_max));
    s.setDepth(depth);
    std::vector<CartesianCoord> pts;
    MOPS_GenerateSamplePoints(&s, pts);
    const int64_t n = static_cast<int64_t>(pts.size());
    for (int64_t i = 0; i < n && i < cap; ++i) {
        xyz_out[3 * i] = pts[i].x();
        xyz_out[3 * i + 1] = pts[i].y();
        xyz_out[3 * i + 2] = pts[i].z();
    }
    return n;
}

static void fill_settings(TrajectorySettings& cfg, int method_rk4, int forward, int64_t delta_t, int64_t duration,
                          int64_t record_t, float depth, const float* depths, int64_t n)
{
    cfg.deltaT = static_cast<size_t>(delta_t);
    cfg.simulationDuration = static_cast<size_t>(duration);
    cfg.recordT = static_cast<size_t>(record_t);
    cfg.depth = depth;
    if (depths != nullptr) cfg.particle_depths.assign(depths, depths + n);
    cfg.directionType = forward ? CalcDirection::kForward : CalcDirection::kBackward;
    cfg.methodType = method_rk4 ? CalcMethodType::kRK4 : CalcMethodType::kEuler;
}

// out_points / out_vel: [n][each+1][3] with each = duration / record_t, i.e. exactly the
// assembled TrajectoryLine::points / ::velocity (seed first, velocity shifted by one and
// zero-padded -- src/Common/TrajectoryCommon.h:43-55,131-157).  Returns lines.size().
int64_t refo_streamline(int method_rk4, int forward, int64_t delta_t, int64_t duration, int64_t record_t,
                        float depth, const float* depths, int64_t n, const double* seeds_xyz,
                        double* out_points, double* out_vel, double* out_last, double* seconds)
{
    TrajectorySettings cfg;
    fill_settings(cfg, method_rk4, forward, delta_t, duration, record_t, depth, depths, n);
    std::vector<CartesianCoord> pts = to_vec3(seeds_xyz, static_cast<size_t>(n));
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<TrajectoryLine> lines = MOPS_RunStreamLine(&cfg, pts);
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    const size_t per = static_cast<size_t>(duration / record_t) + 1;
    for (size_t i = 0; i < lines.size(); ++i) {
        const auto& ln = lines[i];
        if (out_points) {
            for (size_t k = 0; k < per && k < ln.points.size(); ++k) {
                double* d = out_points + (i * per + k) * 3;
                d[0] = ln.points[k].x(); d[1] = ln.points[k].y(); d[2] = ln.points[k].z();
            }
        }
        if (out_vel) {
            for (size_t k = 0; k < per && k < ln.velocity.size(); ++k) {
                double* d = out_vel + (i * per + k) * 3;
                d[0] = ln.velocity[k].x(); d[1] = ln.velocity[k].y(); d[2] = ln.velocity[k].z();
            }
        }
        if (out_last) {
            out_last[3 * i] = ln.lastPoint.x(); out_last[3 * i + 1] = ln.lastPoint.y(); out_last[3 * i + 2] = ln.lastPoint.z();
        }
    }
    return static_cast<int64_t>(lines.size());
}

// As refo_streamline, between the two active snapshots.  seeds_xyz_inout is overwritten
// with each line's lastPoint, as MOPSApp::runPathLine does (src/Core/MOPSApp.cpp:287-290).
// out_temp / out_sal: [n][each+1] TrajectoryLine::temperature / ::salinity (which the
// reference fills from velocity.x / velocity.y, src/Common/TrajectoryCommon.h:179-180).
int64_t refo_pathline(int method_rk4, int forward, int64_t delta_t, int64_t duration, int64_t record_t,
                      float depth, const float* depths, int64_t n, double* seeds_xyz_inout,
                      double* out_points, double* out_vel, double* out_temp, double* out_sal, double* seconds)
{
    TrajectorySettings cfg;
    fill_settings(cfg, method_rk4, forward, delta_t, duration, record_t, depth, depths, n);
    std::vector<CartesianCoord> pts = to_vec3(seeds_xyz_inout, static_cast<size_t>(n));
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<TrajectoryLine> lines = MOPS_RunPathLine(&cfg, pts);
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    const size_t per = static_cast<size_t>(duration / record_t) + 1;
    for (size_t i = 0; i < lines.size(); ++i) {
        const auto& ln = lines[i];
        for (size_t k = 0; k < per; ++k) {
            if (out_points && k < ln.points.size()) {
                double* d = out_points + (i * per + k) * 3;
                d[0] = ln.points[k].x(); d[1] = ln.points[k].y(); d[2] = ln.points[k].z();
            }
            if (out_vel && k < ln.velocity.size()) {
                double* d = out_vel + (i * per + k) * 3;
                d[0] = ln.velocity[k].x(); d[1] = ln.velocity[k].y(); d[2] = ln.velocity[k].z();
            }
            if (out_temp && k < ln.temperature.size()) out_temp[i * per + k] = ln.temperature[k];
            if (out_sal && k < ln.salinity.size()) out_sal[i * per + k] = ln.salinity[k];
        }
    }
    for (size_t i = 0; i < pts.size(); ++i) {
        seeds_xyz_inout[3 * i] = pts[i].x(); seeds_xyz_inout[3 * i + 1] = pts[i].y(); seeds_xyz_inout[3 * i + 2] = pts[i].z();
    }
    return static_cast<int64_t>(lines.size());
}

// img0 / img1: [height][width][4] doubles (ImageBuffer layout, src/Common/ImageBuffer.hpp:13-30).
// Returns the number of images the reference produced.
int refo_remap_fixed_depth(int width, int height, double lat_min, double lat_max, double lon_min, double lon_max,
                           double fixed_depth, double* img0, double* img1, double* seconds)
{
    VisualizationSettings cfg;
    cfg.imageSize = vec2(width, height);
    cfg.LatRange = vec2(lat_min, lat_max);
    cfg.LonRange = vec2(lon_min, lon_max);
    cfg.FixedDepth = fixed_depth;
    cfg.TimeStep = 0.0;
    cfg.CalcType = CalcAttributeType::kZonalMerimoal;
    cfg.VisType = VisualizeType::kFixedDepth;
    cfg.PositionType = CalcPositionType::kPoint;
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<ImageBuffer<double>> imgs = MOPS_RunRemapping(&cfg);
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    const size_t bytes = static_cast<size_t>(width) * height * 4 * sizeof(double);
    if (img0 && imgs.size() > 0) std::memcpy(img0, imgs[0].mPixels.data(), bytes);
    if (img1 && imgs.size() > 1) std::memcpy(img1, imgs[1].mPixels.data(), bytes);
    return static_cast<int>(imgs.size());
}

// VisualizeFixedLayer through the reference's facade (src/Core/MPASOVisualizer.cpp:17-20)
int refo_remap_fixed_layer(int width, int height, double lat_min, double lat_max, double lon_min, double lon_max,
                           int layer, double* img)
{
    VisualizationSettings cfg;
    cfg.imageSize = vec2(width, height);
    cfg.LatRange = vec2(lat_min, lat_max);
    cfg.LonRange = vec2(lon_min, lon_max);
    cfg.FixedLayer = layer;
    ImageBuffer<double> im(width, height);
    CPUContext cpu_ctx;
    cpu_ctx.backend = CPUBackend::kTBB;
    MPASOVisualizer::VisualizeFixedLayer(MOPS_GetFieldSnapshots().get(), &cfg, &im, RuntimeContext::FromCPU(cpu_ctx));
    std::memcpy(img, im.mPixels.data(), static_cast<size_t>(width) * height * 4 * sizeof(double));
    return 0;
}

// VisualizeFixedLatitude through MOPSApp::runReGrid (src/Core/MOPSApp.cpp:198-210)
int refo_regrid_fixed_latitude(int width, int height, double lon_min, double lon_max, double latitude, double* img)
{
    VisualizationSettings cfg;
    cfg.imageSize = vec2(width, height);
    cfg.LonRange = vec2(lon_min, lon_max);
    cfg.LatRange = vec2(-90.0, 90.0);
    cfg.FixedLatitude = latitude;
    cfg.FixedDepth = 0.0;
    ImageBuffer<double> im = app.runReGrid(&cfg);
    std::memcpy(img, im.mPixels.data(), static_cast<size_t>(width) * height * 4 * sizeof(double));
    return 0;
}

// Known answers the reference's own unit tests hold for hot-adjacent math
// (test/test_gaussian.cpp:9-28) and a Wachspress spot check.
int refo_gauss3(const double* a9, const double* b3, double* x3)
{
    double A[MAX_EDGE][MAX_EDGE] = {};
    double b[MAX_EDGE] = {};
    double x[MAX_EDGE] = {};
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) A[i][j] = a9[3 * i + j];
        b[i] = b3[i];
    }
    Interpolator::gauss_elimination_fixed(A, b, 3, x);
    for (int i = 0; i < 3; ++i) x3[i] = x[i];
    return 0;
}

int refo_wachspress(const double* p3, const double* poly, int nv, double* w)
{
    std::vector<vec3> pv = to_vec3(poly, static_cast<size_t>(nv));
    vec3 p(p3[0], p3[1], p3[2]);
    Interpolator::CalcPolygonWachspress(p, pv.data(), w, nv);
    return 0;
}

} // extern "C"
