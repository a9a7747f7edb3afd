// include/api/MOPS.h -- C++ drop-in for the hot-path surface of YosefQiu/MOPS's public API,
// implemented on the B200 engine (include/mops_b200.h).
//
// A program written against the reference's `api/MOPS.h` for streamline / pathline /
// remapping (tutorial/streamLine.cpp, tutorial/pathLine.cpp, tutorial/reMapping.cpp, the
// pyMOPS bindings) keeps compiling against this header: same namespace, function names,
// argument meaning and error behaviour (reference: include/api/MOPS.h:20-148, semantics in
// src/Core/MOPS.cpp:10-127 and src/Core/MOPSApp.cpp:34-337), the same settings / result
// structs field for field (src/Core/MPASOVisualizer.h:12-103), the same setter-based route
// for feeding a grid and a solution without files (src/Core/MPASOGrid.cpp:82-186,
// src/Core/MPASOSolution.cpp:1150-1210; the route pyMOPS uses).  Everything underneath is
// new: no SYCL/HIP/TBB backends, no dispatch factories, no KD-tree, no disk cache, no CPU path.
//
// Out of scope here (SURVEY.md 8, rows marked out of scope / next): VTK writers, fixed-layer /
// fixed-latitude views, the RBF velocity path, netCDF-4/HDF5 files (NetCDF-3 only).
#pragma once

#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <thread>
#include <vector>

// ---- vector types (24-byte vec3 with x()/y()/z() accessors, as every call site of the
//      reference spells them; src/Utils/BackendCompat.hpp) -----------------------------------
struct vec2 {
    double c[2];
    vec2() : c{0.0, 0.0} {}
    vec2(double a, double b) : c{a, b} {}
    double& x() { return c[0]; }
    double& y() { return c[1]; }
    const double& x() const { return c[0]; }
    const double& y() const { return c[1]; }
};
struct vec2i {
    int c[2];
    vec2i() : c{0, 0} {}
    vec2i(int a, int b) : c{a, b} {}
    int& x() { return c[0]; }
    int& y() { return c[1]; }
    const int& x() const { return c[0]; }
    const int& y() const { return c[1]; }
};
struct vec3 {
    double c[3];
    vec3() : c{0.0, 0.0, 0.0} {}
    vec3(double a, double b, double d) : c{a, b, d} {}
    double& x() { return c[0]; }
    double& y() { return c[1]; }
    double& z() { return c[2]; }
    const double& x() const { return c[0]; }
    const double& y() const { return c[1]; }
    const double& z() const { return c[2]; }
};
static_assert(sizeof(vec3) == 24, "vec3 must keep the reference's 24-byte layout");
using SphericalCoord = vec2;
using CartesianCoord = vec3;

#define ONE_SECOND 1
#define ONE_MINUTE 60
#define ONE_HOUR 60 * 60
#define ONE_DAY 60 * 60 * 24
#define ONE_MONTH 60 * 60 * 24 * 30
#define ONE_YEAR 60 * 60 * 24 * 30 * 12

struct mops_ctx;   // the engine context (include/mops_b200.h)
struct mops_multi; // all devices of the box behind one handle (include/mops_b200.h)

namespace MOPS {

// ---- data model: what the caller fills (names as in src/Core/MPASOGrid.h, MPASOSolution.h) ----
enum class GridAttributeType : int {
    kCellSize, kEdgeSize, kVertexSize, kMaxEdgesSize, kVertLevels, kVertLevelsP1,
    kVertexCoord, kCellCoord, kEdgeCoord, kVertexLatLon, kVerticesOnCell, kVerticesOnEdge, kCellsOnVertex, kCellsOnCell,
    kNumberVertexOnCell, kCellsOnEdge, kEdgesOnCell, kCellWeight, krefBottomDepth, kCount
};
enum class AttributeFormat : int { kDouble, kFloat, kChar, kVec3, kCount };
enum class AttributeType : int { kZonalVelocity, kMeridionalVelocity, kVelocity, kNormalVelocity, kZTop, kLayerThickness, kBottomDepth, kCount };

// ---- file ingestion (src/IO/MPASOReader.h; implemented over mops_b200/host/mpas_io.*: YAML-subset
//      stream description + NetCDF-3 reader, no ndarray / netcdf-c / yaml-cpp) ---------------------
class MPASOReader {
public:
    using Ptr = std::shared_ptr<MPASOReader>;
    // mesh substream of the YAML stream description (src/IO/MPASOReader.cpp:128-169)
    static Ptr readGridData(const std::string& yaml_path);
    // snapshot `timestep` of the data file whose name contains `data_name` (src/IO/MPASOReader.cpp:171-245)
    static Ptr readSolData(const std::string& yaml_path, const std::string& data_name, const int& timestep);

    std::string path, mMeshName, mDataName, mFolderName, mTimeStamp;
    int mCellsSize = 0, mEdgesSize = 0, mMaxEdgesSize = 0, mVertexSize = 0, mTimesteps = 0, mVertLevels = 0, mVertLevelsP1 = 0;
    std::vector<vec3> vertexCoord_vec, cellCoord_vec, edgeCoord_vec;
    std::vector<vec2> vertexLatLon_vec;
    std::vector<size_t> verticesOnCell_vec, verticesOnEdge_vec, cellsOnVertex_vec, cellsOnCell_vec, numberVertexOnCell_vec,
        cellsOnEdge_vec, edgesOnCell_vec;
    std::vector<double> cellRefBottomDepth_vec;
    std::vector<double> cellBottomDepth_vec, cellSurfaceHeight_vec, cellZonalVelocity_vec, cellMeridionalVelocity_vec,
        cellLayerThickness_vec, cellZTop_vec, cellNormalVelocity_vec, cellVertVelocity_vec;
    std::shared_ptr<void> mGroupT; // the open record (lets MPASOSolution::addAttribute read further variables)
};

class MPASOGrid {
public:
    int mCellsSize = 0, mEdgesSize = 0, mMaxEdgesSize = 0, mVertexSize = 0, mTimesteps = 0, mVertLevels = 0, mVertLevelsP1 = 0;
    std::string mMeshName, mCachedDataDir, mFolderPath; // kept for source compatibility; nothing is cached on disk
    std::vector<vec3> vertexCoord_vec, cellCoord_vec, edgeCoord_vec;
    std::vector<vec2> vertexLatLon_vec;
    // 1-based, 0-padded, exactly as in an MPAS file (src/IO/MPASOReader.cpp:147-153)
    std::vector<size_t> verticesOnCell_vec, verticesOnEdge_vec, cellsOnVertex_vec, cellsOnCell_vec, numberVertexOnCell_vec,
        cellsOnEdge_vec, edgesOnCell_vec;
    std::vector<float> cellWeight_vec;
    std::vector<double> cellRefBottomDepth_vec;

    void initGrid(MPASOReader* reader); // src/Core/MPASOGrid.cpp:190-217 (moves the arrays out of the reader)
    void initGrid_DemoLoading(const char* yaml_path); // mesh substream of the YAML, as CLI/main.cpp:104 uses it
    void setGridAttribute(GridAttributeType type, int val);
    void setGridAttributesVec3(GridAttributeType type, const std::vector<vec3>& vec);
    void setGridAttributesVec2(GridAttributeType type, const std::vector<vec2>& vec);
    void setGridAttributesInt(GridAttributeType type, const std::vector<size_t>& vec);
    void setGridAttributesFloat(GridAttributeType type, const std::vector<float>& vec);
    bool checkAttribute();
};

class MPASOSolution {
public:
    std::string mTimeStamp, mDataName;
    int mCellsSize = 0, mEdgesSize = 0, mMaxEdgesSize = 0, mVertexSize = 0, mTimesteps = 0, mVertLevels = 0, mVertLevelsP1 = 0;
    int mTotalZTopLayer = 0, mTotalZTopLayerP1 = 0;
    // cell-major [nCells][nVertLevels] inputs (src/IO/MPASOReader.cpp:215-223)
    std::vector<double> cellLayerThickness_vec, cellZTop_vec, cellVertVelocity_vec, cellNormalVelocity_vec,
        cellMeridionalVelocity_vec, cellZonalVelocity_vec, cellBottomDepth_vec, cellSurfaceHeight_vec;
    std::map<std::string, std::vector<double>> mDoubleAttributes; // e.g. "temperature", "salinity"
    std::shared_ptr<void> gt; // open record handed over by the reader

    void initSolution(MPASOReader* reader); // src/Core/MPASOSolution.cpp:278-320
    // global record index `timestep` of the data substream (all files, in order), as CLI/main.cpp:111 uses it
    void initSolution_DemoLoading(const char* yaml_path, int timestep);
    // read one more cell-major variable of the open record into mDoubleAttributes[name]
    // (src/Core/MPASOSolution.cpp:381-411); float variables are widened to double
    void addAttribute(std::string name, AttributeFormat type);
    void setAttribute(GridAttributeType type, int val);
    void setAttributesDouble(AttributeType type, const std::vector<double>& vec);
    void setTimestep(int timestep) { mTimesteps = timestep; }
    std::string getTimeStamp() const { return mTimeStamp; }
    int getID() const; // FNV-1a of "<timeStamp>_<timestep>", as the reference's MPASOSolution::getID
    bool checkAttribute();
};

class MPASOField {
public:
    std::shared_ptr<MPASOGrid> mGrid;
    std::shared_ptr<MPASOSolution> mSol_Front, mSol_Back;
    // device point location (replaces the serial nanoflann loop, src/Core/MPASOField.cpp:23-34)
    void calcInWhichCells(std::vector<CartesianCoord>& points_vec, std::vector<int>& cell_id_vec);
};

// ---- RGBA-double image (src/Common/ImageBuffer.hpp:13-62) -------------------------------------
template <typename T>
class ImageBuffer {
public:
    ImageBuffer() = default;
    ImageBuffer(int w, int h) : mWidth(w), mHeight(h) { mPixels.resize(static_cast<size_t>(w) * h * 4, T(0)); }
    int getIndex(int i, int j) const { return (i < 0 || i >= mHeight || j < 0 || j >= mWidth) ? -1 : (i * mWidth + j) * 4; }
    void setPixel(int i, int j, const vec3& val)
    {
        const int k = getIndex(i, j);
        if (k < 0) return;
        mPixels[k] = val.x(); mPixels[k + 1] = val.y(); mPixels[k + 2] = val.z(); mPixels[k + 3] = 1.0;
    }
    vec3 getPixel(int i, int j) const
    {
        const int k = getIndex(i, j);
        return k < 0 ? vec3(-1, -1, -1) : vec3(mPixels[k], mPixels[k + 1], mPixels[k + 2]);
    }
    std::vector<T> getChannel(int channel) const
    {
        std::vector<T> out;
        if (channel < 0 || channel > 3) return out;
        out.reserve(static_cast<size_t>(mWidth) * mHeight);
        for (size_t p = 0; p < static_cast<size_t>(mWidth) * mHeight; ++p) out.push_back(mPixels[4 * p + channel]);
        return out;
    }
    int getWidth() const { return mWidth; }
    int getHeight() const { return mHeight; }
    std::vector<T> mPixels;

protected:
    int mWidth = 0, mHeight = 0;
};

// ---- run options / results (src/Core/MPASOVisualizer.h:12-103, field for field) ---------------
enum class CalcPositionType : int { kCenter, kVertx, kPoint, kCount };
enum class CalcAttributeType : int { kZonalMerimoal, kVelocity, kZTop, kTemperature, kSalinity, kAll, kCount };
enum class CalcDirection : int { kForward, kBackward, kCount };
enum class CalcMethodType : int { kRK4, kEuler, kCount };
enum class VisualizeType : int { kFixedLayer, kFixedDepth };
enum class SaveType : int { kVTI, kPNG, kNone, kCount };

struct VisualizationSettings {
    vec2 imageSize; // (width, height)
    vec2 LonRange;
    vec2 LatRange;
    vec2 DepthRange;
    double FixedLatitude = 0.0;
    union {
        double FixedDepth;
        double FixedLayer;
    };
    int tile_index = 0;
    CalcAttributeType CalcType = CalcAttributeType::kZonalMerimoal;
    CalcPositionType PositionType = CalcPositionType::kPoint;
    VisualizeType VisType = VisualizeType::kFixedDepth;
    SaveType saveType = SaveType::kNone;
    double TimeStep = 0.0;
    VisualizationSettings() : FixedDepth(0.0) {}
};

struct SamplingSettings {
    void setSampleRange(const vec2i& number) { sampleRange = number; }
    void setGeoBox(const vec2& latRange, const vec2& lonRange) { sampleLatitudeRange = latRange; sampleLongitudeRange = lonRange; }
    void setDepth(double depth) { sampleDepth = depth; }
    void setSamplingRegion(const vec2i& number, const vec2& latRange, const vec2& lonRange, double depth)
    {
        sampleRange = number; sampleLatitudeRange = latRange; sampleLongitudeRange = lonRange; sampleDepth = depth;
    }
    void atCellCenter(bool b) { bAtCellCenter = b; }
    vec2i getSampleRange() const { return sampleRange; }
    vec2 getLatitudeRange() const { return sampleLatitudeRange; }
    vec2 getLongitudeRange() const { return sampleLongitudeRange; }
    bool isAtCellCenter() const { return bAtCellCenter; }
    double getDepth() const { return sampleDepth; }

private:
    vec2i sampleRange;
    vec2 sampleLatitudeRange, sampleLongitudeRange;
    double sampleDepth = 0.0;
    bool bAtCellCenter = false;
};

struct TrajectoryLine {
    int lineID = 0;
    std::vector<CartesianCoord> points;   // [seed, rec_0 .. rec_each-1]
    std::vector<CartesianCoord> velocity; // [vel_0 .. vel_each-1, 0]
    std::vector<double> temperature, salinity;
    CartesianCoord lastPoint;
    double duration = 0.0, timestamp = 0.0, depth = 0.0;
};

struct TrajectorySettings {
    size_t deltaT = 0, simulationDuration = 0, recordT = 0;
    float depth = 0.0f;
    std::vector<float> particle_depths; // per-particle depth (metres, positive down); used when sized like the seeds
    std::string fileName;
    CalcDirection directionType = CalcDirection::kForward;
    CalcMethodType methodType = CalcMethodType::kEuler; // the reference's default
    bool hasPerParticleDepths() const { return !particle_depths.empty(); }
};

// ---- application state (src/Core/MOPSApp.h) -----------------------------------------------------
enum class MOPSState { Uninitialized, Configuring, Ready };

class MOPSApp {
public:
    MOPSApp();
    ~MOPSApp();
    void init(const char* device);
    void addGrid(std::shared_ptr<MPASOGrid> grid);
    void addSol(int solID, std::shared_ptr<MPASOSolution> sol);
    void addField();
    void activeAttribute(int id1, std::optional<int> id2 = std::nullopt);
    std::vector<TrajectoryLine> runStreamLine(TrajectorySettings* config, std::vector<CartesianCoord>& sample_points);
    std::vector<TrajectoryLine> runPathLine(TrajectorySettings* config, std::vector<CartesianCoord>& sample_points);
    std::vector<ImageBuffer<double>> runRemapping(VisualizationSettings* config);
    // depth-vs-longitude section at config->FixedLatitude; rows span refBottomDepth.front()..back()
    // (src/Core/MOPSApp.cpp:198-210 -> VisualizeFixedLatitude)
    ImageBuffer<double> runReGrid(VisualizationSettings* config);
    // velocity of layer config->FixedLayer (MPASOVisualizer::VisualizeFixedLayer, src/Core/MPASOVisualizer.cpp:17-20)
    ImageBuffer<double> runFixedLayer(VisualizationSettings* config);
    void generateSamplePoints(SamplingSettings* config, std::vector<CartesianCoord>& sample_points);
    void generateSamplePointsAtCenter(SamplingSettings* config, std::vector<CartesianCoord>& sample_points);
    MOPSState getState() const { return mState; }
    void setState(MOPSState s) { mState = s; }
    bool checkAttribute() const;
    std::shared_ptr<MPASOField> getField() const { return mpasoField; }
    ::mops_ctx* engine() const { return mCtx; } // the C-ABI context, for callers that want the flat API
    int locate(const std::vector<CartesianCoord>& pts, std::vector<int>& cells);

    ::mops_multi* multi() const { return mMulti; } // non-null when more than one device is in use (MOPS_DEVICES / "gpu:<n>")

private:
    int residentSlot(int solID);
    int uploadSnapshot(int solID, int slot, bool async);
    int pickSlot(int keepA, int keepB);
    void prefetch(int solID);
    void joinPrefetch();
    MOPSState mState = MOPSState::Uninitialized;
    ::mops_ctx* mCtx = nullptr;      // device 0's context (every single-device call: views, point location)
    ::mops_multi* mMulti = nullptr;  // all devices (trajectory calls shard over them)
    std::thread mPreThread;          // background upload of the next snapshot of a chain
    int mPreSol = -1, mPreRc = 0;
    struct PinBuf { void* p = nullptr; size_t cap = 0; };
    PinBuf mPinned[2];               // page-locked staging of the raw trajectory records (grow-only, reused by every call)
    double* pinnedScratch(int which, size_t bytes);
    std::shared_ptr<MPASOGrid> mpasoGrid;
    std::map<int, std::shared_ptr<MPASOSolution>> mpasoAttributeMap;
    std::shared_ptr<MPASOField> mpasoField;
    int mFrontID = 0, mBackID = 0;
    bool mHasBack = false;
    std::map<int, int> mSlotOf;   // solID -> resident snapshot slot
    std::vector<int> mSlotOrder;  // least recently used first
};

extern MOPSApp app; // pyMOPS reaches into it directly (tools/pyMOPS/bindings.cpp:288,302)

// ---- the API (include/api/MOPS.h:20-148 of the reference) ------------------------------------
void MOPS_Init(const char* device = "gpu");
void MOPS_Begin();
void MOPS_AddGridMesh(std::shared_ptr<MPASOGrid> grid);
void MOPS_AddAttribute(int solID, std::shared_ptr<MPASOSolution> sol);
void MOPS_End();
void MOPS_ActiveAttribute(int t1, std::optional<int> t2 = std::nullopt);
std::vector<ImageBuffer<double>> MOPS_RunRemapping(VisualizationSettings* config);
std::vector<TrajectoryLine> MOPS_RunStreamLine(TrajectorySettings* config, std::vector<CartesianCoord>& sample_points);
std::vector<TrajectoryLine> MOPS_RunPathLine(TrajectorySettings* config, std::vector<CartesianCoord>& sample_points);
void MOPS_GenerateSamplePoints(SamplingSettings* config, std::vector<CartesianCoord>& sample_points);
std::shared_ptr<MPASOField> MOPS_GetFieldSnapshots();
// timing (categories: "IO_Read", "IO_Write", "Preprocessing", "MemoryCopy", "GPUKernel", "CPUCompute", "Other")
void MOPS_ResetTiming();
void MOPS_PrintTimingSummary();
void MOPS_PrintTimingDetailed();
double MOPS_GetCategoryTime(const char* category);
double MOPS_GetTotalTime();

} // namespace MOPS
