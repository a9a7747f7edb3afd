// mops_api.cpp -- host side of the C++ drop-in (include/api/MOPS.h) over the C ABI
// (include/mops_b200.h).  Mirrors the call sequence and error behaviour of the reference's
// src/Core/MOPS.cpp:10-127 and src/Core/MOPSApp.cpp:34-337; every compute step is a C-ABI call
// into libmops_b200.so (CUDA kernels) -- there is no host implementation of the path here.
#include "api/MOPS.h"
#include "mops_b200.h"
#include "lines.hpp"

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <thread>

namespace MOPS {

// =================================================================================================
// timing: the five getters of the reference API (src/Utils/Timer.hpp:17-25, 56-264), fed with
// CUDA-event times reported by the engine plus host wall time of each API call
// =================================================================================================
namespace {
const char* kCategories[] = {"IO_Read", "IO_Write", "Preprocessing", "MemoryCopy", "GPUKernel", "CPUCompute", "Other"};
struct TimerEntry {
    std::string name;
    int category;
    double ms;
};
struct TimerBook {
    std::mutex mu;
    std::vector<TimerEntry> entries;
    void add(const std::string& name, int cat, double ms)
    {
        std::lock_guard<std::mutex> g(mu);
        entries.push_back({name, cat, ms});
    }
    double category(int cat)
    {
        std::lock_guard<std::mutex> g(mu);
        double s = 0.0;
        for (auto& e : entries)
            if (e.category == cat) s += e.ms;
        return s;
    }
};
TimerBook& book()
{
    static TimerBook b;
    return b;
}
int category_index(const char* c)
{
    for (int i = 0; i < 6; ++i)
        if (std::strcmp(c, kCategories[i]) == 0) return i;
    return 6; // anything else is booked as "Other", as the reference does (src/Core/MOPS.cpp:103-118)
}
struct Scope {
    std::string name;
    int cat;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    Scope(std::string n, int c) : name(std::move(n)), cat(c) {}
    ~Scope() { book().add(name, cat, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()); }
};
void engine_error(::mops_ctx* c, const char* where, int rc)
{
    std::fprintf(stderr, "[MOPS/b200]::%s failed (%d): %s\n", where, rc, c ? mops_last_error(c) : "no context");
}
} // namespace

// =================================================================================================
// data model setters (src/Core/MPASOGrid.cpp:82-186, src/Core/MPASOSolution.cpp:1150-1210)
// =================================================================================================
void MPASOGrid::setGridAttribute(GridAttributeType type, int val)
{
    switch (type) {
    case GridAttributeType::kCellSize: mCellsSize = val; break;
    case GridAttributeType::kEdgeSize: mEdgesSize = val; break;
    case GridAttributeType::kVertexSize: mVertexSize = val; break;
    case GridAttributeType::kMaxEdgesSize: mMaxEdgesSize = val; break;
    case GridAttributeType::kVertLevels: mVertLevels = val; break;
    case GridAttributeType::kVertLevelsP1: mVertLevelsP1 = val; break;
    default: std::cout << "[MPASOGrid]::Invalid GridAttributeType" << std::endl; break;
    }
}
void MPASOGrid::setGridAttributesVec3(GridAttributeType type, const std::vector<vec3>& vec)
{
    switch (type) {
    case GridAttributeType::kVertexCoord: vertexCoord_vec = vec; break;
    case GridAttributeType::kCellCoord: cellCoord_vec = vec; break;
    case GridAttributeType::kEdgeCoord: edgeCoord_vec = vec; break;
    default: std::cout << "Error: Invalid GridAttributeType" << std::endl;
    }
}
void MPASOGrid::setGridAttributesVec2(GridAttributeType type, const std::vector<vec2>& vec)
{
    if (type == GridAttributeType::kVertexLatLon) vertexLatLon_vec = vec;
    else std::cout << "Error: Invalid GridAttributeType" << std::endl;
}
void MPASOGrid::setGridAttributesInt(GridAttributeType type, const std::vector<size_t>& vec)
{
    switch (type) {
    case GridAttributeType::kVerticesOnCell: verticesOnCell_vec = vec; break;
    case GridAttributeType::kVerticesOnEdge: verticesOnEdge_vec = vec; break;
    case GridAttributeType::kCellsOnVertex: cellsOnVertex_vec = vec; break;
    case GridAttributeType::kCellsOnCell: cellsOnCell_vec = vec; break;
    case GridAttributeType::kNumberVertexOnCell: numberVertexOnCell_vec = vec; break;
    case GridAttributeType::kCellsOnEdge: cellsOnEdge_vec = vec; break;
    case GridAttributeType::kEdgesOnCell: edgesOnCell_vec = vec; break;
    default: std::cout << "Error: Invalid GridAttributeType" << std::endl;
    }
}
void MPASOGrid::setGridAttributesFloat(GridAttributeType type, const std::vector<float>& vec)
{
    if (type == GridAttributeType::kCellWeight) cellWeight_vec = vec;
    else std::cout << "Error: Invalid GridAttributeType" << std::endl;
}
bool MPASOGrid::checkAttribute()
{
    // the reference's checks that matter for the hot path (src/Core/MPASOGrid.cpp:516-598)
    auto bad = [](const char* what) { std::cout << "[MPASOGrid]::Error: " << what << " is not set" << std::endl; return false; };
    if (mCellsSize == 0) return bad("mCellsSize");
    if (mVertexSize == 0) return bad("mVertexSize");
    if (mMaxEdgesSize == 0) return bad("mMaxEdgesSize");
    if (mVertLevels == 0) return bad("mVertLevels");
    if (mVertLevelsP1 == 0) return bad("mVertLevelsP1");
    if (cellCoord_vec.empty()) return bad("cellCoord_vec");
    if (vertexCoord_vec.empty()) return bad("vertexCoord_vec");
    if (verticesOnCell_vec.empty()) return bad("verticesOnCell_vec");
    if (cellsOnVertex_vec.empty()) return bad("cellsOnVertex_vec");
    if (cellsOnCell_vec.empty()) return bad("cellsOnCell_vec");
    if (numberVertexOnCell_vec.empty()) return bad("numberVertexOnCell_vec");
    return true;
}

void MPASOSolution::setAttribute(GridAttributeType type, int val)
{
    switch (type) {
    case GridAttributeType::kCellSize: mCellsSize = val; break;
    case GridAttributeType::kEdgeSize: mEdgesSize = val; break;
    case GridAttributeType::kVertexSize: mVertexSize = val; break;
    case GridAttributeType::kMaxEdgesSize: mMaxEdgesSize = val; break;
    case GridAttributeType::kVertLevels: mVertLevels = val; break;
    case GridAttributeType::kVertLevelsP1: mVertLevelsP1 = val; break;
    default: std::cerr << "[Error]: Invalid GridAttributeType" << std::endl; break;
    }
}
void MPASOSolution::setAttributesDouble(AttributeType type, const std::vector<double>& vec)
{
    switch (type) {
    case AttributeType::kZTop: cellZTop_vec = vec; break;
    case AttributeType::kLayerThickness: cellLayerThickness_vec = vec; break;
    case AttributeType::kBottomDepth: cellBottomDepth_vec = vec; break;
    case AttributeType::kZonalVelocity: cellZonalVelocity_vec = vec; break;
    case AttributeType::kMeridionalVelocity: cellMeridionalVelocity_vec = vec; break;
    case AttributeType::kNormalVelocity: cellNormalVelocity_vec = vec; break;
    default: std::cerr << "[Error]: Invalid AttributeType" << std::endl; break;
    }
}
int MPASOSolution::getID() const
{
    const std::string key = mTimeStamp + "_" + std::to_string(mTimesteps);
    uint32_t h = 2166136261u;
    for (unsigned char ch : key) h = (h ^ ch) * 16777619u;
    return static_cast<int>(h);
}
bool MPASOSolution::checkAttribute()
{
    if (cellZTop_vec.empty() && cellLayerThickness_vec.empty()) { // src/Core/MPASOSolution.cpp:1222-1226
        std::cerr << "[MPASOSolution]::Error: Invalid ZTop Attribute" << std::endl;
        return false;
    }
    return true;
}

void MPASOField::calcInWhichCells(std::vector<CartesianCoord>& points_vec, std::vector<int>& cell_id_vec)
{
    app.locate(points_vec, cell_id_vec);
}

// =================================================================================================
// MOPSApp
// =================================================================================================
MOPSApp app;

MOPSApp::MOPSApp() = default;
MOPSApp::~MOPSApp()
{
    if (mPreThread.joinable()) mPreThread.join();
    for (auto& b : mPinned) if (b.p) mops_host_free(b.p);
    if (mMulti) mops_multi_destroy(mMulti); // owns every context, device 0's included
    else if (mCtx) mops_destroy(mCtx);
}

namespace {
void segv_backtrace(int sig)
{
    void* frames[64];
    const int n = backtrace(frames, 64);
    const char msg[] = "[MOPS/b200] fatal signal, native backtrace:\n";
    (void)!write(2, msg, sizeof(msg) - 1);
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}
} // namespace

void MOPSApp::init(const char* device)
{
    if (std::getenv("MOPS_BACKTRACE")) { // debugging aid: native frames on SIGSEGV / SIGABRT
        signal(SIGSEGV, segv_backtrace);
        signal(SIGABRT, segv_backtrace);
    }
    std::printf(" [ system information ]\n");
    if (device && std::strcmp(device, "cpu") == 0)
        std::printf("Device requested: cpu -- this build has no CPU path; using the CUDA device\n");
    int dev = 0;
    if (const char* e = std::getenv("MOPS_DEVICE")) dev = std::atoi(e);
    // more than one GPU: MOPS_Init("gpu:<n>") / MOPS_Init("gpu:all") or MOPS_DEVICES=<n>|all.  Trajectory calls then shard their
    // seeds over the devices (mops_multi_*: Morton blocks, NCCL gather in caller order); results do not depend on <n>.
    int n_multi = 1;
    const char* spec = (device && std::strncmp(device, "gpu:", 4) == 0) ? device + 4 : std::getenv("MOPS_DEVICES");
    if (spec && *spec) n_multi = (std::strcmp(spec, "all") == 0) ? 0 : std::atoi(spec);
    if (!mCtx) {
        int rc;
        if (n_multi != 1) {
            rc = mops_multi_create(&mMulti, n_multi, nullptr);
            if (rc == MOPS_OK) mCtx = mops_multi_ctx(mMulti, 0);
        } else {
            rc = mops_create(&mCtx, dev);
        }
        if (rc != MOPS_OK) {
            std::fprintf(stderr, " [ MOPS: no usable CUDA device%s (-> %d); there is no CPU fallback ]\n",
                         n_multi != 1 ? "s / NCCL for the multi-GPU form" : "", rc);
            std::exit(1);
        }
    }
    mops_info info;
    std::memset(&info, 0, sizeof(info));
    mops_get_info(mCtx, &info);
    std::printf("Device selected : CUDA device %d (sm_%d%d, %d SMs) -- B200-native engine\n", info.device, info.cc_major, info.cc_minor,
                info.sm_count);
    if (mMulti) std::printf("Devices in use  : %d (particles shard over them; mesh and snapshots replicated)\n", mops_multi_device_count(mMulti));
    std::printf("MOPS Version    : mops-b200 (ABI %d)\n", mops_abi_version());
    std::fflush(stdout);
    mpasoGrid = std::make_shared<MPASOGrid>();
}

void MOPSApp::addGrid(std::shared_ptr<MPASOGrid> grid)
{
    Scope sc("Preprocessing::addGrid", 2);
    mpasoGrid = std::move(grid);
    if (!mCtx) {
        std::cerr << " [ MOPS is not initialised: call MOPS_Init first ]\n";
        std::exit(1);
    }
    const MPASOGrid& g = *mpasoGrid;
    const size_t nC = static_cast<size_t>(g.mCellsSize), nV = static_cast<size_t>(g.mVertexSize), E = static_cast<size_t>(g.mMaxEdgesSize);
    if (g.cellCoord_vec.size() < nC || g.vertexCoord_vec.size() < nV || g.verticesOnCell_vec.size() < nC * E ||
        g.cellsOnCell_vec.size() < nC * E || g.cellsOnVertex_vec.size() < nV * 3 || g.numberVertexOnCell_vec.size() < nC) {
        std::cerr << "[MOPSApp]::addGrid: grid arrays are smaller than the declared sizes\n";
        std::exit(1);
    }
    auto narrow = [](const std::vector<size_t>& v, size_t n) {
        std::vector<int32_t> o(n);
        for (size_t i = 0; i < n; ++i) o[i] = static_cast<int32_t>(v[i]);
        return o;
    };
    const auto voc = narrow(g.verticesOnCell_vec, nC * E), coc = narrow(g.cellsOnCell_vec, nC * E),
               cov = narrow(g.cellsOnVertex_vec, nV * 3), ne = narrow(g.numberVertexOnCell_vec, nC);
    joinPrefetch();
    const int rc = mMulti ? mops_multi_set_mesh(mMulti, g.mCellsSize, g.mVertexSize, g.mMaxEdgesSize, reinterpret_cast<const double*>(g.cellCoord_vec.data()),
                                                reinterpret_cast<const double*>(g.vertexCoord_vec.data()), voc.data(), coc.data(), cov.data(), ne.data())
                          : mops_set_mesh(mCtx, g.mCellsSize, g.mVertexSize, g.mMaxEdgesSize, reinterpret_cast<const double*>(g.cellCoord_vec.data()),
                                          reinterpret_cast<const double*>(g.vertexCoord_vec.data()), voc.data(), coc.data(), cov.data(), ne.data());
    if (rc != MOPS_OK) {
        if (mMulti) std::fprintf(stderr, "[MOPS/b200] addGrid: %s\n", mops_multi_last_error(mMulti));
        engine_error(mCtx, "addGrid", rc);
        std::exit(1);
    }
    mSlotOf.clear();
    mSlotOrder.clear();
}

void MOPSApp::addSol(int solID, std::shared_ptr<MPASOSolution> sol)
{
    Scope sc("Preprocessing::addSol", 2);
    if (mpasoAttributeMap.count(solID)) return; // a repeated solID is ignored (src/Core/MOPSApp.cpp:82-87)
    sol->mCellsSize = mpasoGrid->mCellsSize;
    sol->mEdgesSize = mpasoGrid->mEdgesSize;
    sol->mMaxEdgesSize = mpasoGrid->mMaxEdgesSize;
    sol->mVertexSize = mpasoGrid->mVertexSize;
    mpasoGrid->mVertLevels = sol->mVertLevels;
    mpasoGrid->mVertLevelsP1 = sol->mVertLevelsP1;
    sol->mTotalZTopLayer = sol->mVertLevels;
    sol->mTotalZTopLayerP1 = sol->mVertLevels + 1;
    mpasoAttributeMap[solID] = std::move(sol);
    // the preprocessing chain of the reference (src/Core/MOPSApp.cpp:100-130) runs on the device
    // when the snapshot becomes resident (residentSlot), not on the host here
}

void MOPSApp::addField()
{
    mpasoField = std::make_shared<MPASOField>();
    mpasoField->mGrid = mpasoGrid;
    mpasoField->mSol_Front = mpasoAttributeMap.begin()->second;
    mFrontID = mpasoAttributeMap.begin()->first;
    mHasBack = false;
}

// a snapshot slot for a new resident: a free one, else the least recently used that is neither keepA nor keepB
int MOPSApp::pickSlot(int keepA, int keepB)
{
    if (static_cast<int>(mSlotOf.size()) < MOPS_MAX_SNAPSHOT_SLOTS) {
        std::vector<bool> used(MOPS_MAX_SNAPSHOT_SLOTS, false);
        for (auto& kv : mSlotOf) used[kv.second] = true;
        for (int s = 0; s < MOPS_MAX_SNAPSHOT_SLOTS; ++s)
            if (!used[s]) return s;
    }
    for (int slot : mSlotOrder) {
        if (slot == keepA || slot == keepB) continue;
        for (auto kv = mSlotOf.begin(); kv != mSlotOf.end(); ++kv)
            if (kv->second == slot) { mSlotOf.erase(kv); break; }
        return slot;
    }
    return -1;
}

// upload + device preprocessing of `solID` into `slot` (replaces the reference's host preprocessing chain of addSol,
// src/Core/MOPSApp.cpp:100-130); async = enqueue on the engine's side stream and return
int MOPSApp::uploadSnapshot(int solID, int slot, bool async)
{
    const MPASOSolution& s = *mpasoAttributeMap.at(solID);
    const size_t nC = static_cast<size_t>(mpasoGrid->mCellsSize), L = static_cast<size_t>(s.mVertLevels);
    if (s.cellZonalVelocity_vec.size() < nC * L || s.cellMeridionalVelocity_vec.size() < nC * L ||
        s.cellLayerThickness_vec.size() < nC * L || s.cellBottomDepth_vec.size() < nC) {
        std::fprintf(stderr, "[MOPSApp]::solution %d: zonal/meridional velocity, layerThickness and bottomDepth are required\n", solID);
        std::exit(1);
    }
    // scalar attributes in std::map (alphabetical) order, first two (R11)
    std::vector<const double*> attrs;
    for (auto& kv : s.mDoubleAttributes) {
        if (attrs.size() >= MOPS_MAX_ATTRS) break;
        if (kv.second.size() >= nC * L) attrs.push_back(kv.second.data());
    }
    const double* wtop = s.cellVertVelocity_vec.size() >= nC * (L + 1) ? s.cellVertVelocity_vec.data() : nullptr;
    const int n_attr = static_cast<int>(attrs.size()), n_total = static_cast<int>(s.mDoubleAttributes.size());
    const double* const* ap = attrs.empty() ? nullptr : attrs.data();
    if (mMulti)
        return mops_multi_set_snapshot(mMulti, slot, s.mVertLevels, s.cellZonalVelocity_vec.data(), s.cellMeridionalVelocity_vec.data(),
                                       s.cellLayerThickness_vec.data(), s.cellBottomDepth_vec.data(), wtop, n_attr, ap, n_total, async ? 1 : 0);
    return async ? mops_set_snapshot_async(mCtx, slot, s.mVertLevels, s.cellZonalVelocity_vec.data(), s.cellMeridionalVelocity_vec.data(),
                                           s.cellLayerThickness_vec.data(), s.cellBottomDepth_vec.data(), wtop, n_attr, ap, n_total)
                 : mops_set_snapshot(mCtx, slot, s.mVertLevels, s.cellZonalVelocity_vec.data(), s.cellMeridionalVelocity_vec.data(),
                                     s.cellLayerThickness_vec.data(), s.cellBottomDepth_vec.data(), wtop, n_attr, ap, n_total);
}

void MOPSApp::joinPrefetch()
{
    if (mPreThread.joinable()) mPreThread.join();
    if (mPreSol >= 0 && mPreRc != MOPS_OK) {
        engine_error(mCtx, "snapshot prefetch", mPreRc);
        std::exit(1);
    }
    mPreSol = -1;
}

// Chained pathlines (tutorial/pathLine.cpp:250-296 of the reference advance a front/back pair per interval): while the
// interval (a, b) integrates, the snapshot that follows b in the attribute map is uploaded and preprocessed in the
// background -- a host thread feeds the engine's side stream (the source arrays are pageable std::vectors, so the staging
// copies block that thread, not the caller), the kernels of the running call are untouched because the target slot is
// neither a's nor b's.  The next MOPS_ActiveAttribute then finds it resident.
void MOPSApp::prefetch(int solID)
{
    if (std::getenv("MOPS_NO_PREFETCH")) return;
    if (mSlotOf.count(solID) || !mpasoAttributeMap.count(solID)) return;
    joinPrefetch();
    const int keepA = mSlotOf.count(mFrontID) ? mSlotOf[mFrontID] : -1, keepB = (mHasBack && mSlotOf.count(mBackID)) ? mSlotOf[mBackID] : -1;
    const int slot = pickSlot(keepA, keepB);
    if (slot < 0) return;
    mSlotOf[solID] = slot;
    mSlotOrder.erase(std::remove(mSlotOrder.begin(), mSlotOrder.end(), slot), mSlotOrder.end());
    mSlotOrder.insert(mSlotOrder.begin(), slot); // least recently used until it is activated
    mPreSol = solID;
    mPreRc = MOPS_OK;
    mPreThread = std::thread([this, solID, slot] { mPreRc = uploadSnapshot(solID, slot, true); });
}

// make `solID` resident in one of the engine's snapshot slots (LRU over MOPS_MAX_SNAPSHOT_SLOTS)
int MOPSApp::residentSlot(int solID)
{
    auto touch = [&](int slot) {
        for (size_t i = 0; i < mSlotOrder.size(); ++i)
            if (mSlotOrder[i] == slot) { mSlotOrder.erase(mSlotOrder.begin() + i); break; }
        mSlotOrder.push_back(slot);
    };
    if (mPreSol == solID) joinPrefetch(); // its upload is enqueued; the engine orders the first use after it
    auto it = mSlotOf.find(solID);
    if (it != mSlotOf.end()) {
        touch(it->second);
        return it->second;
    }
    joinPrefetch();
    const int slot = pickSlot(-1, -1);
    Scope sc("Preprocessing::snapshot", 2);
    const int rc = uploadSnapshot(solID, slot, false);
    if (rc != MOPS_OK) {
        if (mMulti) std::fprintf(stderr, "[MOPS/b200] set_snapshot: %s\n", mops_multi_last_error(mMulti));
        engine_error(mCtx, "set_snapshot", rc);
        std::exit(1);
    }
    mSlotOf[solID] = slot;
    touch(slot);
    return slot;
}

void MOPSApp::activeAttribute(int ID1, std::optional<int> ID2)
{
    mpasoField = std::make_shared<MPASOField>();
    auto it = mpasoAttributeMap.find(ID1);
    if (it == mpasoAttributeMap.end()) {
        std::fprintf(stderr, "[MOPSApp]::activeAttribute: solID %d not found\n", ID1);
        return;
    }
    mpasoField->mGrid = mpasoGrid;
    if (ID2.has_value()) {
        auto it2 = mpasoAttributeMap.find(ID2.value());
        if (it2 == mpasoAttributeMap.end()) {
            std::fprintf(stderr, "[MOPSApp]::activeAttribute: solID %d not found\n", ID2.value());
            return;
        }
        mpasoField->mSol_Front = it->second;
        mpasoField->mSol_Back = it2->second;
        mFrontID = ID1; mBackID = ID2.value(); mHasBack = true;
        residentSlot(mFrontID);
        residentSlot(mBackID);
        auto nxt = mpasoAttributeMap.upper_bound(mBackID); // the pair a chain activates next is (back, the one after it)
        if (nxt != mpasoAttributeMap.end()) prefetch(nxt->first);
    } else {
        mpasoField->mSol_Front = it->second;
        mFrontID = ID1; mHasBack = false;
        residentSlot(mFrontID);
    }
}

// page-locked scratch for the raw records, kept between calls (pinning is slow -- ~1 GB/s -- so a chain of intervals pins
// once); small requests are not worth it and use pageable memory
double* MOPSApp::pinnedScratch(int which, size_t bytes)
{
    if (bytes < (64u << 20) || std::getenv("MOPS_NO_PINNED")) return nullptr;
    PinBuf& b = mPinned[which];
    if (bytes > b.cap) {
        if (b.p) mops_host_free(b.p);
        b.p = nullptr; b.cap = 0;
        void* p = nullptr;
        if (mops_host_alloc(&p, bytes) != MOPS_OK) return nullptr;
        b.p = p; b.cap = bytes;
    }
    return static_cast<double*>(b.p);
}

int MOPSApp::locate(const std::vector<CartesianCoord>& pts, std::vector<int>& cells)
{
    cells.assign(pts.size(), -1);
    if (pts.empty()) return 0;
    static_assert(sizeof(int) == sizeof(int32_t), "int32 cells");
    const int rc = mops_locate(mCtx, MOPS_MEM_HOST, static_cast<int64_t>(pts.size()), reinterpret_cast<const double*>(pts.data()), cells.data());
    if (rc != MOPS_OK) engine_error(mCtx, "locate", rc);
    return rc;
}

namespace {
// TrajectoryCommon.h:29-41: per-particle depths only when sized like the seeds, else the uniform depth
std::vector<float> effective_depths(const TrajectorySettings* cfg, size_t n)
{
    if (cfg->hasPerParticleDepths() && cfg->particle_depths.size() == n) return cfg->particle_depths;
    return std::vector<float>(n, cfg->depth);
}

// raw_pos / raw_vel: caller-provided staging for the flat records ([n][each][3] doubles each), page-locked when the caller
// could get it (the D2H then runs at PCIe speed instead of through the driver's bounce buffer); multi != nullptr shards the
// call over every device
std::vector<TrajectoryLine> run_lines(::mops_ctx* ctx, ::mops_multi* multi, bool path, int slot_f, int slot_b, TrajectorySettings* config,
                                      std::vector<CartesianCoord>& seeds, const char* what,
                                      const std::function<double*(int, size_t)>& scratch)
{
    std::vector<TrajectoryLine> lines;
    if (config == nullptr || seeds.empty()) return lines; // VK:659-665
    if (config->deltaT == 0 || config->recordT == 0 || config->simulationDuration == 0) {
        std::fprintf(stderr, "[B200::%s] invalid trajectory settings\n", what); // VK:666-669
        return lines;
    }
    const size_t n = seeds.size();
    const size_t each = config->simulationDuration / config->recordT;
    const size_t times = config->simulationDuration / config->deltaT;
    std::vector<float> depths = effective_depths(config, n);
    const std::vector<float> depths0 = depths;
    if (each == 0 || times == 0) {
        // VK:709-712 returns the initialised lines: seed only
        std::fprintf(stderr, "[B200::%s] invalid integration steps\n", what);
        lines.resize(n);
        for (size_t i = 0; i < n; ++i) {
            lines[i].lineID = static_cast<int>(i);
            lines[i].points.push_back(seeds[i]);
            lines[i].lastPoint = seeds[i];
            lines[i].duration = static_cast<double>(config->simulationDuration);
            lines[i].timestamp = static_cast<double>(config->deltaT);
            lines[i].depth = depths0[i];
        }
        return lines;
    }
    std::vector<CartesianCoord> pos = seeds; // stable_points
    // the engine writes every slot of both buffers (zeros for the slots a stopped particle never reaches), so they are
    // not value-initialised here: at 1 M seeds x 168 records that alone would be 8 GB of serial zero-fill
    std::unique_ptr<double[]> own_pos, own_vel;
    double* raw_pos = scratch(0, n * each * 3 * sizeof(double));
    double* raw_vel = scratch(1, n * each * 3 * sizeof(double));
    if (!raw_pos || !raw_vel) { // no page-locked memory to be had: pageable buffers work too
        own_pos.reset(new double[n * each * 3]); own_vel.reset(new double[n * each * 3]);
        raw_pos = own_pos.get(); raw_vel = own_vel.get();
    }
    mops_traj_cfg cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.method = (config->methodType == CalcMethodType::kEuler) ? MOPS_METHOD_EULER : MOPS_METHOD_RK4;
    cfg.direction = (config->directionType == CalcDirection::kForward) ? MOPS_DIR_FORWARD : MOPS_DIR_BACKWARD;
    cfg.delta_t = static_cast<int64_t>(config->deltaT);
    cfg.duration = static_cast<int64_t>(config->simulationDuration);
    cfg.record_t = static_cast<int64_t>(config->recordT);
    cfg.mem = MOPS_MEM_HOST;
    cfg.sort_particles = 1;
    mops_traj_io io;
    std::memset(&io, 0, sizeof(io));
    io.n = static_cast<int64_t>(n);
    io.xyz = reinterpret_cast<double*>(pos.data());
    io.depth = depths.data();
    io.cell0 = nullptr; // located on the device (replaces MPASOField::calcInWhichCells)
    io.out_pos = raw_pos;
    io.out_vel = raw_vel;
    mops_traj_stats st;
    int rc;
    if (multi) rc = path ? mops_multi_pathline(multi, &cfg, slot_f, slot_b, &io, &st) : mops_multi_streamline(multi, &cfg, slot_f, &io, &st);
    else rc = path ? mops_pathline(ctx, &cfg, slot_f, slot_b, &io, &st) : mops_streamline(ctx, &cfg, slot_f, &io, &st);
    if (rc != MOPS_OK) {
        if (multi) std::fprintf(stderr, "[MOPS/b200] %s: %s\n", what, mops_multi_last_error(multi));
        engine_error(ctx, what, rc);
        return lines;
    }
    if (st.above_surface_particles > 0) // N2: the reference reads ztop[-1] for these (undefined behaviour); here they stop and keep their last position
        std::fprintf(stderr, "[B200::%s] %lld particle(s) stopped above the interpolated sea surface (depth < -zTop[0])\n", what,
                     static_cast<long long>(st.above_surface_particles));
    book().add(std::string("GPUKernel::") + what + "::kernel", 4, st.kernel_ms);
    book().add(std::string("MemoryCopy::") + what, 3, st.total_ms - st.kernel_ms - st.locate_ms);

    // line assembly + NaN trimming (TrajectoryCommon.h:43-190) on the flat buffers
    lines = detail::assemble_lines(n, each, seeds.data(), raw_pos, raw_vel, path, static_cast<double>(config->simulationDuration),
                                   static_cast<double>(config->deltaT), depths0.data());
    return lines;
}
} // namespace

std::vector<TrajectoryLine> MOPSApp::runStreamLine(TrajectorySettings* config, std::vector<CartesianCoord>& sample_points)
{
    Scope sc("GPUKernel::StreamLine", 6); // whole call booked under "Other"; kernel time goes to GPUKernel
    if (!mpasoField || !mpasoField->mSol_Front) {
        std::fprintf(stderr, "[MOPSApp]::mpasoField is nullptr, please activeAttribute first\n");
        return {};
    }
    return run_lines(mCtx, mMulti, false, residentSlot(mFrontID), -1, config, sample_points, "StreamLine",
                     [this](int w, size_t b) { return pinnedScratch(w, b); });
}

std::vector<TrajectoryLine> MOPSApp::runPathLine(TrajectorySettings* config, std::vector<CartesianCoord>& sample_points)
{
    Scope sc("GPUKernel::PathLine", 6);
    if (!mpasoField) { // src/Core/MOPSApp.cpp:259-264
        std::fprintf(stderr, "[MOPSApp]::mpasoField is nullptr, please activeAttribute first\n");
        std::exit(-1);
    }
    if (!mpasoField->mSol_Front || !mpasoField->mSol_Back || !mHasBack) { // :266-271
        std::fprintf(stderr, "[MOPSApp]::Sol_Front or Sol_Back is nullptr, please activeAttribute first\n");
        std::exit(-1);
    }
    const int sf = residentSlot(mFrontID), sb = residentSlot(mBackID);
    auto lines = run_lines(mCtx, mMulti, true, sf, sb, config, sample_points, "PathLine", [this](int w, size_t b) { return pinnedScratch(w, b); });
    // the caller's seeds become each line's lastPoint (src/Core/MOPSApp.cpp:287-290, R14)
    for (size_t i = 0; i < sample_points.size() && i < lines.size(); ++i) sample_points[i] = lines[i].lastPoint;
    return lines;
}

std::vector<ImageBuffer<double>> MOPSApp::runRemapping(VisualizationSettings* config)
{
    Scope sc("GPUKernel::Remapping", 6);
    std::vector<ImageBuffer<double>> img_vec;
    if (!config || !mpasoField || !mpasoField->mSol_Front) {
        std::fprintf(stderr, "[B200::VisualizeFixedDepth] invalid inputs\n");
        return img_vec;
    }
    const int w = static_cast<int>(config->imageSize.x()), h = static_cast<int>(config->imageSize.y());
    // image 0 = velocity, then ceil(nAttr/3) attribute images (src/Core/MOPSApp.cpp:176-185)
    const size_t attr_size = mpasoField->mSol_Front->mDoubleAttributes.size();
    img_vec.emplace_back(w, h);
    if (attr_size > 0)
        for (size_t i = 0; i < (attr_size + 2) / 3; ++i) img_vec.emplace_back(w, h);
    mops_remap_cfg cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.width = w; cfg.height = h;
    cfg.lat_min = config->LatRange.x(); cfg.lat_max = config->LatRange.y();
    cfg.lon_min = config->LonRange.x(); cfg.lon_max = config->LonRange.y();
    cfg.fixed_depth = config->FixedDepth;
    cfg.mem = MOPS_MEM_HOST;
    mops_remap_stats st;
    const int rc = mops_remap_fixed_depth(mCtx, &cfg, residentSlot(mFrontID), img_vec[0].mPixels.data(),
                                          img_vec.size() > 1 ? img_vec[1].mPixels.data() : nullptr, nullptr, &st);
    if (rc != MOPS_OK) engine_error(mCtx, "runRemapping", rc);
    else {
        book().add("GPUKernel::Remapping::kernel", 4, st.kernel_ms);
        book().add("MemoryCopy::Remapping", 3, st.total_ms - st.kernel_ms);
    }
    return img_vec;
}

namespace {
mops_view_cfg make_view_cfg(const VisualizationSettings* config)
{
    mops_view_cfg cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.width = static_cast<int>(config->imageSize.x());
    cfg.height = static_cast<int>(config->imageSize.y());
    cfg.lat_min = config->LatRange.x(); cfg.lat_max = config->LatRange.y();
    cfg.lon_min = config->LonRange.x(); cfg.lon_max = config->LonRange.y();
    cfg.mem = MOPS_MEM_HOST;
    return cfg;
}
} // namespace

ImageBuffer<double> MOPSApp::runReGrid(VisualizationSettings* config)
{
    Scope sc("GPUKernel::ReGrid", 6);
    const int w = config ? static_cast<int>(config->imageSize.x()) : 0, h = config ? static_cast<int>(config->imageSize.y()) : 0;
    ImageBuffer<double> img(std::max(w, 0), std::max(h, 0));
    if (!config || !mpasoField || !mpasoField->mSol_Front) {
        std::fprintf(stderr, "[B200::VisualizeFixedLatitude] invalid inputs\n");
        return img;
    }
    if (w <= 0 || h <= 0) {
        std::fprintf(stderr, "[B200::VisualizeFixedLatitude] invalid image size\n"); // VK:483-486
        return img;
    }
    if (mpasoGrid->cellRefBottomDepth_vec.empty()) {
        std::fprintf(stderr, "[B200::VisualizeFixedLatitude] refBottomDepth is empty\n"); // VK:491-495
        return img;
    }
    mops_view_cfg cfg = make_view_cfg(config);
    cfg.fixed_latitude = config->FixedLatitude;
    cfg.depth_min = mpasoGrid->cellRefBottomDepth_vec.front();
    cfg.depth_max = mpasoGrid->cellRefBottomDepth_vec.back();
    mops_remap_stats st;
    const int rc = mops_regrid_fixed_latitude(mCtx, &cfg, residentSlot(mFrontID), img.mPixels.data(), nullptr, &st);
    if (rc != MOPS_OK) engine_error(mCtx, "runReGrid", rc);
    else book().add("GPUKernel::ReGrid::kernel", 4, st.kernel_ms);
    return img;
}

ImageBuffer<double> MOPSApp::runFixedLayer(VisualizationSettings* config)
{
    Scope sc("GPUKernel::FixedLayer", 6);
    const int w = config ? static_cast<int>(config->imageSize.x()) : 0, h = config ? static_cast<int>(config->imageSize.y()) : 0;
    ImageBuffer<double> img(std::max(w, 0), std::max(h, 0));
    if (!config || !mpasoField || !mpasoField->mSol_Front || w <= 0 || h <= 0) {
        std::fprintf(stderr, "[B200::VisualizeFixedLayer] invalid inputs\n"); // VK:143-163
        return img;
    }
    mops_view_cfg cfg = make_view_cfg(config);
    cfg.fixed_layer = static_cast<int>(config->FixedLayer); // ClampLayer(int) takes the truncated double
    mops_remap_stats st;
    const int rc = mops_remap_fixed_layer(mCtx, &cfg, residentSlot(mFrontID), img.mPixels.data(), nullptr, &st);
    if (rc != MOPS_OK) engine_error(mCtx, "runFixedLayer", rc);
    else book().add("GPUKernel::FixedLayer::kernel", 4, st.kernel_ms);
    return img;
}

void MOPSApp::generateSamplePoints(SamplingSettings* config, std::vector<CartesianCoord>& points)
{
    // MPASOVisualizer::GenerateSamplePoint (src/Core/MPASOVisualizer.cpp:120-149): floating
    // accumulation loops, then EVERY point of the vector is converted (lon, lat, depth) -> XYZ at
    // r = 6371010 regardless of depth (R15)
    const double minLat = config->getLatitudeRange().x(), maxLat = config->getLatitudeRange().y();
    const double minLon = config->getLongitudeRange().x(), maxLon = config->getLongitudeRange().y();
    const double i_step = (maxLat - minLat) / static_cast<double>(config->getSampleRange().x() - 1);
    const double j_step = (maxLon - minLon) / static_cast<double>(config->getSampleRange().y() - 1);
    for (double i = minLat; i < maxLat; i += i_step)
        for (double j = minLon; j < maxLon; j += j_step) points.push_back(CartesianCoord(j, i, config->getDepth()));
    for (auto& p : points) {
        const double theta = p.y() * (M_PI / 180.0), phi = p.x() * (M_PI / 180.0);
        const double r = 6371010.0f;
        const double costheta = std::cos(theta), cosphi = std::cos(phi), sintheta = std::sin(theta), sinphi = std::sin(phi);
        p = CartesianCoord(r * costheta * cosphi, r * costheta * sinphi, r * sintheta);
    }
}

void MOPSApp::generateSamplePointsAtCenter(SamplingSettings*, std::vector<CartesianCoord>& points)
{
    for (auto& c : mpasoGrid->cellCoord_vec) points.push_back(c); // src/Core/MOPSApp.cpp:217-229
}

bool MOPSApp::checkAttribute() const
{
    if (!mpasoGrid) {
        std::fprintf(stderr, "[MOPSApp]::Grid is nullptr\n");
        return false;
    }
    if (!mpasoGrid->checkAttribute()) {
        std::fprintf(stderr, "[MOPSApp]::Grid attribute check failed\n");
        return false;
    }
    for (auto& kv : mpasoAttributeMap) {
        if (!kv.second) {
            std::fprintf(stderr, "[MOPSApp]::Solution at solID %d is nullptr\n", kv.first);
            return false;
        }
        if (!kv.second->checkAttribute()) {
            std::fprintf(stderr, "[MOPSApp]::Attribute check failed at solID %d\n", kv.first);
            return false;
        }
    }
    return true;
}

// =================================================================================================
// free functions (src/Core/MOPS.cpp:10-127)
// =================================================================================================
void MOPS_Init(const char* device) { app.init(device); }
void MOPS_Begin() { app.setState(MOPSState::Configuring); }
void MOPS_AddGridMesh(std::shared_ptr<MPASOGrid> grid) { app.addGrid(std::move(grid)); }
void MOPS_AddAttribute(int solID, std::shared_ptr<MPASOSolution> sol) { app.addSol(solID, std::move(sol)); }
void MOPS_End()
{
    if (app.getState() != MOPSState::Configuring) {
        std::cerr << " [ MOPS is not configuring ]\n";
        std::exit(1);
    }
    if (!app.checkAttribute()) {
        std::cerr << " [ MOPS is not configured ]\n";
        std::exit(1);
    }
    app.setState(MOPSState::Ready);
    app.addField();
}
void MOPS_ActiveAttribute(int t1, std::optional<int> t2) { app.activeAttribute(t1, t2); }
std::vector<ImageBuffer<double>> MOPS_RunRemapping(VisualizationSettings* config) { return app.runRemapping(config); }
std::vector<TrajectoryLine> MOPS_RunStreamLine(TrajectorySettings* config, std::vector<CartesianCoord>& pts) { return app.runStreamLine(config, pts); }
std::vector<TrajectoryLine> MOPS_RunPathLine(TrajectorySettings* config, std::vector<CartesianCoord>& pts) { return app.runPathLine(config, pts); }
void MOPS_GenerateSamplePoints(SamplingSettings* config, std::vector<CartesianCoord>& pts)
{
    if (!config->isAtCellCenter()) app.generateSamplePoints(config, pts);
    else app.generateSamplePointsAtCenter(config, pts);
}
std::shared_ptr<MPASOField> MOPS_GetFieldSnapshots() { return app.getField(); }

void MOPS_ResetTiming()
{
    std::lock_guard<std::mutex> g(book().mu);
    book().entries.clear();
}
void MOPS_PrintTimingSummary()
{
    std::printf("==== MOPS timing summary (ms) ====\n");
    double total = 0.0;
    for (int c = 0; c < 7; ++c) {
        const double t = book().category(c);
        total += t;
        if (t > 0.0) std::printf("  %-14s %12.3f\n", kCategories[c], t);
    }
    std::printf("  %-14s %12.3f\n", "Total", total);
}
void MOPS_PrintTimingDetailed()
{
    std::lock_guard<std::mutex> g(book().mu);
    std::printf("==== MOPS timing detail (ms) ====\n");
    for (auto& e : book().entries) std::printf("  [%-13s] %-40s %12.3f\n", kCategories[e.category], e.name.c_str(), e.ms);
}
double MOPS_GetCategoryTime(const char* category) { return book().category(category_index(category)); }
double MOPS_GetTotalTime()
{
    double t = 0.0;
    for (int c = 0; c < 7; ++c) t += book().category(c);
    return t;
}

} // namespace MOPS
