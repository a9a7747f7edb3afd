// lines.hpp -- flat trajectory buffers -> std::vector<TrajectoryLine> (InitTrajectoryLines / FinalizeTrajectoryLines[WithAttrs] /
// RemoveNaNTrajectoriesAndReindex of the reference, src/Common/TrajectoryCommon.h:43-190).  The trimming itself is
// mops_finalize_lines (C ABI, host only); this header adds the copy into the reference's per-line vectors.  Both steps are
// independent per line and run on all host cores for large seed sets (MOPS_HOST_THREADS overrides the thread count).
#pragma once
#include "api/MOPS.h"
#include "mops_b200.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

namespace MOPS {
namespace detail {

// raw_pos / raw_vel: [n][each][3] as the engine writes them; depths0: the per-particle depths the call started with
inline std::vector<TrajectoryLine> assemble_lines(size_t n, size_t each, const CartesianCoord* seeds, const double* raw_pos,
                                                  const double* raw_vel, bool pathline, double duration, double timestamp,
                                                  const float* depths0)
{
    std::vector<TrajectoryLine> lines;
    if (n == 0 || each == 0) return lines;
    const size_t per = each + 1;
    std::unique_ptr<double[]> pts(new double[n * per * 3]), vel(new double[n * per * 3]), temp(new double[n * per]),
        sal(new double[n * per]), last(new double[n * 3]);
    mops_finalize_lines(static_cast<int64_t>(n), static_cast<int32_t>(each), reinterpret_cast<const double*>(seeds), raw_pos, raw_vel,
                        pathline ? 1 : 0, pts.get(), vel.get(), temp.get(), sal.get(), last.get());
    lines.resize(n);
    auto assemble = [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            TrajectoryLine& ln = lines[i];
            ln.lineID = static_cast<int>(i);
            ln.points.resize(per);
            ln.velocity.resize(per);
            std::memcpy(ln.points.data(), pts.get() + i * per * 3, per * 24);
            std::memcpy(ln.velocity.data(), vel.get() + i * per * 3, per * 24);
            ln.temperature.assign(temp.get() + i * per, temp.get() + (i + 1) * per);
            ln.salinity.assign(sal.get() + i * per, sal.get() + (i + 1) * per);
            ln.lastPoint = CartesianCoord(last[3 * i], last[3 * i + 1], last[3 * i + 2]);
            ln.duration = duration;
            ln.timestamp = timestamp;
            ln.depth = depths0[i];
        }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("MOPS_HOST_THREADS")) nt = static_cast<unsigned>(std::max(1, std::atoi(e)));
    nt = static_cast<unsigned>(std::min<size_t>(std::max(1u, std::min(nt, 64u)), (n * per + (1u << 20) - 1) >> 20));
    if (nt <= 1) {
        assemble(0, n);
    } else {
        std::vector<std::thread> pool;
        const size_t chunk = (n + nt - 1) / nt;
        for (unsigned t = 0; t < nt; ++t) {
            const size_t lo = t * chunk, hi = std::min(n, lo + chunk);
            if (lo < hi) pool.emplace_back(assemble, lo, hi);
        }
        for (auto& th : pool) th.join();
    }
    return lines;
}

} // namespace detail
} // namespace MOPS
