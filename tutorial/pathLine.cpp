// tutorial/pathLine.cpp -- the reference's pathline tutorial (tutorial/pathLine.cpp:160-300 of
// YosefQiu/MOPS) on a synthetic fixture: intervals between consecutive snapshots are chained by the
// CALLER -- MOPS_RunPathLine overwrites the seeds with each line's last point, the front/back pair is
// advanced with MOPS_ActiveAttribute, per-particle depths are recomputed from the radius.
//   usage: pathLine <fixture.bin with >= 2 snapshots> <lines_out_prefix>
#include "api/MOPS.h"
#include "fixture.hpp"

#include <cmath>
#include <iostream>

int main(int argc, char** argv)
{
    if (argc < 3) {
        std::cerr << "usage: pathLine <fixture.bin> <out_prefix>\n";
        return 2;
    }
    auto fx = fixture::load(argv[1]);
    if (fx.sols.size() < 2) {
        std::cerr << "pathLine needs at least two snapshots\n";
        return 2;
    }
    MOPS::MOPS_Init("gpu");
    MOPS::MOPS_Begin();
    MOPS::MOPS_AddGridMesh(fx.grid);
    for (size_t s = 0; s < fx.sols.size(); ++s) MOPS::MOPS_AddAttribute((int)(100 + s), fx.sols[s]);
    MOPS::MOPS_End();

    std::vector<CartesianCoord> pts;
    MOPS::SamplingSettings sampling;
    sampling.setSampleRange(vec2i{21, 21});
    sampling.setGeoBox(vec2{-60.0, 60.0}, vec2{-170.0, 170.0});
    sampling.setDepth(800.0);
    MOPS::MOPS_GenerateSamplePoints(&sampling, pts);

    MOPS::TrajectorySettings traj;
    traj.directionType = MOPS::CalcDirection::kForward;
    traj.methodType = MOPS::CalcMethodType::kRK4;
    traj.depth = 800.0f;
    traj.deltaT = ONE_MINUTE * 2;
    traj.simulationDuration = ONE_HOUR * 6;
    traj.recordT = ONE_HOUR;

    for (size_t s = 0; s + 1 < fx.sols.size(); ++s) {
        MOPS::MOPS_ActiveAttribute((int)(100 + s), (int)(100 + s + 1));
        auto lines = MOPS::MOPS_RunPathLine(&traj, pts); // pts <- each line's lastPoint (R14)
        fixture::dump_lines(std::string(argv[2]) + "_" + std::to_string(s) + ".bin", lines);
        // continuation exactly as the reference tutorial does it (tutorial/pathLine.cpp:196-235): the
        // next seed is the last recorded point that is not (0,0,0) -- a line that stopped early has a
        // zero-filled tail -- and the per-particle depth is earthRadius - |x| of that point
        traj.particle_depths.resize(pts.size());
        for (size_t i = 0; i < lines.size(); ++i) {
            const auto& lp = lines[i].points;
            CartesianCoord p = lp.back();
            for (int k = (int)lp.size() - 1; k >= 0; --k) {
                if (!(lp[k].x() == 0.0 && lp[k].y() == 0.0 && lp[k].z() == 0.0)) { p = lp[k]; break; }
            }
            pts[i] = p;
            const double r = std::sqrt(p.x() * p.x() + p.y() * p.y() + p.z() * p.z());
            traj.particle_depths[i] = (float)(6371010.0 - r);
        }
        std::cout << "interval " << s << ": " << lines.size() << " lines" << std::endl;
    }
    MOPS::MOPS_PrintTimingSummary();
    return 0;
}
