// tutorial/streamLine.cpp -- the reference's streamline tutorial (tutorial/streamLine.cpp:12-103 of
// YosefQiu/MOPS) on a synthetic fixture: same API call sequence, BASELINE config C1 parameters
// (100 uniform seeds, depth 800 m, dt = 120 s, 1 day, RK4).
//   usage: streamLine <fixture.bin | stream.yaml> <lines_out.bin> [date-tag]
// With a YAML stream description the grid and the solution come from MPAS NetCDF-3 files through
// MPASOReader, exactly as in the reference tutorial (tutorial/streamLine.cpp:72-103).
#include "api/MOPS.h"
#include "fixture.hpp"

#include <iostream>

int main(int argc, char** argv)
{
    if (argc < 3) {
        std::cerr << "usage: streamLine <fixture.bin> <lines_out.bin>\n";
        return 2;
    }
    const std::string input = argv[1];
    fixture::Loaded fx;
    if (input.size() > 5 && input.substr(input.size() - 5) == ".yaml") {
        const std::string timeStamp = argc > 3 ? argv[3] : "0001-01-01";
        auto grid = std::make_shared<MOPS::MPASOGrid>();
        auto sol = std::make_shared<MOPS::MPASOSolution>();
        sol->initSolution(MOPS::MPASOReader::readSolData(input, timeStamp, 0).get());
        sol->addAttribute("temperature", MOPS::AttributeFormat::kFloat);
        sol->addAttribute("salinity", MOPS::AttributeFormat::kFloat);
        grid->initGrid(MOPS::MPASOReader::readGridData(input).get());
        fx.grid = grid;
        fx.sols.push_back(sol);
    } else {
        fx = fixture::load(input);
    }

    MOPS::MOPS_Init("gpu");
    MOPS::MOPS_Begin();
    MOPS::MOPS_AddGridMesh(fx.grid);
    MOPS::MOPS_AddAttribute(fx.sols[0]->getID(), fx.sols[0]);
    MOPS::MOPS_End();
    MOPS::MOPS_ActiveAttribute(fx.sols[0]->getID());

    std::vector<CartesianCoord> sample_points;
    MOPS::SamplingSettings sampling;
    sampling.setSampleRange(vec2i{11, 11});
    sampling.setGeoBox(vec2{-60.0, 60.0}, vec2{-170.0, 170.0});
    sampling.atCellCenter(false);
    sampling.setDepth(800.0);
    MOPS::MOPS_GenerateSamplePoints(&sampling, sample_points);

    MOPS::TrajectorySettings traj;
    traj.directionType = MOPS::CalcDirection::kForward;
    traj.methodType = MOPS::CalcMethodType::kRK4;
    traj.depth = 800.0f;
    traj.deltaT = ONE_MINUTE * 2;
    traj.simulationDuration = ONE_DAY;
    traj.recordT = ONE_HOUR;

    auto lines = MOPS::MOPS_RunStreamLine(&traj, sample_points);
    std::cout << "seeds " << sample_points.size() << " lines " << lines.size() << " length "
              << (lines.empty() ? 0 : lines[0].points.size()) << std::endl;
    fixture::dump_lines(argv[2], lines);
    MOPS::MOPS_PrintTimingSummary();
    return 0;
}
