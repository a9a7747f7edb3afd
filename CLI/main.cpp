// CLI/main.cpp -- command-line driver with the reference's options and flow (YosefQiu/MOPS CLI/main.cpp:
// -i/--input yaml, -p/--prefix, -t/--timestep, -r/--range, -g/--day, -d/--depth, -h/--help): for every
// requested timestep a fixed-depth remap of the whole globe, then -- for a single timestep -- a streamline
// run from a 31 x 31 seed grid, dumped as TXT in the reference's format (and as legacy-VTK polylines).
// Extra options of this build: --imagesize WxH (default 3601x1801 as the reference), --out DIR, --no-png.
#include "api/MOPS.h"
#include "writers.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace {
struct Options {
    std::string input, prefix, out = ".";
    int timestep = 0, day_gap = 1, width = 3601, height = 1801;
    double depth = 10.0;
    std::vector<int> range;
    bool png = true;
};

void usage(const char* argv0)
{
    std::printf("Usage:\n  %s [OPTION...]\n\n"
                "  -i, --input arg      Input yaml file\n"
                "  -p, --prefix arg     Data path prefix\n"
                "  -t, --timestep arg   single timestep (default: 0)\n"
                "  -r, --range arg      Timestep range (comma separated)\n"
                "  -g, --day arg        Day Gap (default: 1)\n"
                "  -d, --depth arg      Fixed depth (default: 10.0)\n"
                "      --imagesize WxH  remap image size (default: 3601x1801)\n"
                "      --out DIR        output directory (default: .)\n"
                "      --no-png         skip the PNG dumps\n"
                "  -h, --help           Print this information\n", argv0);
}

bool parse(int argc, char** argv, Options& o)
{
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i], v;
        auto eq = a.find('=');
        if (a.rfind("--", 0) == 0 && eq != std::string::npos) { v = a.substr(eq + 1); a = a.substr(0, eq); }
        auto value = [&]() -> std::string {
            if (!v.empty()) return v;
            if (i + 1 >= argc) { std::printf("[ERROR]::Option %s needs a value.\n", a.c_str()); std::exit(1); }
            return argv[++i];
        };
        if (a == "-h" || a == "--help") { usage(argv[0]); return false; }
        else if (a == "-i" || a == "--input") o.input = value();
        else if (a == "-p" || a == "--prefix") o.prefix = value();
        else if (a == "-t" || a == "--timestep") o.timestep = std::atoi(value().c_str());
        else if (a == "-g" || a == "--day") o.day_gap = std::atoi(value().c_str());
        else if (a == "-d" || a == "--depth") o.depth = std::atof(value().c_str());
        else if (a == "-r" || a == "--range") {
            std::stringstream ss(value());
            std::string tok;
            while (std::getline(ss, tok, ',')) if (!tok.empty()) o.range.push_back(std::atoi(tok.c_str()));
        } else if (a == "--imagesize") {
            const std::string s = value();
            if (std::sscanf(s.c_str(), "%dx%d", &o.width, &o.height) != 2) { std::printf("[ERROR]::--imagesize expects WxH\n"); std::exit(1); }
        } else if (a == "--out") o.out = value();
        else if (a == "--no-png") o.png = false;
        else { std::printf("[ERROR]::Unknown option %s\n", a.c_str()); usage(argv[0]); return false; }
    }
    if (o.input.empty()) {
        std::printf("[ERROR]::Input yaml file is required.\n");
        return false;
    }
    return true;
}
} // namespace

int main(int argc, char** argv)
{
    Options opt;
    if (!parse(argc, argv, opt)) return 1;
    std::printf("== command line arguments ==\n== input_yaml_filename: %s\n== data_path_prefix: %s\n== timestep: %d\n== day_gap: %d\n== fixed_depth: %g\n== time_range_vec: ",
                opt.input.c_str(), opt.prefix.c_str(), opt.timestep, opt.day_gap, opt.depth);
    for (int t : opt.range) std::printf("%d ", t);
    std::printf("\n");

    MOPS::MOPS_Init("gpu");
    std::vector<int> timesteps = opt.range.empty() ? std::vector<int>{opt.timestep} : opt.range;

    auto grid = std::make_shared<MOPS::MPASOGrid>();
    grid->initGrid_DemoLoading(opt.input.c_str());
    std::vector<std::shared_ptr<MOPS::MPASOSolution>> sols(timesteps.size());
    for (size_t i = 0; i < timesteps.size(); ++i) {
        sols[i] = std::make_shared<MOPS::MPASOSolution>();
        sols[i]->initSolution_DemoLoading(opt.input.c_str(), timesteps[i]);
        sols[i]->addAttribute("temperature", MOPS::AttributeFormat::kFloat);
        sols[i]->addAttribute("salinity", MOPS::AttributeFormat::kFloat);
    }
    MOPS::MOPS_Begin();
    MOPS::MOPS_AddGridMesh(grid);
    for (size_t i = 0; i < timesteps.size(); ++i) MOPS::MOPS_AddAttribute(timesteps[i], sols[i]);
    MOPS::MOPS_End();

    for (size_t i = 0; i < timesteps.size(); ++i) {
        MOPS::MOPS_ActiveAttribute(timesteps[i]);
        MOPS::VisualizationSettings vis;
        vis.imageSize = vec2{static_cast<double>(opt.width), static_cast<double>(opt.height)};
        vis.LatRange = vec2{-90.0, 90.0};
        vis.LonRange = vec2{-180.0, 180.0};
        vis.FixedDepth = opt.depth;
        vis.TimeStep = timesteps[i];
        vis.saveType = MOPS::SaveType::kPNG;
        auto imgs = MOPS::MOPS_RunRemapping(&vis);
        std::printf("== timestep %d: remap %d x %d, %zu image(s) ==\n", timesteps[i], opt.width, opt.height, imgs.size());
        // raw doubles for downstream tools + the reference's per-channel PNGs
        const std::string raw = opt.out + "/remap_t" + std::to_string(timesteps[i]) + ".bin";
        if (FILE* f = std::fopen(raw.c_str(), "wb")) {
            const int32_t hdr[3] = {static_cast<int32_t>(imgs.size()), opt.width, opt.height};
            std::fwrite(hdr, 4, 3, f);
            for (auto& im : imgs) std::fwrite(im.mPixels.data(), 8, im.mPixels.size(), f);
            std::fclose(f);
        }
        if (opt.png)
            for (size_t k = 0; k < imgs.size(); ++k)
                for (int ch = 0; ch < 3; ++ch)
                    MOPS::writers::SaveToPNG(imgs[k], opt.out + "/output_" + std::to_string(k) + "_ch" + std::to_string(ch) + ".png", ch);
    }

    std::vector<CartesianCoord> seeds;
    std::printf("== generate sample points ==\n");
    MOPS::SamplingSettings sampling;
    sampling.setSampleRange(vec2i{31, 31});
    sampling.setGeoBox(vec2{35.0, 45.0}, vec2{-90.0, -15.0});
    sampling.atCellCenter(false);
    sampling.setDepth(opt.depth);
    MOPS::MOPS_GenerateSamplePoints(&sampling, seeds);

    MOPS::TrajectorySettings traj; // methodType stays at the API default (Euler), as in the reference CLI
    traj.depth = static_cast<float>(opt.depth);
    traj.deltaT = ONE_HOUR * 1;
    traj.simulationDuration = static_cast<size_t>(ONE_DAY) * static_cast<size_t>(opt.day_gap);
    traj.recordT = ONE_HOUR * 6;
    traj.fileName = opt.out + "/traj_line_" + std::to_string(timesteps[0]);
    if (timesteps.size() == 1) {
        std::printf("== single timestep [streamline] ==\n");
        MOPS::MOPS_ActiveAttribute(timesteps[0]);
        auto lines = MOPS::MOPS_RunStreamLine(&traj, seeds);
        MOPS::writers::SaveTrajectoryLinesAsVTK(lines, traj.fileName + ".vtk");
        if (MOPS::writers::SaveTrajectoryLinesAsTXT(lines, traj.fileName + ".txt"))
            std::printf("[ok] Trajectory lines saved to %s.txt\n", traj.fileName.c_str());
        else
            std::fprintf(stderr, "[Error] Unable to open file for writing: %s.txt\n", traj.fileName.c_str());
    }
    MOPS::MOPS_PrintTimingSummary();
    return 0;
}
