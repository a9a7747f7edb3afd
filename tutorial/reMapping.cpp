// tutorial/reMapping.cpp -- the reference's remapping tutorial (tutorial/reMapping.cpp:12-60 of
// YosefQiu/MOPS) on a synthetic fixture: lat/lon image at a fixed depth (BASELINE config C2).
//   usage: reMapping <fixture.bin> <image_out.bin> [width height depth]
#include "api/MOPS.h"
#include "fixture.hpp"

#include <iostream>

int main(int argc, char** argv)
{
    if (argc < 3) {
        std::cerr << "usage: reMapping <fixture.bin> <image_out.bin> [width height depth]\n";
        return 2;
    }
    const int w = argc > 3 ? std::atoi(argv[3]) : 360, h = argc > 4 ? std::atoi(argv[4]) : 180;
    const double depth = argc > 5 ? std::atof(argv[5]) : 800.0;
    auto fx = fixture::load(argv[1]);
    MOPS::MOPS_Init("gpu");
    MOPS::MOPS_Begin();
    MOPS::MOPS_AddGridMesh(fx.grid);
    MOPS::MOPS_AddAttribute(fx.sols[0]->getID(), fx.sols[0]);
    MOPS::MOPS_End();
    MOPS::MOPS_ActiveAttribute(fx.sols[0]->getID());

    MOPS::VisualizationSettings vis;
    vis.imageSize = vec2(w, h);
    vis.LatRange = vec2(-90.0, 90.0);
    vis.LonRange = vec2(-180.0, 180.0);
    vis.FixedDepth = depth;
    vis.VisType = MOPS::VisualizeType::kFixedDepth;
    auto imgs = MOPS::MOPS_RunRemapping(&vis);
    std::cout << "images " << imgs.size() << " (" << w << " x " << h << ")" << std::endl;

    FILE* f = std::fopen(argv[2], "wb");
    const int32_t hdr[3] = {(int32_t)imgs.size(), w, h};
    std::fwrite(hdr, 4, 3, f);
    for (auto& im : imgs) std::fwrite(im.mPixels.data(), 8, im.mPixels.size(), f);
    std::fclose(f);
    MOPS::MOPS_PrintTimingSummary();
    return 0;
}
