// TEST INFRASTRUCTURE ONLY (oracle build shim) -- the single deliberate deviation
// from the unmodified reference.  The reference's TBB remap calls the template
//   SetPixel(Accessor img_acc, ...)            (src/Common/ImageBuffer.hpp:65-75)
// with a std::vector<double> BY VALUE (src/CPU/TBB/Kernel/MPASOVisualizerKernels.cpp:466),
// so every pixel write lands in a temporary copy of the whole image and the returned
// image is all zeros (and the call is O(pixels^2)).  The CUDA kernel of the same
// reference takes a raw double* and has the intended behaviour.  This non-template
// by-reference overload wins overload resolution and restores that behaviour; it is
// force-included (-include) only when compiling that one translation unit.
#pragma once
#include "ggl.h"
#include <vector>
namespace MOPS {
inline void SetPixel(std::vector<double>& img_acc, const int w, const int h, const int i, const int j, const vec3& val)
{
    if (i < 0 || i >= h || j < 0 || j >= w) return;
    const size_t index = (static_cast<size_t>(i) * static_cast<size_t>(w) + static_cast<size_t>(j)) * 4;
    img_acc[index + 0] = val.x();
    img_acc[index + 1] = val.y();
    img_acc[index + 2] = val.z();
    img_acc[index + 3] = 1.0;
}
} // namespace MOPS
