// TEST INFRASTRUCTURE ONLY (oracle build shim) -- stand-in for oneTBB's
// <tbb/parallel_for.h>, which is not installed in this image.  It provides the
// one overload the reference's TBB backend uses: parallel_for(first, last, f).
// Iterations of every reference call site are independent (one particle / pixel /
// vertex-level each), so a serial loop and an OpenMP loop give identical results.
//   -DMOPS_SHIM_OMP : run the range with OpenMP (used for the timed CPU baseline)
//   default         : serial (used for bit-reproducible golden vectors)
#pragma once
#include <cstddef>
namespace tbb {
template <class Index, class Func>
inline void parallel_for(Index first, Index last, const Func& f)
{
#ifdef MOPS_SHIM_OMP
    const long long lo = static_cast<long long>(first);
    const long long hi = static_cast<long long>(last);
#pragma omp parallel for schedule(dynamic, 64)
    for (long long i = lo; i < hi; ++i) f(static_cast<Index>(i));
#else
    for (Index i = first; i < last; ++i) f(i);
#endif
}
} // namespace tbb
