// TEST INFRASTRUCTURE ONLY (oracle build shim) -- netCDF-C is not installed.  The
// reference's reader (src/IO/MPASOReader.cpp:38-119) references these five entry
// points; the oracle never opens a file (meshes go in through the public setters),
// so every call reports failure.
#pragma once
#include <cstddef>
#define NC_NOWRITE 0
#define NC_NOERR 0
inline int nc_open(const char*, int, int* ncid) { if (ncid) *ncid = -1; return -1; }
inline int nc_close(int) { return 0; }
inline int nc_inq_dimid(int, const char*, int* id) { if (id) *id = -1; return -1; }
inline int nc_inq_dimlen(int, int, size_t* len) { if (len) *len = 0; return -1; }
inline const char* nc_strerror(int) { return "netcdf stub (oracle shim)"; }
