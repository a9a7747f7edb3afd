// TEST INFRASTRUCTURE ONLY (oracle build shim): CMake normally generates this from
// src/version.h.in; src/Core/MOPSApp.cpp:7,59 only prints it.
#pragma once
#define MOPS_VERSION "reference-oracle"
