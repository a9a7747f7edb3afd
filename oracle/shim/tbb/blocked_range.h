// TEST INFRASTRUCTURE ONLY (oracle build shim): the reference includes this header
// (src/Core/MPASOVisualizer.cpp:8) but uses nothing from it.
#pragma once
