// TEST INFRASTRUCTURE (oracle/): stand-in for the third-party plotting header `matplotlibcpp.h` (lava/matplotlib-cpp, not vendored
// in the reference, Python/NumPy C API behind it).  The reference's hot-path headers include it (L/Derivatives.cuh:16,
// L/AutonomousRungeKuttaStepper.cuh:9) but only call into it inside #ifdef DEBUG_* blocks that are never defined here.  This
// stand-in declares the namespace (for `namespace plt = matplotlibcpp;`) and pulls in the standard headers the real one
// includes, which the reference's headers rely on transitively (std::vector, std::map, std::string, ...).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <functional>
#include <iostream>
#include <map>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>
namespace matplotlibcpp {}
