// Build shim (test infrastructure): what the reference's cmake would generate
// from include/common/config.h.in for a CPU-only default build
// (WEED_ENABLE_OPENCL=OFF, WEED_BLAS=OFF, defaults from cmake/FpMath.cmake:1,
// cmake/TCapPow.cmake:1, cmake/Pstridepow.cmake:1-2, cmake/CppStd.cmake:1).
#pragma once
#define WEED_ENABLE_ENV_VARS 1
#define WEED_ENABLE_PTHREAD 1
#define WEED_ENABLE_ASYNC 1
#define WEED_FPPOW 5
#define WEED_PSTRIDEPOW 18
#define WEED_TCAPPOW 5
#define WEED_CPP_STD 14
#define WEED_TILE_SIZE 32
