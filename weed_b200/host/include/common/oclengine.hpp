// Source-compatibility forwarder: the reference's tests include "common/oclengine.hpp" under ENABLE_GPU (test/tests.hpp:16-18); on this
// backend the engine singleton is CUDAEngine (WEED_GPU_SINGLETON), with the same method names.
#pragma once
#include "weed_b200/core.hpp"
