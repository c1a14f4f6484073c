// Source-compatibility forwarder: client code written against vm6502q/weed includes "devices/gpu_device.hpp".
#pragma once
#include "weed_b200/core.hpp"
