// Source-compatibility forwarder: client code written against vm6502q/weed includes "ops/real_extremum.hpp".
#pragma once
#include "weed_b200/ops.hpp"
