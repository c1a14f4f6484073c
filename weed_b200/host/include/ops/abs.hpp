// Source-compatibility forwarder: client code written against vm6502q/weed includes "ops/abs.hpp".
#pragma once
#include "weed_b200/ops.hpp"
