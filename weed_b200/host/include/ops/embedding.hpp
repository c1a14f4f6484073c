// Source-compatibility forwarder: client code written against vm6502q/weed includes "ops/embedding.hpp".
#pragma once
#include "weed_b200/ops.hpp"
