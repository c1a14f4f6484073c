// Source-compatibility forwarder: client code written against vm6502q/weed includes "autograd/node.hpp".
#pragma once
#include "weed_b200/core.hpp"
