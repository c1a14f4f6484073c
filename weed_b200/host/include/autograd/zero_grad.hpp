// Source-compatibility forwarder: client code written against vm6502q/weed includes "autograd/zero_grad.hpp".
#pragma once
#include "weed_b200/autograd.hpp"
