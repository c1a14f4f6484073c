// Source-compatibility forwarder: client code written against vm6502q/weed includes "autograd/cross_entropy_loss.hpp".
#pragma once
#include "weed_b200/autograd.hpp"
