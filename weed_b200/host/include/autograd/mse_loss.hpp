// Source-compatibility forwarder: client code written against vm6502q/weed includes "autograd/mse_loss.hpp".
#pragma once
#include "weed_b200/autograd.hpp"
