// Source-compatibility forwarder: client code written against vm6502q/weed includes "autograd/sgd.hpp".
#pragma once
#include "weed_b200/autograd.hpp"
