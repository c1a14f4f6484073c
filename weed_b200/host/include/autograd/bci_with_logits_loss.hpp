// Source-compatibility forwarder: client code written against vm6502q/weed includes "autograd/bci_with_logits_loss.hpp".
#pragma once
#include "weed_b200/autograd.hpp"
