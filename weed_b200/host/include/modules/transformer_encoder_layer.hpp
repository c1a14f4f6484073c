// Source-compatibility forwarder: client code written against vm6502q/weed includes "modules/transformer_encoder_layer.hpp".
#pragma once
#include "weed_b200/modules.hpp"
