// Source-compatibility forwarder: client code written against vm6502q/weed includes "modules/min.hpp".
#pragma once
#include "weed_b200/modules.hpp"
