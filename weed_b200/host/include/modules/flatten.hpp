// Source-compatibility forwarder: client code written against vm6502q/weed includes "modules/flatten.hpp".
#pragma once
#include "weed_b200/modules.hpp"
