// Source-compatibility forwarder: client code written against vm6502q/weed includes "enums/module_type.hpp".
#pragma once
#include "weed_b200/modules.hpp"
