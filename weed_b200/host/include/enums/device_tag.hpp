// Source-compatibility forwarder: client code written against vm6502q/weed includes "enums/device_tag.hpp".
#pragma once
#include "weed_b200/core.hpp"
