// Source-compatibility forwarder: client code written against vm6502q/weed includes "common/config.h".
#pragma once
#include "weed_b200/core.hpp"
