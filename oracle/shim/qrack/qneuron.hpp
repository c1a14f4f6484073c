// Build shim (test infrastructure, not product code).
// /root/reference/include/common/serializer.hpp:19 includes the un-vendored
// third-party header <qrack/qneuron.hpp> unconditionally, but only needs two
// names from it (serializer.hpp:44-49 bitLenInt, :91-99 QNeuronActivationFn).
// This stand-in declares exactly those so the CPU path compiles without Qrack.
#pragma once
#include <cstdint>
typedef uint16_t bitLenInt;
namespace Qrack {
enum QNeuronActivationFn { Sigmoid = 0, ReLU = 1, GeLU = 2, Generalized_Logistic = 3, Leaky_ReLU = 4 };
}
