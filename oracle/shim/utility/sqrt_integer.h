/* utility/sqrt_integer.h — TEST INFRASTRUCTURE: sqrt_uint32_approx, analyze_fft256iq.cpp:107. */
#ifndef ORACLE_SHIM_SQRT_INTEGER_H
#define ORACLE_SHIM_SQRT_INTEGER_H
#include "../../teensy_shim.h"
static inline uint32_t sqrt_uint32_approx(uint32_t in) { return oracle_sqrt_uint32_approx(in); }
#endif
