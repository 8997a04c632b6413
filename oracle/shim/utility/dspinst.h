/* utility/dspinst.h — TEST INFRASTRUCTURE: the one DSP intrinsic analyze_fft256iq.cpp:89 uses. */
#ifndef ORACLE_SHIM_DSPINST_H
#define ORACLE_SHIM_DSPINST_H
#include "../../teensy_shim.h"
static inline int32_t multiply_16tx16t_add_16bx16b(uint32_t a, uint32_t b) { return oracle_smuad(a, b); }
#endif
