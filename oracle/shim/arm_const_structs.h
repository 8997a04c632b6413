/* arm_const_structs.h — TEST INFRASTRUCTURE: arm_cfft_sR_f32_len256 is declared in cmsis_shim.h. */
#ifndef ORACLE_SHIM_ARM_CONST_STRUCTS_H
#define ORACLE_SHIM_ARM_CONST_STRUCTS_H
#include "arm_math.h"
#endif
