/* arm_math.h — TEST INFRASTRUCTURE (oracle/_ref build): forwards to the CMSIS primitive shim. */
#ifndef ORACLE_SHIM_ARM_MATH_H
#define ORACLE_SHIM_ARM_MATH_H
#include "Arduino.h"
#include "../cmsis_shim.h"
#endif
