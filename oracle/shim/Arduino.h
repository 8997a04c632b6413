/* Arduino.h — TEST INFRASTRUCTURE (oracle/_ref build).  Minimal stand-in for the Teensy
 * core so the reference's in-tree DSP sources compile unmodified on x86. */
#ifndef ORACLE_SHIM_ARDUINO_H
#define ORACLE_SHIM_ARDUINO_H
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>
#include <stdlib.h>
#ifdef __cplusplus
/* pull in every C++ header the shim needs BEFORE the abs() macro below exists */
#include <cstdlib>
#include <cmath>
#include <deque>
#include <vector>
#endif

typedef bool boolean;
typedef unsigned long ulong;

/* Arduino's PI / TWO_PI are double macros and win over CMSIS's float PI */
#undef PI
#undef TWO_PI
#define PI      3.1415926535897932384626433832795
#define TWO_PI  6.283185307179586476925286766559
/* type-generic abs, as wiring.h defines it */
#ifdef abs
#undef abs
#endif
#define abs(x) ((x)>0?(x):-(x))

#define AUDIO_BLOCK_SAMPLES      128
#define AUDIO_SAMPLE_RATE_EXACT  44100.0f

static inline void AudioNoInterrupts(void) {}
static inline void AudioInterrupts(void) {}
#endif
