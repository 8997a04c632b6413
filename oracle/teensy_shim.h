/*
 * teensy_shim.h — TEST INFRASTRUCTURE (oracle).  Not part of the product path.
 *
 * CPU restatement of the Teensy Audio library pieces the reference uses but
 * does not ship.  The library is a third-party
 * dependency absent from /root/reference (Teensyduino-bundled, unpinned) =>
 * PARITY IS UNPINNED by any reference test; the window and sqrt tables and the
 * biquad update loop are pinned against the shipped firmware image.
 *
 * Call sites in the reference:
 *   sqrt_uint32_approx               analyze_fft256iq.cpp:107
 *   multiply_16tx16t_add_16bx16b     analyze_fft256iq.cpp:89,95
 *   AudioWindowHanning256 / 1024     RadioDSP_SDR_RX.ino:144,147
 *   AudioFilterBiquad::setHighpass   RadioDSP_SDR_RX.ino:155-156
 *   AudioAnalyzeFFT1024              RadioDSP_SDR_RX.ino:58,87,147-148
 */
#ifndef ORACLE_TEENSY_SHIM_H
#define ORACLE_TEENSY_SHIM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

extern const uint16_t sqrt_integer_guess_table[33];
uint32_t oracle_sqrt_uint32_approx(uint32_t in);
int32_t  oracle_smuad(uint32_t a, uint32_t b);       /* multiply_16tx16t_add_16bx16b */

const int16_t *oracle_hanning256(void);              /* AudioWindowHanning256[256]   */
const int16_t *oracle_hanning1024(void);             /* AudioWindowHanning1024[1024] */

/* AudioFilterBiquad, one stage: definition[8] = b0,b1,b2,a1,a2 (Q30, a's negated), bprev, aprev, sum */
typedef struct { int32_t def[8]; } oracle_biquad_t;
void oracle_biquad_set_highpass(oracle_biquad_t *bq, float frequency, float q);
void oracle_biquad_update(oracle_biquad_t *bq, int16_t *data /* 128 samples, in place */);

/* AudioAnalyzeFFT1024 (Appendix A.3) */
typedef struct {
    int16_t  blocks[8][128];     /* blocklist */
    int16_t  buffer[2048];
    uint16_t output[512];
    uint8_t  state;
    uint8_t  outputflag;
} oracle_fft1024_t;
void oracle_fft1024_init(oracle_fft1024_t *f);
void oracle_fft1024_update(oracle_fft1024_t *f, const int16_t *block /* 128 */);

#ifdef __cplusplus
}
#endif
#endif
