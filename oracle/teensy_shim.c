/*
 * teensy_shim.c — TEST INFRASTRUCTURE (oracle).  See teensy_shim.h.
 * Semantics: SURVEY.md Appendix A.3 (documented Teensy Audio behaviour) and
 * Appendix G.2 (AudioFilterBiquad::update as shipped in the firmware image).
 */
#include "teensy_shim.h"
#include "cmsis_shim.h"
#include <math.h>
#include <string.h>

/* utility/sqrt_integer.h — firmware image offset 0x1f558 */
const uint16_t sqrt_integer_guess_table[33] = {
    55109, 38968, 27555, 19484, 13778, 9742, 6889, 4871, 3445, 2436, 1723, 1218, 862, 609, 431, 305,
    216, 153, 108, 77, 54, 39, 27, 20, 14, 10, 7, 5, 4, 3, 2, 1, 0
};

uint32_t oracle_sqrt_uint32_approx(uint32_t in)
{
    if (in == 0) return 0;                      /* ARM UDIV by zero yields 0 */
    uint32_t n = sqrt_integer_guess_table[__builtin_clz(in)];
    n = ((in / n) + n) / 2;
    n = ((in / n) + n) / 2;
    return n;
}

int32_t oracle_smuad(uint32_t a, uint32_t b)
{
    int32_t al = (int16_t)(a & 0xFFFF), ah = (int16_t)(a >> 16);
    int32_t bl = (int16_t)(b & 0xFFFF), bh = (int16_t)(b >> 16);
    return (int32_t)((uint32_t)(ah * bh) + (uint32_t)(al * bl));
}

/* Hann windows: min(32767, round(32768 * 0.5 * (1 - cos(2 pi n / (N-1))))) reproduces the
 * firmware tables at 0x1f2f4 (256) and 0x1eaf4 (1024) exactly (tests/test_oracle_tables.py). */
static int16_t g_hann256[256], g_hann1024[1024];
static int g_win_ready = 0;
static void build_windows(void)
{
    if (g_win_ready) return;
    for (int n = 0; n < 256; n++) {
        double v = floor(32768.0 * 0.5 * (1.0 - cos(2.0 * M_PI * n / 255.0)) + 0.5);
        g_hann256[n] = (int16_t)(v > 32767.0 ? 32767.0 : v);
    }
    for (int n = 0; n < 1024; n++) {
        double v = floor(32768.0 * 0.5 * (1.0 - cos(2.0 * M_PI * n / 1023.0)) + 0.5);
        g_hann1024[n] = (int16_t)(v > 32767.0 ? 32767.0 : v);
    }
    g_win_ready = 1;
}
const int16_t *oracle_hanning256(void) { build_windows(); return g_hann256; }
const int16_t *oracle_hanning1024(void) { build_windows(); return g_hann1024; }

/* ---------------------------------------------------------------- biquad */

void oracle_biquad_set_highpass(oracle_biquad_t *bq, float frequency, float q)
{
    /* AudioFilterBiquad::setHighpass: the w0 product is evaluated in float, the rest in double */
    int coef[5];
    double w0 = frequency * (2.0f * 3.141592654f / 44100.0f);
    double sinW0 = sin(w0);
    double alpha = sinW0 / ((double)q * 2.0);
    double cosW0 = cos(w0);
    double scale = 1073741824.0 / (1.0 + alpha);
    coef[0] = (int)(((1.0 + cosW0) / 2.0) * scale);
    coef[1] = (int)(-(1.0 + cosW0) * scale);
    coef[2] = coef[0];
    coef[3] = (int)((-2.0 * cosW0) * scale);
    coef[4] = (int)((1.0 - alpha) * scale);
    /* setCoefficients(stage 0): feedback coefficients negated, state cleared */
    memset(bq, 0, sizeof(*bq));
    bq->def[0] = coef[0];
    bq->def[1] = coef[1];
    bq->def[2] = coef[2];
    bq->def[3] = coef[3] * -1;
    bq->def[4] = coef[4] * -1;
}

static inline int32_t smlawb(int32_t c, uint32_t x, int32_t acc)
{ return (int32_t)((uint32_t)acc + (uint32_t)(int32_t)(((int64_t)c * (int16_t)(x & 0xFFFF)) >> 16)); }
static inline int32_t smlawt(int32_t c, uint32_t x, int32_t acc)
{ return (int32_t)((uint32_t)acc + (uint32_t)(int32_t)(((int64_t)c * (int16_t)(x >> 16)) >> 16)); }
static inline int32_t ssat16(int32_t v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }

void oracle_biquad_update(oracle_biquad_t *bq, int16_t *data16)
{
    int32_t b0 = bq->def[0], b1 = bq->def[1], b2 = bq->def[2], a1 = bq->def[3], a2 = bq->def[4];
    uint32_t bprev = (uint32_t)bq->def[5], aprev = (uint32_t)bq->def[6];
    int32_t sum = bq->def[7];
    uint32_t *data = (uint32_t *)(void *)data16;
    for (int i = 0; i < 64; i++) {
        uint32_t in2 = data[i];
        sum = smlawb(b0, in2, sum);
        sum = smlawt(b1, bprev, sum);
        sum = smlawb(b2, bprev, sum);
        sum = smlawt(a1, aprev, sum);
        sum = smlawb(a2, aprev, sum);
        uint32_t out2 = (uint32_t)ssat16(sum >> 14) & 0xFFFFu;
        sum &= 0x3FFF;
        sum = smlawt(b0, in2, sum);
        sum = smlawb(b1, in2, sum);
        sum = smlawt(b2, bprev, sum);
        sum = smlawb(a1, out2, sum);
        sum = smlawt(a2, aprev, sum);
        bprev = in2;
        aprev = out2 | ((uint32_t)ssat16(sum >> 14) << 16);
        sum &= 0x3FFF;
        data[i] = aprev;
    }
    bq->def[5] = (int32_t)bprev;
    bq->def[6] = (int32_t)aprev;
    bq->def[7] = sum;            /* single stage: bit 31 ("another stage follows") stays clear */
}

/* ------------------------------------------------------ AudioAnalyzeFFT1024 */

void oracle_fft1024_init(oracle_fft1024_t *f) { memset(f, 0, sizeof(*f)); }

void oracle_fft1024_update(oracle_fft1024_t *f, const int16_t *block)
{
    memcpy(f->blocks[f->state], block, 128 * sizeof(int16_t));
    if (f->state < 7) { f->state++; return; }

    const int16_t *win = oracle_hanning1024();
    for (int b = 0; b < 8; b++)
        for (int i = 0; i < 128; i++) {
            f->buffer[2 * (b * 128 + i)] = f->blocks[b][i];      /* real sample, imaginary 0 */
            f->buffer[2 * (b * 128 + i) + 1] = 0;
        }
    for (int i = 0; i < 1024; i++) {
        int32_t val = (int32_t)f->buffer[2 * i] * win[i];
        f->buffer[2 * i] = (int16_t)(val >> 15);
    }
    arm_cfft_radix4_instance_q15 inst;
    arm_cfft_radix4_init_q15(&inst, 1024, 0, 1);
    arm_cfft_radix4_q15(&inst, f->buffer);
    for (int i = 0; i < 512; i++) {
        uint32_t tmp = ((uint32_t)(uint16_t)f->buffer[2 * i]) | ((uint32_t)(uint16_t)f->buffer[2 * i + 1] << 16);
        uint32_t magsq = (uint32_t)oracle_smuad(tmp, tmp);
        f->output[i] = (uint16_t)oracle_sqrt_uint32_approx(magsq);
    }
    f->outputflag = 1;
    for (int b = 0; b < 4; b++) memcpy(f->blocks[b], f->blocks[b + 4], 128 * sizeof(int16_t));
    f->state = 4;
}
