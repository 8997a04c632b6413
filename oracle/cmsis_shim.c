/*
 * cmsis_shim.c — TEST INFRASTRUCTURE (oracle).  See cmsis_shim.h for scope,
 * call sites and the "parity unpinned" statement.  Semantics follow SURVEY.md
 * Appendix A.1 (documented CMSIS-DSP behaviour) and Appendix G (q15 radix-4
 * butterflies as shipped in pre_compiled/RadioDSP_SDR_RX.ino.hex).
 */
#include "cmsis_shim.h"
#include <math.h>
#include <string.h>

/* ------------------------------------------------------------------ tables */

static q15_t     g_tw4096_q15[6144];
static float32_t g_tw256_f32[512];
static float32_t g_sin512_f32[513];
static int       g_tables_ready = 0;

static void build_tables(void)
{
    if (g_tables_ready) return;
    /* twiddleCoef_4096_q15: floor(x * 32768) clipped to q15 — reproduces the
     * firmware table at image offset 0x2012c exactly (tests/test_oracle_tables.py) */
    for (int k = 0; k < 3072; k++) {
        double a = 2.0 * M_PI * (double)k / 4096.0;
        double c = floor(cos(a) * 32768.0), s = floor(sin(a) * 32768.0);
        if (c > 32767.0) c = 32767.0;
        if (s > 32767.0) s = 32767.0;
        g_tw4096_q15[2 * k] = (q15_t)c;
        g_tw4096_q15[2 * k + 1] = (q15_t)s;
    }
    /* twiddleCoef_256: (cos,sin)(2*pi*k/256) rounded to f32, firmware offset 0x1f92c */
    for (int k = 0; k < 256; k++) {
        double a = 2.0 * M_PI * (double)k / 256.0;
        g_tw256_f32[2 * k] = (float32_t)cos(a);
        g_tw256_f32[2 * k + 1] = (float32_t)sin(a);
    }
    for (int k = 0; k <= 512; k++)
        g_sin512_f32[k] = (float32_t)sin(2.0 * M_PI * (double)k / 512.0);
    g_tables_ready = 1;
}

const q15_t *oracle_twiddle_4096_q15(void) { build_tables(); return g_tw4096_q15; }
const float32_t *oracle_twiddle_256_f32(void) { build_tables(); return g_tw256_f32; }

const arm_cfft_instance_f32 arm_cfft_sR_f32_len256 = { 256, g_tw256_f32, 0, 0 };

/* ------------------------------------------------------ conversions, copies */

void arm_q15_to_float(const q15_t *pSrc, float32_t *pDst, uint32_t blockSize)
{
    for (uint32_t i = 0; i < blockSize; i++) pDst[i] = (float32_t)pSrc[i] / 32768.0f;
}

void arm_float_to_q15(const float32_t *pSrc, q15_t *pDst, uint32_t blockSize)
{
    /* truncation toward zero, then saturation; no ARM_MATH_ROUNDING (A.1) */
    for (uint32_t i = 0; i < blockSize; i++) {
        float32_t v = pSrc[i] * 32768.0f;
        int32_t q;
        if (v >= 2147483648.0f) q = INT32_MAX;
        else if (v <= -2147483648.0f) q = INT32_MIN;
        else if (v != v) q = 0;
        else q = (int32_t)v;
        if (q > 32767) q = 32767;
        if (q < -32768) q = -32768;
        pDst[i] = (q15_t)q;
    }
}

void arm_copy_f32(const float32_t *pSrc, float32_t *pDst, uint32_t blockSize)
{
    memmove(pDst, pSrc, (size_t)blockSize * sizeof(float32_t));
}

void arm_fill_f32(float32_t value, float32_t *pDst, uint32_t blockSize)
{
    for (uint32_t i = 0; i < blockSize; i++) pDst[i] = value;
}

void arm_cmplx_mult_cmplx_f32(const float32_t *a, const float32_t *b, float32_t *d, uint32_t numSamples)
{
    for (uint32_t i = 0; i < numSamples; i++) {
        float32_t ar = a[2 * i], ai = a[2 * i + 1], br = b[2 * i], bi = b[2 * i + 1];
        d[2 * i] = ar * br - ai * bi;
        d[2 * i + 1] = ar * bi + ai * br;
    }
}

void arm_cmplx_mag_f32(const float32_t *pSrc, float32_t *pDst, uint32_t numSamples)
{
    for (uint32_t i = 0; i < numSamples; i++) {
        float32_t re = pSrc[2 * i], im = pSrc[2 * i + 1];
        pDst[i] = sqrtf(re * re + im * im);
    }
}

/* --------------------------------------------------------------- f32 FFT
 * In-place, interleaved, unnormalised forward DFT e^{-j2pi kn/N}; inverse =
 * conjugate, forward, conjugate, scale 1/N; natural-order output.  Radix-2
 * DIT here; CMSIS uses radix-8/4 — only f32 rounding differs (A.1). */
void arm_cfft_f32(const arm_cfft_instance_f32 *S, float32_t *p, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    build_tables();
    (void)bitReverseFlag;            /* the reference always passes 1 */
    const uint32_t N = S->fftLen;
    const float32_t *tw = S->pTwiddle;
    uint32_t logN = 0;
    while ((1u << logN) < N) logN++;

    if (ifftFlag)
        for (uint32_t i = 0; i < N; i++) p[2 * i + 1] = -p[2 * i + 1];

    for (uint32_t i = 0; i < N; i++) {           /* bit reversal */
        uint32_t r = 0;
        for (uint32_t b = 0; b < logN; b++) r |= ((i >> b) & 1u) << (logN - 1 - b);
        if (r > i) {
            float32_t tr = p[2 * i], ti = p[2 * i + 1];
            p[2 * i] = p[2 * r]; p[2 * i + 1] = p[2 * r + 1];
            p[2 * r] = tr; p[2 * r + 1] = ti;
        }
    }
    for (uint32_t len = 2; len <= N; len <<= 1) {
        uint32_t half = len >> 1, step = N / len;
        for (uint32_t base = 0; base < N; base += len) {
            for (uint32_t j = 0; j < half; j++) {
                float32_t wr = tw[2 * j * step], wi = -tw[2 * j * step + 1];   /* e^{-j theta} */
                uint32_t a = base + j, b = a + half;
                float32_t xr = p[2 * b] * wr - p[2 * b + 1] * wi;
                float32_t xi = p[2 * b] * wi + p[2 * b + 1] * wr;
                float32_t ur = p[2 * a], ui = p[2 * a + 1];
                p[2 * a] = ur + xr; p[2 * a + 1] = ui + xi;
                p[2 * b] = ur - xr; p[2 * b + 1] = ui - xi;
            }
        }
    }
    if (ifftFlag) {
        float32_t invL = 1.0f / (float32_t)N;
        for (uint32_t i = 0; i < N; i++) {
            p[2 * i] = p[2 * i] * invL;
            p[2 * i + 1] = -p[2 * i + 1] * invL;
        }
    }
}

/* 512-entry sine table + linear interpolation (CMSIS fast math) */
float32_t arm_sin_f32(float32_t x)
{
    build_tables();
    float32_t in = x * 0.159154943092f;
    int32_t n = (int32_t)in;
    if (in < 0.0f) n--;
    in = in - (float32_t)n;
    float32_t findex = 512.0f * in;
    uint16_t index = (uint16_t)findex & 0x1ff;
    float32_t fract = findex - (float32_t)((uint16_t)findex);
    float32_t a = g_sin512_f32[index], b = g_sin512_f32[index + 1];
    return (1.0f - fract) * a + fract * b;
}

float32_t arm_cos_f32(float32_t x)
{
    build_tables();
    float32_t in = x * 0.159154943092f + 0.25f;
    int32_t n = (int32_t)in;
    if (in < 0.0f) n--;
    in = in - (float32_t)n;
    float32_t findex = 512.0f * in;
    uint16_t index = (uint16_t)findex & 0x1ff;
    float32_t fract = findex - (float32_t)((uint16_t)findex);
    float32_t a = g_sin512_f32[index], b = g_sin512_f32[index + 1];
    return (1.0f - fract) * a + fract * b;
}

/* ------------------------------------------------------------------ NLMS */

void arm_lms_norm_init_f32(arm_lms_norm_instance_f32 *S, uint16_t numTaps, float32_t *pCoeffs,
                           float32_t *pState, float32_t mu, uint32_t blockSize)
{
    S->numTaps = numTaps;
    S->pCoeffs = pCoeffs;                                   /* coefficients untouched */
    memset(pState, 0, (numTaps + (blockSize - 1u)) * sizeof(float32_t));
    S->pState = pState;
    S->mu = mu;
    S->energy = 0.0f;
    S->x0 = 0.0f;
}

void arm_lms_norm_f32(arm_lms_norm_instance_f32 *S, const float32_t *pSrc, float32_t *pRef,
                      float32_t *pOut, float32_t *pErr, uint32_t blockSize)
{
    float32_t *pState = S->pState;
    float32_t *pCoeffs = S->pCoeffs;
    const uint32_t numTaps = S->numTaps;
    const float32_t mu = S->mu;
    float32_t energy = S->energy, x0 = S->x0;
    float32_t *pStateCurnt = &S->pState[numTaps - 1u];

    for (uint32_t n = 0; n < blockSize; n++) {
        float32_t in = pSrc[n];
        *pStateCurnt++ = in;
        energy -= x0 * x0;
        energy += in * in;
        float32_t sum = 0.0f;
        for (uint32_t k = 0; k < numTaps; k++) sum += pState[k] * pCoeffs[k];   /* sequential f32 */
        pOut[n] = sum;
        float32_t d = pRef[n];
        float32_t e = d - sum;
        pErr[n] = e;
        float32_t w = (e * mu) / (energy + 0.000000119209289f);
        for (uint32_t k = 0; k < numTaps; k++) pCoeffs[k] += w * pState[k];
        x0 = *pState;
        pState++;
    }
    S->energy = energy;
    S->x0 = x0;
    memmove(S->pState, pState, (numTaps - 1u) * sizeof(float32_t));
}

/* ------------------------------------------------------------- q15 SIMD ops
 * 32-bit word = (lo = real, hi = imag); semantics of the ARMv7E-M DSP
 * instructions used by arm_radix4_butterfly_q15 (Appendix G table). */
typedef uint32_t w32;
static inline int32_t LO(w32 a) { return (int16_t)(a & 0xFFFFu); }
static inline int32_t HI(w32 a) { return (int16_t)(a >> 16); }
static inline w32 MK(int32_t lo, int32_t hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }
static inline int32_t sat16(int32_t v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }
static inline w32 SHADD16(w32 a, w32 b) { return MK((LO(a) + LO(b)) >> 1, (HI(a) + HI(b)) >> 1); }
static inline w32 SHSUB16(w32 a, w32 b) { return MK((LO(a) - LO(b)) >> 1, (HI(a) - HI(b)) >> 1); }
static inline w32 QADD16(w32 a, w32 b) { return MK(sat16(LO(a) + LO(b)), sat16(HI(a) + HI(b))); }
static inline w32 QSUB16(w32 a, w32 b) { return MK(sat16(LO(a) - LO(b)), sat16(HI(a) - HI(b))); }
static inline w32 QASX(w32 a, w32 b) { return MK(sat16(LO(a) - HI(b)), sat16(HI(a) + LO(b))); }
static inline w32 QSAX(w32 a, w32 b) { return MK(sat16(LO(a) + HI(b)), sat16(HI(a) - LO(b))); }
static inline w32 SHASX(w32 a, w32 b) { return MK((LO(a) - HI(b)) >> 1, (HI(a) + LO(b)) >> 1); }
static inline w32 SHSAX(w32 a, w32 b) { return MK((LO(a) + HI(b)) >> 1, (HI(a) - LO(b)) >> 1); }
static inline uint32_t SMUAD(w32 c, w32 r)
{ return (uint32_t)(LO(c) * LO(r)) + (uint32_t)(HI(c) * HI(r)); }
static inline uint32_t SMUSDX(w32 c, w32 r)
{ return (uint32_t)(LO(c) * HI(r)) - (uint32_t)(HI(c) * LO(r)); }
/* both products keep their top 16 bits */
static inline w32 CMULPACK(w32 c, w32 r)
{ return (SMUSDX(c, r) & 0xFFFF0000u) | (SMUAD(c, r) >> 16); }

int arm_cfft_radix4_init_q15(arm_cfft_radix4_instance_q15 *S, uint16_t fftLen, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    build_tables();
    S->fftLen = fftLen;
    S->ifftFlag = ifftFlag;
    S->bitReverseFlag = bitReverseFlag;
    S->pTwiddle = g_tw4096_q15;
    switch (fftLen) {
    case 4096: S->twidCoefModifier = 1; S->bitRevFactor = 1; break;
    case 1024: S->twidCoefModifier = 4; S->bitRevFactor = 4; break;
    case 256:  S->twidCoefModifier = 16; S->bitRevFactor = 16; break;
    case 64:   S->twidCoefModifier = 64; S->bitRevFactor = 64; break;
    case 16:   S->twidCoefModifier = 256; S->bitRevFactor = 256; break;
    default: return -1;
    }
    return 0;
}

/* forward transform only (the sketch passes ifftFlag = 0, analyze_fft256iq.h:58) */
void arm_cfft_radix4_q15(const arm_cfft_radix4_instance_q15 *S, q15_t *pSrc16)
{
    const uint32_t N = S->fftLen;
    const w32 *C = (const w32 *)(const void *)S->pTwiddle;    /* word k = (cos, sin)(2 pi k / 4096) */
    w32 *src = (w32 *)(void *)pSrc16;
    uint32_t mod = S->twidCoefModifier;
    uint32_t n1, n2 = N >> 2, ic = 0;

    /* stage 1: inputs scaled by 1/4, stage gain 1/8 */
    for (uint32_t i = 0; i < n2; i++) {
        w32 *p0 = src + i, *p1 = p0 + n2, *p2 = p1 + n2, *p3 = p2 + n2;
        w32 T = SHADD16(SHADD16(*p0, 0), 0);
        w32 Sx = SHADD16(SHADD16(*p2, 0), 0);
        w32 R = QADD16(T, Sx);
        Sx = QSUB16(T, Sx);
        w32 Tb = SHADD16(SHADD16(*p1, 0), 0);
        w32 U = SHADD16(SHADD16(*p3, 0), 0);
        w32 T2 = QADD16(Tb, U);
        *p0 = SHADD16(R, T2);
        R = QSUB16(R, T2);
        *p1 = CMULPACK(C[2 * ic], R);
        T = QSUB16(Tb, U);
        R = QASX(Sx, T);
        Sx = QSAX(Sx, T);
        *p2 = CMULPACK(C[ic], Sx);
        *p3 = CMULPACK(C[3 * ic], R);
        ic += mod;
    }
    mod <<= 2;

    /* middle stages */
    for (uint32_t k = N >> 2; k > 4; k >>= 2) {
        n1 = n2;
        n2 >>= 2;
        ic = 0;
        for (uint32_t j = 0; j < n2; j++) {
            w32 C1 = C[ic], C2 = C[2 * ic], C3 = C[3 * ic];
            ic += mod;
            for (uint32_t i0 = j; i0 < N; i0 += n1) {
                w32 *p0 = src + i0, *p1 = p0 + n2, *p2 = p1 + n2, *p3 = p2 + n2;
                w32 T = *p0, Sx = *p2;
                w32 R = QADD16(T, Sx);
                Sx = QSUB16(T, Sx);
                w32 Tb = *p1, U = *p3;
                T = QADD16(Tb, U);
                *p0 = SHADD16(SHADD16(R, T), 0);
                R = SHSUB16(R, T);
                *p1 = CMULPACK(C2, R);
                T = QSUB16(Tb, U);
                R = SHASX(Sx, T);
                Sx = SHSAX(Sx, T);
                *p2 = CMULPACK(C1, Sx);
                *p3 = CMULPACK(C3, R);
            }
        }
        mod <<= 2;
    }

    /* last stage: no twiddles */
    for (uint32_t i = 0; i < N; i += 4) {
        w32 xa = src[i], xb = src[i + 1], xc = src[i + 2], xd = src[i + 3];
        w32 R = QADD16(xa, xc), T = QADD16(xb, xd);
        w32 Sx = QSUB16(xa, xc), U = QSUB16(xb, xd);
        src[i] = SHADD16(R, T);
        src[i + 1] = SHSUB16(R, T);
        src[i + 2] = SHSAX(Sx, U);
        src[i + 3] = SHASX(Sx, U);
    }

    if (S->bitReverseFlag) {
        uint32_t logN = 0;
        while ((1u << logN) < N) logN++;
        for (uint32_t i = 0; i < N; i++) {
            uint32_t r = 0;
            for (uint32_t b = 0; b < logN; b++) r |= ((i >> b) & 1u) << (logN - 1 - b);
            if (r > i) { w32 t = src[i]; src[i] = src[r]; src[r] = t; }
        }
    }
}

/* ----------------------------------------------------------------- q15 FIR */
void oracle_fir_q15(const q15_t *taps, uint32_t numTaps, const q15_t *hist, const q15_t *x, q15_t *y, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) {
        uint32_t acc = 0;                                    /* wrap-around 32-bit accumulator */
        for (uint32_t k = 0; k < numTaps; k++) {
            int32_t idx = (int32_t)i - (int32_t)k;
            int32_t s = idx >= 0 ? x[idx] : hist[(int32_t)(numTaps - 1) + idx];
            acc += (uint32_t)((int32_t)taps[k] * s);
        }
        y[i] = (q15_t)sat16((int32_t)acc >> 15);
    }
}
