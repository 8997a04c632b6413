// host_design.cpp — see host_design.h.  Plain C++; no CUDA, no dependency on oracle/.
#include "host_design.h"
#include "../../include/rdsp_gpu.h"
#include <cmath>
#include <cstring>
#include <vector>

namespace rdsp_host {

static const double kPi = 3.1415926535897932384626433832795;
static const double kTwoPi = 6.283185307179586476925286766559;
static const double kFs = RDSP_SAMPLE_RATE_HZ;

void design_cplx_fir(double *cI, double *cQ, int n, double lo, double hi, double fs)
{
    const double nFL = lo / fs, nFH = hi / fs;
    const double nFc = (nFH - nFL) / 2.0;          // prototype low-pass cutoff
    const double nFs = kPi * (nFH + nFL);          // 2 pi * centre frequency
    const double fCenter = 0.5 * (double)(n - 1);
    for (int i = 0; i < n; i++) {
        const double x = (float)i - fCenter;
        const double d = (double)i - fCenter;
        double z;
        if ((d > 0 ? d : -d) < 0.01) {
            z = 2.0 * nFc;
        } else {
            const double win = 0.35875 - 0.48829 * cos((kTwoPi * i) / (n - 1))
                             + 0.14128 * cos((2.0 * kTwoPi * i) / (n - 1))
                             - 0.01168 * cos((3.0 * kTwoPi * i) / (n - 1));
            z = (double)sin(kTwoPi * x * nFc) / (kPi * x) * win;
        }
        cI[i] = z * cos(nFs * x);
        cQ[i] = z * sin(nFs * x);
    }
}

void design_mask(double lo, double hi, float *mask)
{
    double cI[RDSP_FIR_TAPS], cQ[RDSP_FIR_TAPS];
    design_cplx_fir(cI, cQ, RDSP_FIR_TAPS, lo, hi, kFs);
    // time-domain buffer exactly as the reference fills it: taps rounded to f32, and the zeroing loop
    // that starts at float index 257 wipes the imaginary part of tap 128 (RDSP_convolutional.h:96-105)
    float t[2 * RDSP_FFT_LEN];
    memset(t, 0, sizeof(t));
    for (int i = 0; i < RDSP_FIR_TAPS; i++) { t[2 * i] = (float)cI[i]; t[2 * i + 1] = (float)cQ[i]; }
    for (int i = RDSP_FFT_LEN + 1; i < 2 * RDSP_FFT_LEN; i++) t[i] = 0.0f;
    // forward DFT evaluated in double and rounded once (the reference uses an f32 FFT; the two agree to f32 rounding)
    for (int k = 0; k < RDSP_FFT_LEN; k++) {
        double re = 0.0, im = 0.0;
        for (int n = 0; n <= RDSP_FIR_TAPS - 1; n++) {
            const int idx = (k * n) & (RDSP_FFT_LEN - 1);
            const double c = cos(kTwoPi * idx / RDSP_FFT_LEN), s = sin(kTwoPi * idx / RDSP_FFT_LEN);
            const double xr = t[2 * n], xi = t[2 * n + 1];
            re += xr * c + xi * s;                 // (xr + j xi) * (c - j s)
            im += xi * c - xr * s;
        }
        mask[2 * k] = (float)re;
        mask[2 * k + 1] = (float)im;
    }
}

static int16_t q15_round(double v)
{
    double r = floor(v * 32768.0 + 0.5);
    if (r > 32767.0) r = 32767.0;
    if (r < -32768.0) r = -32768.0;
    return (int16_t)r;
}

// pass-bands of the shim-defined tap bank (zero-IF): see DESIGN.md "Tap bank"
static const double kHilBand[RDSP_DEMOD_COUNT][2] = {
    {100.0, 3600.0}, {100.0, 3600.0}, {200.0, 1200.0}, {200.0, 1200.0}, {-4500.0, 4500.0}};
static const double kBpBand[RDSP_FILTER_COUNT][2] = {
    {450.0, 950.0}, {150.0, 2100.0}, {150.0, 2700.0}, {150.0, 3100.0}, {150.0, 3900.0}};

// unity gain at the centre of the pass-band (the 129-tap window's main lobe is wider than the CW bands)
static void normalise_centre_gain(double *cI, double *cQ, int n, double lo, double hi, double fs)
{
    const double w = kPi * (hi / fs + lo / fs);
    const double fCenter = 0.5 * (double)(n - 1);
    double gr = 0.0, gi = 0.0;
    for (int k = 0; k < n; k++) {
        const double ang = w * ((double)k - fCenter);
        gr += cI[k] * cos(ang) + cQ[k] * sin(ang);
        gi += cQ[k] * cos(ang) - cI[k] * sin(ang);
    }
    const double g = sqrt(gr * gr + gi * gi);
    for (int k = 0; k < n; k++) { cI[k] = cI[k] / g; cQ[k] = cQ[k] / g; }
}

void design_hilbert_pair(int demod, int16_t *ti, int16_t *tq)
{
    double cI[RDSP_FIR_TAPS], cQ[RDSP_FIR_TAPS];
    const double rs2 = 0.70710678118654752440;
    design_cplx_fir(cI, cQ, RDSP_FIR_TAPS, kHilBand[demod][0], kHilBand[demod][1], kFs);
    normalise_centre_gain(cI, cQ, RDSP_FIR_TAPS, kHilBand[demod][0], kHilBand[demod][1], kFs);
    for (int k = 0; k < RDSP_FIR_TAPS; k++) {
        if (demod == RDSP_DEMOD_AM) {
            ti[k] = q15_round(cI[k]);               // symmetric low-pass on both arms, envelope follows
            tq[k] = q15_round(cI[k]);
        } else {
            ti[k] = q15_round((cI[k] + cQ[k]) * rs2);   // Re / Im of c * e^{-j 45 deg}: a +-45 degree pair
            tq[k] = q15_round((cQ[k] - cI[k]) * rs2);
        }
    }
}

void design_bandpass_hz(double lo, double hi, int16_t *t)
{
    double cI[RDSP_FIR_TAPS], cQ[RDSP_FIR_TAPS];
    design_cplx_fir(cI, cQ, RDSP_FIR_TAPS, lo, hi, kFs);
    normalise_centre_gain(cI, cQ, RDSP_FIR_TAPS, lo, hi, kFs);
    for (int k = 0; k < RDSP_FIR_TAPS; k++) t[k] = q15_round(2.0 * cI[k]);
}

void design_bandpass(int filter, int16_t *t)
{
    design_bandpass_hz(kBpBand[filter][0], kBpBand[filter][1], t);
}

float lms_mu(int strength)
{
    float mu = (float)strength;
    mu /= 2;
    mu += 2;
    mu /= 10;
    mu = powf(10, mu);
    mu = 1 / mu;
    return mu;
}

float agc_alpha(float ms)
{
    return (float)(1.0 - exp(-1.0 / ((double)ms * 1e-3 * kFs)));
}

void make_twiddle_4096_q15(uint32_t *w)
{
    for (int k = 0; k < 3072; k++) {
        const double a = kTwoPi * (double)k / 4096.0;
        double c = floor(cos(a) * 32768.0), s = floor(sin(a) * 32768.0);
        if (c > 32767.0) c = 32767.0;
        if (s > 32767.0) s = 32767.0;
        w[k] = ((uint32_t)(uint16_t)(int16_t)c) | ((uint32_t)(uint16_t)(int16_t)s << 16);
    }
}

void make_twiddle_256_f32(float *cs)
{
    for (int k = 0; k < 256; k++) {
        const double a = kTwoPi * (double)k / 256.0;
        cs[2 * k] = (float)cos(a);
        cs[2 * k + 1] = (float)sin(a);
    }
}

// sinTable_f32 of CMSIS-DSP's arm_sin_f32 / arm_cos_f32: 513 entries, sin(2 pi k / 512) rounded to f32
void make_sin512_f32(float *t513)
{
    for (int k = 0; k <= 512; k++) t513[k] = (float)sin(kTwoPi * (double)k / 512.0);
}

void make_hann_q15(int16_t *w, int n)
{
    for (int i = 0; i < n; i++) {
        double v = floor(32768.0 * 0.5 * (1.0 - cos(kTwoPi * i / (double)(n - 1))) + 0.5);
        w[i] = (int16_t)(v > 32767.0 ? 32767.0 : v);
    }
}

void biquad_highpass_q30(float frequency, float q, int32_t coef[5])
{
    const double w0 = frequency * (2.0f * 3.141592654f / 44100.0f);   // float product, as the Teensy library evaluates it
    const double sinW0 = sin(w0), cosW0 = cos(w0);
    const double alpha = sinW0 / ((double)q * 2.0);
    const double scale = 1073741824.0 / (1.0 + alpha);
    coef[0] = (int32_t)(((1.0 + cosW0) / 2.0) * scale);
    coef[1] = (int32_t)(-(1.0 + cosW0) * scale);
    coef[2] = coef[0];
    coef[3] = -(int32_t)((-2.0 * cosW0) * scale);
    coef[4] = -(int32_t)((1.0 - alpha) * scale);
}

}  // namespace rdsp_host
