/*
 * rdsp_oracle.c — TEST INFRASTRUCTURE.  CPU oracle ("port") of the receive chain.
 * See rdsp_oracle.h for who may call this and for the per-stage parity status.
 *
 * Every function cites the reference lines it restates; paths are relative to
 * /root/reference/src/RadioDSP_SDR_RX/ unless they start with backup/.
 * Stages that live in libraries absent from the reference tree (AudioSDR,
 * Teensy Audio) follow SURVEY.md Appendix A.3/A.4 and are the oracle of record
 * for them ("parity unpinned").
 */
#include "rdsp_oracle.h"
#include "cmsis_shim.h"
#include "teensy_shim.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define BLK   RDSP_BLOCK_SAMPLES
#define NTAPS RDSP_FIR_TAPS
#define FFT_L 256

/* Arduino's abs() is a type-generic macro */
#define ARD_ABS(x) ((x) > 0 ? (x) : -(x))

/* ------------------------------------------------------------ filter design */

/* calc_cplx_FIR_coeffs, RDSP_convolutional.h:127-185, FIR_filter_window == 1
 * (4-term Blackman-Harris, :66,153-158).  PI / TWO_PI are Arduino's double macros. */
void rdsp_oracle_calc_cplx_fir(double *cI, double *cQ, int n, double lo, double hi, double fs)
{
    const double PI_ = 3.1415926535897932384626433832795;
    const double TWO_PI_ = 6.283185307179586476925286766559;
    const double FOURPI_ = 2.0 * TWO_PI_, SIXPI_ = 3.0 * TWO_PI_;
    double nFL = lo / fs;
    double nFH = hi / fs;
    double nFc = (nFH - nFL) / 2.0;
    double nFs = PI_ * (nFH + nFL);
    double fCenter = 0.5 * (double)(n - 1);
    for (int i = 0; i < n; i++) { cI[i] = 0.0; cQ[i] = 0.0; }
    for (int i = 0; i < n; i++) {
        double x = (float)i - fCenter;
        double z;
        if (ARD_ABS((double)i - fCenter) < 0.01)
            z = 2.0 * nFc;
        else
            z = (double)sin(TWO_PI_ * x * nFc) / (PI_ * x) *
                (0.35875 - 0.48829 * cos((TWO_PI_ * i) / (n - 1))
                 + 0.14128 * cos((FOURPI_ * i) / (n - 1))
                 - 0.01168 * cos((SIXPI_ * i) / (n - 1)));
        cI[i] = z * cos(nFs * x);
        cQ[i] = z * sin(nFs * x);
    }
}

/* reInitializeFilter + init_filter_mask, RDSP_convolutional.h:209-224,87-110 */
void rdsp_oracle_design_mask(double lo, double hi, float *mask)
{
    double cI[NTAPS], cQ[NTAPS];
    rdsp_oracle_calc_cplx_fir(cI, cQ, NTAPS, lo, hi, RDSP_SAMPLE_RATE_HZ);
    memset(mask, 0, 2 * FFT_L * sizeof(float));          /* static storage starts zeroed */
    for (unsigned i = 0; i < NTAPS; i++) {
        mask[i * 2] = (float)cI[i];
        mask[i * 2 + 1] = (float)cQ[i];
    }
    for (unsigned i = FFT_L + 1; i < FFT_L * 2; i++) mask[i] = 0.0f;   /* :102-105, wipes Q of tap 128 */
    arm_cfft_f32(&arm_cfft_sR_f32_len256, mask, 0, 1);
}

static int16_t q15_round(double v)
{
    double r = floor(v * 32768.0 + 0.5);
    if (r > 32767.0) r = 32767.0;
    if (r < -32768.0) r = -32768.0;
    return (int16_t)r;
}

/* Shim-defined tap bank: every filter is designed with the in-tree
 * designer above and quantised round-to-nearest to q15.
 *   Hilbert pair for a pass-band [lo,hi] of the analytic prototype c = cI + j cQ:
 *     A = (cI + cQ)/sqrt2 (applied to I),  B = (cQ - cI)/sqrt2 (applied to Q)
 *     = real / imaginary part of c * e^{-j45deg}; USB = A(I) - B(Q), LSB = A(I) + B(Q).
 *   AM: A = B = real low-pass (lo = -hi), envelope of (A(I), B(Q)).
 *   Band-pass bank: real taps 2*cI.
 *   Every designed pair is first scaled to unity gain at the centre of its pass-band. */
/* SAM loop (shim-defined; the same constants in radiodsp_sdr_rx_b200/csrc/rdsp_common.cuh): natural frequency 100 Hz,
 * damping 0.707 at 44.1 kHz; pull-in limited to +-1 kHz; carrier-level tracker 100 ms */
#define RDSP_SAM_K1   0.020146f
#define RDSP_SAM_K2   2.02995e-4f
#define RDSP_SAM_WMAX 0.142476f
#define RDSP_SAM_ADC  2.2673e-4f
#define RDSP_SAM_PI   3.14159265358979f

static const double k_hil_band[RDSP_DEMOD_COUNT][2] = {
    { 100.0, 3600.0 },   /* LSB    */
    { 100.0, 3600.0 },   /* USB    */
    { 200.0, 1200.0 },   /* CW_LSB */
    { 200.0, 1200.0 },   /* CW_USB */
    { -4500.0, 4500.0 }, /* AM     */
};
static const double k_bp_band[RDSP_FILTER_COUNT][2] = {
    { 450.0, 950.0 },    /* audioCW   "500 Hz"  */
    { 150.0, 2100.0 },   /* audio2100 */
    { 150.0, 2700.0 },   /* audio2700 */
    { 150.0, 3100.0 },   /* audio3100 */
    { 150.0, 3900.0 },   /* audioAM   "3.9 kHz" */
};

/* scale a designed pair to unity gain at the centre of its pass-band (a 129-tap window cannot reach
 * unity for the narrow CW filters: its main lobe is wider than the band) */
static void normalise_centre_gain(double *cI, double *cQ, int n, double lo, double hi, double fs)
{
    const double w = 3.1415926535897932384626433832795 * (hi / fs + lo / fs);
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

static int16_t g_hil_i[RDSP_DEMOD_COUNT][NTAPS], g_hil_q[RDSP_DEMOD_COUNT][NTAPS], g_bp[RDSP_FILTER_COUNT][NTAPS];
static int g_taps_ready = 0;

static void build_taps(void)
{
    if (g_taps_ready) return;
    double cI[NTAPS], cQ[NTAPS];
    const double rs2 = 0.70710678118654752440;
    for (int m = 0; m < RDSP_DEMOD_COUNT; m++) {
        rdsp_oracle_calc_cplx_fir(cI, cQ, NTAPS, k_hil_band[m][0], k_hil_band[m][1], RDSP_SAMPLE_RATE_HZ);
        normalise_centre_gain(cI, cQ, NTAPS, k_hil_band[m][0], k_hil_band[m][1], RDSP_SAMPLE_RATE_HZ);
        for (int k = 0; k < NTAPS; k++) {
            if (m == RDSP_DEMOD_AM) {
                g_hil_i[m][k] = q15_round(cI[k]);
                g_hil_q[m][k] = q15_round(cI[k]);
            } else {
                g_hil_i[m][k] = q15_round((cI[k] + cQ[k]) * rs2);
                g_hil_q[m][k] = q15_round((cQ[k] - cI[k]) * rs2);
            }
        }
    }
    for (int f = 0; f < RDSP_FILTER_COUNT; f++) {
        rdsp_oracle_calc_cplx_fir(cI, cQ, NTAPS, k_bp_band[f][0], k_bp_band[f][1], RDSP_SAMPLE_RATE_HZ);
        normalise_centre_gain(cI, cQ, NTAPS, k_bp_band[f][0], k_bp_band[f][1], RDSP_SAMPLE_RATE_HZ);
        for (int k = 0; k < NTAPS; k++) g_bp[f][k] = q15_round(2.0 * cI[k]);
    }
    g_taps_ready = 1;
}

static int16_t *taps_ptr(int kind, int index)
{
    build_taps();
    if (kind == RDSP_TAPS_HILBERT_I && index >= 0 && index < RDSP_DEMOD_COUNT) return g_hil_i[index];
    if (kind == RDSP_TAPS_HILBERT_Q && index >= 0 && index < RDSP_DEMOD_COUNT) return g_hil_q[index];
    if (kind == RDSP_TAPS_BANDPASS && index >= 0 && index < RDSP_FILTER_COUNT) return g_bp[index];
    return 0;
}
void rdsp_oracle_get_taps(int kind, int index, int16_t *t)
{ int16_t *p = taps_ptr(kind, index); if (p) memcpy(t, p, NTAPS * sizeof(int16_t)); }
void rdsp_oracle_set_taps(int kind, int index, const int16_t *t)
{ int16_t *p = taps_ptr(kind, index); if (p) memcpy(p, t, NTAPS * sizeof(int16_t)); }

/* mu from the "DSP strength" setting, RDSP_noise_reduction.h:48-56 */
float rdsp_oracle_lms_mu(int strength)
{
    float mu_calc = strength;
    mu_calc /= 2;
    mu_calc += 2;
    mu_calc /= 10;
    mu_calc = powf(10, mu_calc);
    mu_calc = 1 / mu_calc;
    return mu_calc;
}

/* ------------------------------------------------------------ NLMS instance
 * One copy of the globals of RDSP_noise_reduction.h:26-32 plus the function
 * statics of :69 (which Init_LMS_NR does NOT reset). */
typedef struct {
    float errsig[256 + 10];
    arm_lms_norm_instance_f32 inst;
    float state[RDSP_LMS_TAPS + 128];
    float coeff[RDSP_LMS_TAPS + 128];
    float delay[256 + 128];
    unsigned long inbuf, outbuf;
} nlms_t;

/* Init_LMS_NR, RDSP_noise_reduction.h:35-64 — coefficients are NOT cleared */
static void nlms_init(nlms_t *s, int strength)
{
    uint16_t calc_taps = RDSP_LMS_TAPS;
    float mu_calc = rdsp_oracle_lms_mu(strength);
    s->inst.numTaps = calc_taps;
    s->inst.pCoeffs = s->coeff;
    s->inst.pState = s->state;
    s->inst.mu = mu_calc;
    arm_fill_f32(0.0f, s->delay, 256 + 128);
    arm_fill_f32(0.0f, s->state, calc_taps + 128);
    arm_lms_norm_init_f32(&s->inst, calc_taps, &s->coeff[0], &s->state[0], mu_calc, 128);
}

/* LMS_NoiseReduction, RDSP_noise_reduction.h:66-80 */
static void nlms_run(nlms_t *s, int blockSize, float *nrbuffer)
{
    arm_copy_f32(nrbuffer, &s->delay[s->inbuf], blockSize);
    arm_lms_norm_f32(&s->inst, nrbuffer, &s->delay[s->outbuf], nrbuffer, s->errsig, blockSize);
    s->inbuf += blockSize;
    s->outbuf = s->inbuf + blockSize;
    s->inbuf %= 256;
    s->outbuf %= 256;
}

/* ----------------------------------------------------------- channel state */
struct rdsp_oracle_chan {
    rdsp_gpu_config_t cfg;
    rdsp_chan_params_t par;
    /* K0 */
    int32_t mult_i, mult_q;
    /* K1/K2 q15 delay lines: the 128 samples before the current block */
    int16_t hist_i[BLK], hist_q[BLK], hist_d[BLK];
    /* SAM carrier PLL: phase, frequency (rad/sample), carrier level */
    float sam_phi, sam_omega, sam_dc;
    /* noise blanker: running IQ magnitude (|I| + |Q| averaged over 32-sample chunks) */
    int32_t nb_ref;
    /* K3 notch */
    nlms_t notch;
    int notch_old_level;
    /* K4 AGC */
    float agc_env, agc_alpha_a, agc_alpha_d[4];
    /* K5 FFT filter, RDSP_convolutional.h:42,50-56,77 */
    uint8_t first_block;
    float last_L[BLK], last_R[BLK];
    float mask[2 * FFT_L];
    /* K6 DNR */
    nlms_t dnr;
    int old_nr_level;                    /* oldNRLevel, RDSP_convolutional.h:80 */
    /* K8 spectral NR */
    float nfloor;                        /* NFloor, backup/RDSP_convolutional_spec.h:109 */
    /* a11 + K9 */
    oracle_biquad_t bq_i, bq_q;
    int have_prev;
    int16_t prev_i[BLK], prev_q[BLK];
    uint32_t sum[256];
    uint8_t count, naverage, outputflag;
    uint16_t output[256];
    arm_cfft_radix4_instance_q15 fft_inst;
    /* K10 */
    oracle_fft1024_t fft1024;
    /* K11 */
    uint16_t spectrum_view[256], spectrum_view_old[256];
    uint16_t waterfall[50][128];             /* WaterfallData[MAX_WATERFALL][..], RDSP_display.h:30; cols 0..127 used */
};

/* setup() defaults, RadioDSP_SDR_RX.ino:117-148,183 (the oracle keeps its own copy so that it
 * does not link against the product library) */
void rdsp_oracle_default_params(rdsp_chan_params_t *p)
{
    memset(p, 0, sizeof(*p));
    p->demod = RDSP_DEMOD_LSB;
    p->audio_filter = RDSP_FILTER_2700;
    p->agc_mode = RDSP_AGC_MEDIUM;
    p->notch_on = 0;
    p->notch_level = 20;
    p->nr_kind = RDSP_NR_OFF;
    p->nr_level = 0;
    p->pbt_lo_hz = 300.0f;
    p->pbt_hi_hz = 4000.0f;
    p->in_gain = 1.0f;
    p->out_gain = 0.5f;
    p->iq_balance = 1.020f;
    p->als_peak = 0;
    p->nb_on = 0;
    p->nb_threshold_db = 20.0f;
}

void rdsp_oracle_default_config(rdsp_gpu_config_t *cfg)
{
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = sizeof(*cfg);
    cfg->n_channels = 1;
    cfg->stage_mask = RDSP_STAGE_ALL;
    cfg->max_blocks_per_call = 1;
    cfg->spec256_naverage = 30;
    cfg->agc_target = 0.25f;
    cfg->agc_max_gain = 1000.0f;
    cfg->agc_attack_ms = 5.0f;
    cfg->agc_decay_ms[RDSP_AGC_FAST] = 100.0f;
    cfg->agc_decay_ms[RDSP_AGC_MEDIUM] = 500.0f;
    cfg->agc_decay_ms[RDSP_AGC_SLOW] = 2000.0f;
}

static float agc_alpha(float ms)
{
    return (float)(1.0 - exp(-1.0 / ((double)ms * 1e-3 * RDSP_SAMPLE_RATE_HZ)));
}

static void apply_params(rdsp_oracle_chan_t *c, const rdsp_chan_params_t *p, int first)
{
    /* K0: AudioMixer4 gain convention, mult = gain * 65536 */
    c->mult_i = (int32_t)((double)p->in_gain * 65536.0);
    c->mult_q = (int32_t)((double)p->in_gain * (double)p->iq_balance * 65536.0);
    if (first || p->pbt_lo_hz != c->par.pbt_lo_hz || p->pbt_hi_hz != c->par.pbt_hi_hz)
        rdsp_oracle_design_mask((double)p->pbt_lo_hz, (double)p->pbt_hi_hz, c->mask);
    c->par = *p;
}

rdsp_oracle_chan_t *rdsp_oracle_chan_create(const rdsp_gpu_config_t *cfg)
{
    rdsp_oracle_chan_t *c = (rdsp_oracle_chan_t *)calloc(1, sizeof(*c));
    if (!c) return 0;
    c->cfg = *cfg;
    build_taps();
    c->first_block = 1;
    c->old_nr_level = 15;                              /* RDSP_convolutional.h:80 */
    nlms_init(&c->dnr, 15);                            /* Init_LMS_NR(15), RadioDSP_SDR_RX.ino:172 */
    c->notch_old_level = -1;
    c->agc_alpha_a = agc_alpha(cfg->agc_attack_ms);
    for (int m = 1; m < 4; m++) c->agc_alpha_d[m] = agc_alpha(cfg->agc_decay_ms[m]);
    oracle_biquad_set_highpass(&c->bq_i, 500, 0.5f);   /* RadioDSP_SDR_RX.ino:155-156 */
    oracle_biquad_set_highpass(&c->bq_q, 500, 0.5f);
    c->naverage = (uint8_t)(cfg->spec256_naverage ? cfg->spec256_naverage : 1);   /* analyze_fft256iq.h:88-91 */
    arm_cfft_radix4_init_q15(&c->fft_inst, 256, 0, 1); /* analyze_fft256iq.h:58 */
    oracle_fft1024_init(&c->fft1024);
    c->spectrum_view[0] = 1;                           /* = {1}, RDSP_display.h:31-32 */
    c->spectrum_view_old[0] = 1;
    rdsp_chan_params_t p;
    rdsp_oracle_default_params(&p);
    apply_params(c, &p, 1);
    return c;
}

void rdsp_oracle_chan_destroy(rdsp_oracle_chan_t *c) { free(c); }

int rdsp_oracle_chan_set_mode(rdsp_oracle_chan_t *c, const rdsp_chan_params_t *p)
{
    apply_params(c, p, 0);
    return 0;
}

void rdsp_oracle_chan_get_mask(rdsp_oracle_chan_t *c, float *m) { memcpy(m, c->mask, sizeof(c->mask)); }
void rdsp_oracle_chan_set_mask(rdsp_oracle_chan_t *c, const float *m) { memcpy(c->mask, m, sizeof(c->mask)); }

/* ------------------------------------------------------------------ stages */

static inline int32_t sat16(int32_t v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }

/* AudioMixer4 gain: unity = pass, else ((int64)mult * in) >> 16 saturated (A.4) */
static inline int16_t mix_gain(int16_t in, int32_t mult)
{
    if (mult == 65536) return in;
    int64_t v = ((int64_t)mult * (int64_t)in) >> 16;
    if (v > 32767) v = 32767;
    if (v < -32768) v = -32768;
    return (int16_t)v;
}

/* K0+K1+K2 (shim-defined AudioSDR front end, all q15) */
static void stage_frontend(rdsp_oracle_chan_t *c, const int16_t *iq, int16_t *audio)
{
    int16_t xi[BLK], xq[BLK], a[BLK], b[BLK], d[BLK];
    const int mode = c->par.demod;
    const int m = mode == RDSP_DEMOD_SAM ? RDSP_DEMOD_AM : mode;         /* SAM filters its arms with the AM rows */
    for (int n = 0; n < BLK; n++) {
        xi[n] = mix_gain(iq[2 * n], c->mult_i);
        xq[n] = mix_gain(iq[2 * n + 1], c->mult_q);
    }
    if (c->par.nb_on) {
        /* Noise blanker (SDR.enableNoiseBlanker, RadioDSP_SDR_RX.ino:129-131; AudioSDR absent, shim-defined, integer):
         * a frame whose |I| + |Q| exceeds threshold x the running magnitude is zeroed; the running magnitude is the mean
         * of the (clipped) magnitudes of a 32-sample chunk, smoothed over 8 chunks; the first chunk only seeds it. */
        const uint32_t mult_q8 = (uint32_t)(pow(10.0, (double)c->par.nb_threshold_db / 20.0) * 256.0 + 0.5);
        for (int k = 0; k < BLK; k += 32) {
            const int armed = c->nb_ref > 0;
            const uint32_t thr = armed ? (uint32_t)(((uint32_t)c->nb_ref * mult_q8) >> 8) : 0xFFFFFFFFu;
            uint32_t sum = 0;
            for (int n = k; n < k + 32; n++) {
                const uint32_t mag = (uint32_t)abs((int)xi[n]) + (uint32_t)abs((int)xq[n]);
                if (mag > thr) { xi[n] = 0; xq[n] = 0; sum += thr; } else sum += mag;
            }
            const int32_t cm = (int32_t)(sum >> 5);
            c->nb_ref = armed ? c->nb_ref + ((cm - c->nb_ref) >> 3) : cm;
        }
    }
    oracle_fir_q15(g_hil_i[m], NTAPS, c->hist_i, xi, a, BLK);
    oracle_fir_q15(g_hil_q[m], NTAPS, c->hist_q, xq, b, BLK);
    memcpy(c->hist_i, xi, sizeof(xi));
    memcpy(c->hist_q, xq, sizeof(xq));
    for (int n = 0; n < BLK; n++) {
        switch (mode) {
        case RDSP_DEMOD_SAM: {
            /* SAMmode (RDSP_controls.h:384-391; AudioSDR absent, shim-defined): second-order PLL on the carrier of
             * the low-passed pair, coherent detection, carrier level removed by a slow tracker.  f32. */
            float sn = sinf(c->sam_phi), cs = cosf(c->sam_phi);
            float re = (float)a[n] * cs + (float)b[n] * sn;
            float im = (float)b[n] * cs - (float)a[n] * sn;
            float err = (re == 0.0f && im == 0.0f) ? 0.0f : atan2f(im, re);
            c->sam_omega += RDSP_SAM_K2 * err;
            c->sam_omega = fminf(fmaxf(c->sam_omega, -RDSP_SAM_WMAX), RDSP_SAM_WMAX);
            c->sam_phi += c->sam_omega + RDSP_SAM_K1 * err;
            if (c->sam_phi >= RDSP_SAM_PI) c->sam_phi -= 2.0f * RDSP_SAM_PI;
            if (c->sam_phi < -RDSP_SAM_PI) c->sam_phi += 2.0f * RDSP_SAM_PI;
            c->sam_dc += (re - c->sam_dc) * RDSP_SAM_ADC;
            float o = re - c->sam_dc;
            o = fminf(fmaxf(o, -32768.0f), 32767.0f);
            d[n] = (int16_t)(int32_t)o;                                           /* truncation toward zero */
            break;
        }
        case RDSP_DEMOD_USB: case RDSP_DEMOD_CW_USB:
            d[n] = (int16_t)sat16((int32_t)a[n] - (int32_t)b[n]); break;          /* QSUB16 */
        case RDSP_DEMOD_LSB: case RDSP_DEMOD_CW_LSB:
            d[n] = (int16_t)sat16((int32_t)a[n] + (int32_t)b[n]); break;          /* QADD16 */
        default: {                                                                 /* AM envelope */
            uint32_t e = oracle_sqrt_uint32_approx((uint32_t)((int32_t)a[n] * a[n]) + (uint32_t)((int32_t)b[n] * b[n]));
            d[n] = (int16_t)(e > 32767u ? 32767u : e);
        } }
    }
    oracle_fir_q15(g_bp[c->par.audio_filter], NTAPS, c->hist_d, d, audio, BLK);
    memcpy(c->hist_d, d, sizeof(d));
}

/* K3: ALS auto-notch = error output of the NLMS structure of RDSP_noise_reduction.h:35-80 */
static void stage_notch(rdsp_oracle_chan_t *c, float *x)
{
    if (c->par.notch_level != c->notch_old_level) {
        nlms_init(&c->notch, c->par.notch_level);
        c->notch_old_level = c->par.notch_level;
    }
    nlms_run(&c->notch, BLK, x);                       /* leaves the estimate in x, the error in errsig */
    if (!c->par.als_peak) memcpy(x, c->notch.errsig, BLK * sizeof(float));   /* notch: error; peak: estimate */
}

/* K4: AGC (peak follower, shim-defined) then SDR.setOutputGain, RadioDSP_SDR_RX.ino:134 */
static void stage_agc(rdsp_oracle_chan_t *c, float *x)
{
    const int mode = c->par.agc_mode;
    const float target = c->cfg.agc_target, max_gain = c->cfg.agc_max_gain;
    const float knee = target / max_gain;
    const float aa = c->agc_alpha_a, ad = c->agc_alpha_d[mode];
    float env = c->agc_env;
    for (int n = 0; n < BLK; n++) {
        float v = x[n];
        if (mode != RDSP_AGC_OFF) {
            float mag = fabsf(v);
            float diff = mag - env;
            env = env + (diff > 0.0f ? aa : ad) * diff;
            float g = env > knee ? target / env : max_gain;
            v = v * g;
        }
        x[n] = v * c->par.out_gain;
    }
    c->agc_env = env;
}

/* K8: spectral subtraction, backup/RDSP_convolutional_spec.h:181-238, with the
 * loop bounds restated as FFT_length (the shipped `j < FFT_length*2` overruns its
 * arrays). */
static void spectral_subtract(rdsp_oracle_chan_t *c, const float *fft_buf, float *ifft_buf, float level)
{
    float mag[FFT_L];
    float beta = 0.65;
    arm_cmplx_mag_f32(fft_buf, mag, FFT_L);
    float specVal = 0.0;
    for (int m = 30; m <= 180; m++) specVal = specVal + mag[m];
    float th = specVal / (180 - 30);
    th = th * (level * 1.5);
    c->nfloor += (th - c->nfloor) * beta;
    c->nfloor = (c->nfloor > 0) ? c->nfloor : 0;
    for (int j = 0; j < FFT_L; j++) {
        if (mag[j] <= c->nfloor) mag[j] = mag[j] * 0.2;
        else mag[j] = mag[j] - c->nfloor;
    }
    for (int j = 0; j < FFT_L; j++) {
        float r1 = fft_buf[2 * j], i1 = fft_buf[2 * j + 1];
        float phi = atan2f(i1, r1);
        ifft_buf[2 * j] = mag[j] * arm_cos_f32(phi);
        ifft_buf[2 * j + 1] = mag[j] * arm_sin_f32(phi);
    }
}

/* K6 + the post-NR gain: the DNR branch of doConvolutionalProcessing, RDSP_convolutional.h:326-337 */
static void stage_dnr(rdsp_oracle_chan_t *c, float *float_buffer_L, float *float_buffer_R)
{
    if (c->par.nr_level != c->old_nr_level) {
        nlms_init(&c->dnr, c->par.nr_level);
        c->old_nr_level = c->par.nr_level;
    }
    nlms_run(&c->dnr, 128, float_buffer_L);
    for (int i = 0; i < BLK; i++) {
        float_buffer_L[i] = float_buffer_L[i] * 1.1;                 /* double multiply, :334 */
        float_buffer_R[i] = float_buffer_L[i];
    }
}

/* test hook: K6 alone on f32 blocks x [n_blocks][128] (in the chain: the L output of K5), y = what the chain emits.
 * Lets a test hand the DNR of both sides IDENTICAL inputs: K6 multiplies a 1e-7 difference in its input (two FFT
 * algorithms) by ~1e3 in the two blocks after its same-block-reference first call. */
void rdsp_oracle_chan_dnr_f32(rdsp_oracle_chan_t *c, uint32_t n_blocks, const float *x, float *y)
{
    float L[BLK], R[BLK];
    for (uint32_t b = 0; b < n_blocks; b++) {
        memcpy(L, x + (size_t)b * BLK, sizeof(L));
        if (c->par.nr_kind == RDSP_NR_LMS && c->par.nr_level > 0) stage_dnr(c, L, R);
        memcpy(y + (size_t)b * BLK, L, sizeof(L));
    }
}

/* K5 + K6/K8: doConvolutionalProcessing body, RDSP_convolutional.h:250-337 */
static void stage_conv(rdsp_oracle_chan_t *c, const int16_t *sp_L, const int16_t *sp_R, float *out_L, float *out_R)
{
    float float_buffer_L[BLK], float_buffer_R[BLK];
    float FFT_buffer[2 * FFT_L], iFFT_buffer[2 * FFT_L];
    arm_q15_to_float(sp_L, float_buffer_L, BLK);                     /* :241-242 */
    arm_q15_to_float(sp_R, float_buffer_R, BLK);
    memset(FFT_buffer, 0, sizeof(FFT_buffer));
    if (c->first_block) {                                            /* :256-263 */
        c->first_block = 0;
    } else {
        for (unsigned i = 0; i < BLK; i++) {                         /* :267-271 */
            FFT_buffer[i * 2] = c->last_L[i];
            FFT_buffer[i * 2 + 1] = c->last_R[i];
        }
    }
    for (unsigned i = 0; i < BLK; i++) {                             /* :274-285 */
        c->last_L[i] = float_buffer_L[i];
        c->last_R[i] = float_buffer_R[i];
        FFT_buffer[FFT_L + i * 2] = float_buffer_L[i];
        FFT_buffer[FFT_L + i * 2 + 1] = float_buffer_R[i];
    }
    arm_cfft_f32(&arm_cfft_sR_f32_len256, FFT_buffer, 0, 1);          /* :291 */

    const int nr_on = (c->cfg.stage_mask & RDSP_STAGE_NR) != 0;
    const int kind = nr_on ? c->par.nr_kind : RDSP_NR_OFF;
    if (kind == RDSP_NR_SPECTRAL && c->par.nr_level > 0)
        spectral_subtract(c, FFT_buffer, iFFT_buffer, (float)c->par.nr_level);
    else
        arm_cmplx_mult_cmplx_f32(FFT_buffer, c->mask, iFFT_buffer, FFT_L);   /* :301 */
    arm_cfft_f32(&arm_cfft_sR_f32_len256, iFFT_buffer, 1, 1);         /* :309 */
    for (unsigned i = 0; i < FFT_L / 2; i++) {                       /* :314-318 */
        float_buffer_L[i] = iFFT_buffer[FFT_L + i * 2];
        float_buffer_R[i] = iFFT_buffer[FFT_L + i * 2 + 1];
    }
    if (kind == RDSP_NR_LMS && c->par.nr_level > 0) stage_dnr(c, float_buffer_L, float_buffer_R);
    memcpy(out_L, float_buffer_L, sizeof(float_buffer_L));
    memcpy(out_R, float_buffer_R, sizeof(float_buffer_R));
}

/* K9: AudioAnalyzeFFT256IQ::update, analyze_fft256iq.cpp:65-118, on one block per port */
static void spec256_update(rdsp_oracle_chan_t *c, const int16_t *bi, const int16_t *bqv)
{
    int16_t buffer[512] __attribute__((aligned(4)));
    if (!c->have_prev) {                                             /* :73-77 */
        memcpy(c->prev_i, bi, BLK * sizeof(int16_t));
        memcpy(c->prev_q, bqv, BLK * sizeof(int16_t));
        c->have_prev = 1;
        return;
    }
    uint32_t *dst = (uint32_t *)(void *)buffer;                      /* copy_to_fft_buffer, :38-48 */
    for (int i = 0; i < BLK; i++) dst[i] = (uint16_t)c->prev_i[i] | ((uint32_t)(uint16_t)c->prev_q[i] << 16);
    for (int i = 0; i < BLK; i++) dst[128 + i] = (uint16_t)bi[i] | ((uint32_t)(uint16_t)bqv[i] << 16);
    const int16_t *win = oracle_hanning256();                        /* apply_window_to_fft_buffer, :50-63 */
    for (int i = 0; i < 256; i++) {
        buffer[2 * i] = (int16_t)((buffer[2 * i] * win[i]) >> 15);
        buffer[2 * i + 1] = (int16_t)((buffer[2 * i + 1] * win[i]) >> 15);
    }
    arm_cfft_radix4_q15(&c->fft_inst, buffer);                       /* :82 */
    for (int i = 0; i < 256; i++) {                                  /* :86-98 */
        uint32_t tmp = dst[i];
        uint32_t magsq = (uint32_t)oracle_smuad(tmp, tmp);
        if (c->count == 0) c->sum[i] = magsq / c->naverage;
        else c->sum[i] += magsq / c->naverage;
    }
    if (++c->count == c->naverage) {                                 /* :99-113 */
        c->count = 0;
        for (int i = 0; i < 256; i++) c->output[255 - (i ^ 128)] = (uint16_t)oracle_sqrt_uint32_approx(c->sum[i]);
        c->outputflag = 1;
    }
    memcpy(c->prev_i, bi, BLK * sizeof(int16_t));                    /* :114-117 */
    memcpy(c->prev_q, bqv, BLK * sizeof(int16_t));
}

/* test hook: K9 alone on already separated I / Q blocks (no biquads), to pin it against the compiled reference */
void rdsp_oracle_chan_spec256_raw(rdsp_oracle_chan_t *c, const int16_t *i_blk, const int16_t *q_blk)
{
    spec256_update(c, i_blk, q_blk);
}

/* a11 + K9: HP biquads (RadioDSP_SDR_RX.ino:75-78,155-156) feeding the IQ spectrum */
static void stage_spec256(rdsp_oracle_chan_t *c, const int16_t *iq)
{
    int16_t bi[BLK], bqv[BLK];
    for (int n = 0; n < BLK; n++) { bi[n] = iq[2 * n]; bqv[n] = iq[2 * n + 1]; }
    oracle_biquad_update(&c->bq_i, bi);
    oracle_biquad_update(&c->bq_q, bqv);
    spec256_update(c, bi, bqv);
}

/* one AudioStream tick of the graph wired at RadioDSP_SDR_RX.ino:71-89 */
static void tick(rdsp_oracle_chan_t *c, const int16_t *iq, int16_t *audio, float *f32)
{
    const uint32_t sm = c->cfg.stage_mask;
    int16_t L[BLK], R[BLK];
    float fL[BLK], fR[BLK];

    if (sm & RDSP_STAGE_SPEC256) stage_spec256(c, iq);

    if (sm & RDSP_STAGE_FRONTEND) {
        int16_t m[BLK];
        stage_frontend(c, iq, m);
        const int do_notch = (sm & RDSP_STAGE_NOTCH) && c->par.notch_on;
        if (do_notch || (sm & RDSP_STAGE_AGC)) {
            float x[BLK];
            arm_q15_to_float(m, x, BLK);
            if (do_notch) stage_notch(c, x);
            if (sm & RDSP_STAGE_AGC) stage_agc(c, x);
            memcpy(fL, x, sizeof(x));
            arm_float_to_q15(x, m, BLK);
        } else {
            arm_q15_to_float(m, fL, BLK);
        }
        memcpy(L, m, sizeof(m));
        memcpy(R, m, sizeof(m));                                     /* SDR outputs 0 and 1 carry the same audio */
        memcpy(fR, fL, sizeof(fL));
    } else {
        for (int n = 0; n < BLK; n++) { L[n] = iq[2 * n]; R[n] = iq[2 * n + 1]; }
        arm_q15_to_float(L, fL, BLK);
        arm_q15_to_float(R, fR, BLK);
    }

    if (sm & RDSP_STAGE_FFTFILT) {
        stage_conv(c, L, R, fL, fR);
        arm_float_to_q15(fL, L, BLK);                                /* :346-347 */
        arm_float_to_q15(fR, R, BLK);
    }
    for (int n = 0; n < BLK; n++) {
        audio[2 * n] = L[n];
        audio[2 * n + 1] = R[n];
        if (f32) { f32[2 * n] = fL[n]; f32[2 * n + 1] = fR[n]; }
    }
    if (sm & RDSP_STAGE_SPEC1024) oracle_fft1024_update(&c->fft1024, L);  /* RadioDSP_SDR_RX.ino:87 */
}

void rdsp_oracle_chan_process(rdsp_oracle_chan_t *c, uint32_t n_blocks,
                              const int16_t *iq, size_t stride_in,
                              int16_t *audio, size_t stride_out,
                              float *f32, size_t stride_f32)
{
    for (uint32_t b = 0; b < n_blocks; b++)
        tick(c, iq + b * stride_in, audio + b * stride_out, f32 ? f32 + b * stride_f32 : 0);
}

void rdsp_oracle_bank_process(rdsp_oracle_chan_t **chans, uint32_t ch_first, uint32_t ch_count,
                              uint32_t n_total, uint32_t n_blocks, const int16_t *iq, int16_t *audio)
{
    const size_t stride = (size_t)n_total * 2 * BLK;
    for (uint32_t ch = ch_first; ch < ch_first + ch_count; ch++)
        rdsp_oracle_chan_process(chans[ch], n_blocks, iq + (size_t)ch * 2 * BLK, stride,
                                 audio + (size_t)ch * 2 * BLK, stride, 0, 0);
}

/* waterfall rows [50][128] (row 0 newest) and the colour class of every cell, thresholds of RDSP_display.h:299-318
 * with low = 0: 6 red >= 75, 5 magenta >= 50, 4 orange >= 40, 3 yellow >= 25, 2 blue >= 15, 1 navy >= 5, 0 black */
void rdsp_oracle_chan_read_waterfall(rdsp_oracle_chan_t *c, uint16_t *rows, uint8_t *colour)
{
    memcpy(rows, c->waterfall, sizeof(c->waterfall));
    if (!colour) return;
    for (int i = 0; i < 50 * 128; i++) {
        const int v = rows[i], low = 0;
        colour[i] = v >= low + 75 ? 6 : v >= low + 50 ? 5 : v >= low + 40 ? 4 : v >= low + 25 ? 3 : v >= low + 15 ? 2 : v >= low + 5 ? 1 : 0;
    }
}

int rdsp_oracle_chan_read_spectrum(rdsp_oracle_chan_t *c, uint16_t *out)
{
    int avail = c->outputflag;                                       /* available(), analyze_fft256iq.h:61-67 */
    c->outputflag = 0;
    memcpy(out, c->output, sizeof(c->output));
    return avail;
}

int rdsp_oracle_chan_read_audio_spectrum(rdsp_oracle_chan_t *c, uint16_t *out)
{
    int avail = c->fft1024.outputflag;
    c->fft1024.outputflag = 0;
    memcpy(out, c->fft1024.output, sizeof(c->fft1024.output));
    return avail;
}

/* K11: Update_Panadapter pre-processing, RDSP_display.h:260-280, and Update_smeter, :366-374 */
void rdsp_oracle_chan_read_panadapter(rdsp_oracle_chan_t *c, uint16_t *trace, float *smeter)
{
    int scale = 5;
    float avg = 0.0;
    float LPFcoeff = 0.7;
    for (int x = 0; x < 256; x++) {
        if ((x > 1) && (x < 254))
            avg = c->output[x] * 0.7 + c->output[x - 1] * 0.3 + c->output[x - 2] * 0.15
                + c->output[x + 1] * 0.3 + c->output[x + 2] * 0.15;
        else
            avg = c->output[x];
        c->spectrum_view[x] = (uint16_t)(LPFcoeff * 2 * sqrtf(ARD_ABS(avg) * scale) + (1 - LPFcoeff) * c->spectrum_view_old[x]);
        c->spectrum_view_old[x] = c->spectrum_view[x];
    }
    memcpy(trace, c->spectrum_view, sizeof(c->spectrum_view));
    /* waterfall, RDSP_display.h:282-297: the new line is SpectrumView[2x], x <= 127; every older line moves one row
     * down.  The reference's loop also copies row -1 into row 0 (out of bounds); restated without
     * that read: row 0 keeps the new line. */
    for (int row = 50 - 1; row >= 1; row--) memcpy(c->waterfall[row], c->waterfall[row - 1], sizeof(c->waterfall[0]));
    for (int x = 0; x <= 127; x++) c->waterfall[0][x] = c->spectrum_view[x * 2];
    float specVal = 0.0;
    for (int m = 75; m <= 85; m++) specVal = specVal + c->output[m];
    *smeter = ARD_ABS(specVal / 5);
}
