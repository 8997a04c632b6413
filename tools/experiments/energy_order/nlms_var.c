// experiment: sensitivity of the NLMS (arm_lms_norm_f32 semantics, 96 taps, desired = previous block) to the order in which the
// running energy is evaluated: mode 0 = reference order, mode 1 = anchored every 16 samples (fma-summed squares), slides between
#include <math.h>
#include <string.h>
#include <stdlib.h>
#define NT 96
void nlms_run(const float *x, int nblocks, float mu, int anchored, int emit_err, float *out)
{
    float coef[NT]; memset(coef, 0, sizeof coef);
    float *hist = calloc((size_t)nblocks * 128 + 256, sizeof(float));    // hist[128 + n] = x[n]
    memcpy(hist + 128, x, (size_t)nblocks * 128 * sizeof(float));
    float energy = 0.f;
    for (int b = 0; b < nblocks; b++) {
        const float *cur = hist + 128 + b * 128;
        const float *des = b == 0 ? cur : cur - 128;
        float anchor[9]; anchor[0] = energy;
        if (anchored) for (int k = 0; k < 8; k++) {
            float qi = 0.f, qo = 0.f;
            for (int j = 0; j < 16; j++) { float xn = cur[16 * k + j], xo = cur[16 * k + j - 96]; qi = fmaf(xn, xn, qi); qo = fmaf(xo, xo, qo); }
            anchor[k + 1] = anchor[k] + (qi - qo);
        }
        for (int n = 0; n < 128; n++) {
            if (anchored && (n & 15) == 0) energy = anchor[n >> 4];
            float in = cur[n], x0 = cur[n - 96];
            energy -= x0 * x0; energy += in * in;
            float sum = 0.f;
            for (int k = 0; k < NT; k++) sum += cur[n - 95 + k] * coef[k];      // pState[k] = x[n - 95 + k]
            float e = des[n] - sum;
            out[b * 128 + n] = emit_err ? e : sum;
            float den = energy + 0.000000119209289f; if (den < 0.000000119209289f) den = 0.000000119209289f;
            float w = (e * mu) / den;
            for (int k = 0; k < NT; k++) coef[k] += w * cur[n - 95 + k];
        }
        if (anchored) energy = anchor[8];
    }
    free(hist);
}
