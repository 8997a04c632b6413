// test_device_math.cu — CPU execution of the host+device arithmetic used by the kernels
// (fft_f32_lanes.cuh, fft_q15.cuh, host_design.cpp), checked against the oracle.  Built with nvcc as a
// plain host program; needs no GPU.  Run by tests/test_host_math.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../radiodsp_sdr_rx_b200/csrc/fft_f32_lanes.cuh"
#include "../../radiodsp_sdr_rx_b200/csrc/fft_q15.cuh"
#include "../../radiodsp_sdr_rx_b200/csrc/host_design.h"
#include "../../oracle/cmsis_shim.h"
#include "../../oracle/teensy_shim.h"
#include "../../oracle/rdsp_oracle.h"

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); g_fail++; } } while (0)

static uint32_t rng_state = 12345;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static void test_fft256_lanes()
{
    float cs[512];
    rdsp_host::make_twiddle_256_f32(cs);
    const float2 *tw = reinterpret_cast<const float2 *>(cs);
    std::vector<double> xr(256), xi(256);
    for (int i = 0; i < 256; i++) { xr[i] = (double)((int)(rnd() % 65536) - 32768) / 32768.0; xi[i] = (double)((int)(rnd() % 65536) - 32768) / 32768.0; }
    float2 v[32][8], buf[FFT256_BUF];
    for (int l = 0; l < 32; l++) for (int j = 0; j < 8; j++) v[l][j] = make_float2((float)xr[l + 32 * j], (float)xi[l + 32 * j]);
    for (int l = 0; l < 32; l++) fft256_phaseA(l, v[l], buf, tw);
    for (int l = 0; l < 32; l++) fft256_phaseB1_load(l, v[l], buf);
    for (int l = 0; l < 32; l++) fft256_phaseB1_store(l, v[l], buf, tw);
    for (int l = 0; l < 32; l++) fft256_phaseB2(l, v[l], buf);
    double maxerr = 0, maxmag = 0;
    for (int k = 0; k < 256; k++) {
        double re = 0, im = 0;
        for (int n = 0; n < 256; n++) {
            const double a = -2.0 * M_PI * (double)((k * n) & 255) / 256.0;
            re += xr[n] * cos(a) - xi[n] * sin(a);
            im += xr[n] * sin(a) + xi[n] * cos(a);
        }
        const float2 g = v[k & 31][k >> 5];
        maxerr = fmax(maxerr, fmax(fabs(g.x - re), fabs(g.y - im)));
        maxmag = fmax(maxmag, hypot(re, im));
    }
    printf("fft256 lanes: max abs err %.3g (max |X| %.3g)\n", maxerr, maxmag);
    CHECK(maxerr < 2e-5 * maxmag, "fft256 lanes error too large");
    // oracle arm_cfft_f32 agrees too
    std::vector<float> p(512);
    for (int i = 0; i < 256; i++) { p[2 * i] = (float)xr[i]; p[2 * i + 1] = (float)xi[i]; }
    arm_cfft_f32(&arm_cfft_sR_f32_len256, p.data(), 0, 1);
    double d = 0;
    for (int k = 0; k < 256; k++) { const float2 g = v[k & 31][k >> 5]; d = fmax(d, fmax(fabs(g.x - p[2 * k]), fabs(g.y - p[2 * k + 1]))); }
    CHECK(d < 2e-5 * maxmag, "fft256 lanes vs oracle arm_cfft_f32: %g", d);
}

static void q15_fft_product(int2 *buf, const int2 *tw, int N)
{
    const int nb = N / 4;
    for (int b = 0; b < nb; b++) q15fft::first(buf, tw, N, 4096 / N, b);
    int mod = (4096 / N) * 4, n1 = N / 4;
    for (int k = N / 4; k > 4; k >>= 2) {
        const int n2 = n1 / 4;
        for (int b = 0; b < nb; b++) q15fft::middle(buf, tw, n1, n2, mod, b);
        n1 = n2; mod *= 4;
    }
    for (int b = 0; b < nb; b++) q15fft::last(buf, b);
}

static void test_q15_fft(int N, int mode)
{
    std::vector<uint32_t> tww(3072);
    rdsp_host::make_twiddle_4096_q15(tww.data());
    CHECK(memcmp(tww.data(), oracle_twiddle_4096_q15(), 3072 * 4) == 0, "q15 twiddle table differs from oracle");
    std::vector<int2> tw(3072);
    for (int k = 0; k < 3072; k++) tw[k] = make_int2((int16_t)(tww[k] & 0xFFFFu), (int16_t)(tww[k] >> 16));
    std::vector<uint32_t> a(N);
    std::vector<int2> b(q15fft::padded(N));
    for (int i = 0; i < N; i++) {
        int re, im;
        if (mode == 0) { re = (int)(rnd() % 65536) - 32768; im = (int)(rnd() % 65536) - 32768; }
        else if (mode == 1) { re = (rnd() & 1) ? 32767 : -32768; im = (rnd() & 1) ? 32767 : -32768; }
        else if (mode == 3) { re = (i & 1) ? 32767 : -32768; im = (i & 2) ? -32768 : 32767; }
        else { re = (int)lrint(20000 * cos(2 * M_PI * 7 * i / N)); im = (int)lrint(20000 * sin(2 * M_PI * 7 * i / N)); }
        a[i] = ((uint32_t)(uint16_t)(int16_t)re) | ((uint32_t)(uint16_t)(int16_t)im << 16);
        b[q15fft::P(i)] = make_int2(re, im);
    }
    arm_cfft_radix4_instance_q15 inst;
    arm_cfft_radix4_init_q15(&inst, (uint16_t)N, 0, 1);
    arm_cfft_radix4_q15(&inst, reinterpret_cast<q15_t *>(a.data()));
    q15_fft_product(b.data(), tw.data(), N);
    int bits = 0; while ((1 << bits) < N) bits++;
    int bad = 0;
    for (int i = 0; i < N; i++) {
        const int2 g = b[q15fft::P((int)q15fft::bitrev((uint32_t)i, bits))];
        const int re = (int16_t)(a[i] & 0xFFFF), im = (int16_t)(a[i] >> 16);
        if (g.x != re || g.y != im) bad++;
    }
    CHECK(bad == 0, "q15 FFT N=%d mode=%d: %d bins differ", N, mode, bad);
    if (mode == 2) {   // sanity: tone at bin 7 with gain 1/N
        const int re = (int16_t)(a[7] & 0xFFFF);
        CHECK(abs(re - 20000) < 64, "q15 FFT tone bin value %d", re);
    }
}

static void test_div_magic()
{
    for (uint32_t d = 1; d <= 255; d++) {
        int lg = 0; while ((1u << lg) < d) lg++;
        const int sh = 32 + lg;
        const unsigned long long M = ((1ull << sh) + d - 1) / d;
        const uint32_t edge[] = {0u, 1u, d - 1, d, d + 1, 0x7FFFFFFFu, 0x80000000u, 0x7FFFFFFEu, 2u * d, 1000u * d - 1};
        for (uint32_t x : edge) { if (x > 0x80000000u) continue; CHECK((uint32_t)(((unsigned long long)x * M) >> sh) == x / d, "div magic d=%u x=%u", d, x); }
        for (int i = 0; i < 20000; i++) {
            uint32_t x = (rnd() << 8) ^ rnd(); x &= 0x7FFFFFFFu;
            if ((uint32_t)(((unsigned long long)x * M) >> sh) != x / d) { CHECK(false, "div magic d=%u x=%u", d, x); break; }
        }
        for (uint32_t q = 1; q < 200; q++) {        // multiples of d around the top of the range
            const uint32_t x = (0x80000000u / d - q) * d;
            CHECK((uint32_t)(((unsigned long long)x * M) >> sh) == x / d && (uint32_t)(((unsigned long long)(x - 1) * M) >> sh) == (x - 1) / d, "div magic top d=%u", d);
        }
    }
}

static void test_design_tables()
{
    int16_t w[1024];
    rdsp_host::make_hann_q15(w, 256);
    CHECK(memcmp(w, oracle_hanning256(), 512) == 0, "hann256 differs");
    rdsp_host::make_hann_q15(w, 1024);
    CHECK(memcmp(w, oracle_hanning1024(), 2048) == 0, "hann1024 differs");
    float cs[512];
    rdsp_host::make_twiddle_256_f32(cs);
    CHECK(memcmp(cs, oracle_twiddle_256_f32(), sizeof(cs)) == 0, "f32 twiddles differ");
    int16_t a[129], b[129], c[129];
    for (int m = 0; m < RDSP_DEMOD_COUNT; m++) {
        rdsp_host::design_hilbert_pair(m, a, b);
        rdsp_oracle_get_taps(RDSP_TAPS_HILBERT_I, m, c);
        CHECK(memcmp(a, c, sizeof(a)) == 0, "hilbert I taps differ, mode %d", m);
        rdsp_oracle_get_taps(RDSP_TAPS_HILBERT_Q, m, c);
        CHECK(memcmp(b, c, sizeof(b)) == 0, "hilbert Q taps differ, mode %d", m);
    }
    for (int f = 0; f < RDSP_FILTER_COUNT; f++) {
        rdsp_host::design_bandpass(f, a);
        rdsp_oracle_get_taps(RDSP_TAPS_BANDPASS, f, c);
        CHECK(memcmp(a, c, sizeof(a)) == 0, "band-pass taps differ, filter %d", f);
    }
    float m1[512], m2[512];
    rdsp_host::design_mask(300.0, 4000.0, m1);
    rdsp_oracle_design_mask(300.0, 4000.0, m2);
    double md = 0;
    for (int i = 0; i < 512; i++) md = fmax(md, fabs((double)m1[i] - m2[i]));
    printf("mask: max |product - oracle| = %.3g\n", md);
    CHECK(md < 2e-6, "mask differs by %g", md);
    for (int s = 0; s <= 60; s += 5) CHECK(rdsp_host::lms_mu(s) == rdsp_oracle_lms_mu(s), "mu differs at %d", s);
    int32_t coef[5];
    oracle_biquad_t bq;
    rdsp_host::biquad_highpass_q30(500.0f, 0.5f, coef);
    oracle_biquad_set_highpass(&bq, 500.0f, 0.5f);
    CHECK(memcmp(coef, bq.def, sizeof(coef)) == 0, "biquad coefficients differ");
    printf("biquad Q30: %d %d %d %d %d\n", coef[0], coef[1], coef[2], coef[3], coef[4]);
}

int main()
{
    test_fft256_lanes();
    for (int mode = 0; mode < 4; mode++) { test_q15_fft(256, mode); test_q15_fft(1024, mode); }
    for (int r = 0; r < 200; r++) { test_q15_fft(256, r & 1); test_q15_fft(1024, r & 1); }
    test_div_magic();
    test_design_tables();
    printf(g_fail ? "FAILED (%d)\n" : "ALL OK\n", g_fail);
    return g_fail ? 1 : 0;
}
