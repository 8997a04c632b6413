// fft_q15.cuh — radix-4 decimation-in-frequency q15 FFT, butterfly by butterfly.
//
// Replaces arm_cfft_radix4_q15 (call sites analyze_fft256iq.cpp:82 and AudioAnalyzeFFT1024::update);
// the per-stage fixed-point arithmetic is the ARMv7E-M code path of CMSIS-DSP as shipped in the
// reference's firmware image: stage 1 pre-scales by 1/4 and halves, middle
// stages quarter / halve, the last stage halves => net gain 1/N.  Every butterfly of a stage is
// independent of the others, so any assignment of butterflies to threads reproduces the sequential
// result bit for bit; stages are separated by a barrier.  Output is in bit-reversed order (the caller
// reads word bitrev(i) for bin i instead of permuting).
//
// Sample word = (re | im << 16); twiddle word k of tw = (cos | sin << 16)(2 pi k / 4096).
#pragma once
#include "rdsp_common.cuh"

#define QHD __host__ __device__ __forceinline__

namespace q15fft {

QHD int32_t s16(int32_t v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }
QHD int32_t lo(uint32_t a) { return (int32_t)(int16_t)(a & 0xFFFFu); }
QHD int32_t hi(uint32_t a) { return ((int32_t)a) >> 16; }
QHD uint32_t mk(int32_t l, int32_t h) { return ((uint32_t)l & 0xFFFFu) | ((uint32_t)h << 16); }
QHD uint32_t shadd(uint32_t a, uint32_t b) { return mk((lo(a) + lo(b)) >> 1, (hi(a) + hi(b)) >> 1); }
QHD uint32_t shsub(uint32_t a, uint32_t b) { return mk((lo(a) - lo(b)) >> 1, (hi(a) - hi(b)) >> 1); }
QHD uint32_t qadd(uint32_t a, uint32_t b) { return mk(s16(lo(a) + lo(b)), s16(hi(a) + hi(b))); }
QHD uint32_t qsub(uint32_t a, uint32_t b) { return mk(s16(lo(a) - lo(b)), s16(hi(a) - hi(b))); }
QHD uint32_t qasx(uint32_t a, uint32_t b) { return mk(s16(lo(a) - hi(b)), s16(hi(a) + lo(b))); }
QHD uint32_t qsax(uint32_t a, uint32_t b) { return mk(s16(lo(a) + hi(b)), s16(hi(a) - lo(b))); }
QHD uint32_t shasx(uint32_t a, uint32_t b) { return mk((lo(a) - hi(b)) >> 1, (hi(a) + lo(b)) >> 1); }
QHD uint32_t shsax(uint32_t a, uint32_t b) { return mk((lo(a) + hi(b)) >> 1, (hi(a) - lo(b)) >> 1); }
QHD uint32_t quarter(uint32_t a) { return mk(lo(a) >> 2, hi(a) >> 2); }      // SHADD16(SHADD16(a,0),0)
// twiddle * sample, both products keep their top 16 bits (SMUAD >> 16, SMUSDX & 0xFFFF0000)
QHD uint32_t cmul(uint32_t c, uint32_t r)
{
    const uint32_t re = (uint32_t)(lo(c) * lo(r)) + (uint32_t)(hi(c) * hi(r));
    const uint32_t im = (uint32_t)(lo(c) * hi(r)) - (uint32_t)(hi(c) * lo(r));
    return (im & 0xFFFF0000u) | (re >> 16);
}

// first stage, butterfly i in [0, N/4); mod = 4096 / N
QHD void first(uint32_t *src, const uint32_t *tw, int N, int mod, int i)
{
    const int n2 = N >> 2, ic = i * mod;
    uint32_t *p0 = src + i, *p1 = p0 + n2, *p2 = p1 + n2, *p3 = p2 + n2;
    const uint32_t xa = quarter(*p0), xb = quarter(*p1), xc = quarter(*p2), xd = quarter(*p3);
    uint32_t R = qadd(xa, xc), S = qsub(xa, xc);
    const uint32_t T2 = qadd(xb, xd);
    *p0 = shadd(R, T2);
    R = qsub(R, T2);
    *p1 = cmul(tw[2 * ic], R);
    const uint32_t T = qsub(xb, xd);
    R = qasx(S, T);
    S = qsax(S, T);
    *p2 = cmul(tw[ic], S);
    *p3 = cmul(tw[3 * ic], R);
}

// middle stage with group span n1 and quarter span n2 = n1/4, twiddle step mod; butterfly b in [0, N/4)
QHD void middle(uint32_t *src, const uint32_t *tw, int n1, int n2, int mod, int b)
{
    const int j = b % n2, grp = b / n2, ic = j * mod;
    uint32_t *p0 = src + j + grp * n1, *p1 = p0 + n2, *p2 = p1 + n2, *p3 = p2 + n2;
    const uint32_t xa = *p0, xb = *p1, xc = *p2, xd = *p3;
    uint32_t R = qadd(xa, xc), S = qsub(xa, xc);
    uint32_t T = qadd(xb, xd);
    *p0 = shadd(shadd(R, T), 0u);
    R = shsub(R, T);
    *p1 = cmul(tw[2 * ic], R);
    T = qsub(xb, xd);
    R = shasx(S, T);
    S = shsax(S, T);
    *p2 = cmul(tw[ic], S);
    *p3 = cmul(tw[3 * ic], R);
}

// last stage, butterfly b in [0, N/4) on words 4b..4b+3
QHD void last(uint32_t *src, int b)
{
    uint32_t *w = src + 4 * b;
    const uint32_t xa = w[0], xb = w[1], xc = w[2], xd = w[3];
    const uint32_t R = qadd(xa, xc), T = qadd(xb, xd), S = qsub(xa, xc), U = qsub(xb, xd);
    w[0] = shadd(R, T);
    w[1] = shsub(R, T);
    w[2] = shsax(S, U);
    w[3] = shasx(S, U);
}

QHD uint32_t bitrev(uint32_t i, int bits)
{
    uint32_t r = 0;
    for (int b = 0; b < bits; b++) r |= ((i >> b) & 1u) << (bits - 1 - b);
    return r;
}

}  // namespace q15fft
#undef QHD
