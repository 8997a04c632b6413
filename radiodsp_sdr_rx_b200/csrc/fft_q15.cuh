// fft_q15.cuh — radix-4 decimation-in-frequency q15 FFT, butterfly by butterfly.
//
// Replaces arm_cfft_radix4_q15 (call sites analyze_fft256iq.cpp:82 and AudioAnalyzeFFT1024::update);
// the per-stage fixed-point arithmetic is the ARMv7E-M code path of CMSIS-DSP as shipped in the
// reference's firmware image: stage 1 pre-scales by 1/4 and halves, middle
// stages quarter / halve, the last stage halves => net gain 1/N.  Every butterfly of a stage is
// independent of the others, so any assignment of butterflies to threads reproduces the sequential
// result bit for bit; stages are separated by a barrier.  Output is in bit-reversed order (the caller
// reads element bitrev(i) for bin i instead of permuting).
//
// The ARM code works on packed (re | im << 16) words with SIMD instructions; on B200 shifts, permutes and
// min/max issue at half the IMAD rate (tools/ubench_pipes.cu), so here a sample is an UNPACKED int2 (re, im)
// in shared memory and every lane operation is written on 32-bit integers: no pack / unpack per operation,
// the same results.  Two facts keep it exact and short:
//   * stage 1 cannot saturate: its inputs are pre-scaled to [-8192, 8191], so every sum / difference it forms
//     stays inside int16 and QADD16 / QSUB16 / QASX / QSAX reduce to plain adds;
//   * (SMUAD >> 16, SMUSDX & 0xFFFF0000) = arithmetic >> 16 of the two 32-bit wrap-around sums.
// Element i is stored at P(i) = i + 4 * (i >> 4) (one 32-byte skew per 128 bytes) to spread the strided
// accesses of the late stages over the banks.  Twiddle k = (cos, sin)(2 pi k / 4096) as int2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define QHD __host__ __device__ __forceinline__

namespace q15fft {

QHD int P(int i) { return i + ((i >> 4) << 2); }
QHD int padded(int n) { return n + (n >> 2); }

QHD int s16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }
// SAT = false: the caller has PROVED that no sum of this frame can leave int16 (see no_saturation_bound below), so
// QADD16 / QSUB16 are plain adds.  The saturating min / max pairs are a quarter of a butterfly's ALU-pipe work.
template <bool SAT> QHD int s16c(int v) { return SAT ? s16(v) : v; }
// QADD16 / QSUB16 of one half-word pair on 32-bit lanes.  On the device the clamp is written in PTX: left to the compiler,
// min(max(a + b)) of two values it has proved to be int16 becomes a 16-bit saturating add that it then expands into
// ~8 compare / select / permute instructions; ptxas turns this form into VIADDMNMX + VIMNMX.
template <bool SAT> QHD int qadd(int a, int b)
{
#ifdef __CUDA_ARCH__
    if (SAT) { int r; asm("{.reg .s32 t;\n\tadd.s32 t, %1, %2;\n\tmax.s32 t, t, -32768;\n\tmin.s32 %0, t, 32767;}" : "=r"(r) : "r"(a), "r"(b)); return r; }
#endif
    return s16c<SAT>(a + b);
}
template <bool SAT> QHD int qsub(int a, int b)
{
#ifdef __CUDA_ARCH__
    if (SAT) { int r; asm("{.reg .s32 t;\n\tsub.s32 t, %1, %2;\n\tmax.s32 t, t, -32768;\n\tmin.s32 %0, t, 32767;}" : "=r"(r) : "r"(a), "r"(b)); return r; }
#endif
    return s16c<SAT>(a - b);
}

// A frame whose windowed input components are all <= this bound in magnitude cannot saturate in any stage of the
// 256- or 1024-point transform.  With input bound m0: stage 1 works on x >> 2 (<= a0 = m0/4 + 1), forms sums <= 4 a0 and
// twiddle products <= 0.70711 * 4 a0 + 1, so its outputs are <= A1 = 0.7075 m0 + 4; a middle stage with input bound A
// saturates only if 2 A > 32767 and leaves A' = 1.4143 A + 2 (its largest outputs are the twiddle products of
// (R - T) >> 1 <= 2 A + 1 per component); the last stage saturates only if 2 A > 32767.  1024 points: A4 <= 16383 needs
// A1 <= 5788; 256 points: A3 <= 16383 needs A1 <= 8188.  11000 (1024-point REAL input, A1 = m0/2 + 3) and
// 11000 (256-point complex input) leave margin for the rounding terms.  (k_spec1024 uses it: 97 -> 92 us; in k_spec256 the
// second code path cost more registers than the min / max pairs it saved: 81 -> 85 us, not used there.)
constexpr int NO_SAT_BOUND_1024_REAL = 11000;
constexpr int NO_SAT_BOUND_256 = 11000;
// twiddle * sample: top 16 bits of the two wrap-around 32-bit sums
QHD int2 cmul(int2 c, int2 r)
{
    const uint32_t re = (uint32_t)(c.x * r.x) + (uint32_t)(c.y * r.y);   // SMUAD
    const uint32_t im = (uint32_t)(c.x * r.y) - (uint32_t)(c.y * r.x);   // SMUSDX
    return make_int2(((int32_t)re) >> 16, ((int32_t)im) >> 16);
}

// first stage, butterfly i in [0, N/4); mod = 4096 / N.  No saturation can occur (see header).
QHD void first(int2 *src, const int2 *tw, int N, int mod, int i)
{
    const int n2 = N >> 2, ic = i * mod;
    int2 *p0 = src + P(i), *p1 = src + P(i + n2), *p2 = src + P(i + 2 * n2), *p3 = src + P(i + 3 * n2);
    int2 xa = *p0, xb = *p1, xc = *p2, xd = *p3;
    xa.x >>= 2; xa.y >>= 2; xb.x >>= 2; xb.y >>= 2; xc.x >>= 2; xc.y >>= 2; xd.x >>= 2; xd.y >>= 2;
    const int Rx = xa.x + xc.x, Ry = xa.y + xc.y, Sx = xa.x - xc.x, Sy = xa.y - xc.y;
    const int Tx = xb.x + xd.x, Ty = xb.y + xd.y, Ux = xb.x - xd.x, Uy = xb.y - xd.y;
    *p0 = make_int2((Rx + Tx) >> 1, (Ry + Ty) >> 1);
    *p1 = cmul(tw[2 * ic], make_int2(Rx - Tx, Ry - Ty));
    *p2 = cmul(tw[ic], make_int2(Sx + Uy, Sy - Ux));          // QSAX(S, T): xa - xc - j (xb - xd)
    *p3 = cmul(tw[3 * ic], make_int2(Sx - Uy, Sy + Ux));      // QASX(S, T): xa - xc + j (xb - xd)
}

// first stage of a REAL input frame (imaginary parts are zero: AudioAnalyzeFFT1024 feeds (sample, 0) pairs): the same
// arithmetic with the zero terms dropped
QHD void first_real(int2 *src, const int2 *tw, int N, int mod, int i)
{
    const int n2 = N >> 2, ic = i * mod;
    int2 *p0 = src + P(i), *p1 = src + P(i + n2), *p2 = src + P(i + 2 * n2), *p3 = src + P(i + 3 * n2);
    const int xa = p0->x >> 2, xb = p1->x >> 2, xc = p2->x >> 2, xd = p3->x >> 2;
    const int Rx = xa + xc, Sx = xa - xc, Tx = xb + xd, Ux = xb - xd;
    *p0 = make_int2((Rx + Tx) >> 1, 0);
    *p1 = cmul(tw[2 * ic], make_int2(Rx - Tx, 0));
    *p2 = cmul(tw[ic], make_int2(Sx, -Ux));
    *p3 = cmul(tw[3 * ic], make_int2(Sx, Ux));
}

// middle stage with group span n1 and quarter span n2 = n1/4, twiddle step mod; butterfly b in [0, N/4)
template <bool SAT = true>
QHD void middle(int2 *src, const int2 *tw, int n1, int n2, int mod, int b)
{
    const int j = b % n2, grp = b / n2, ic = j * mod, i0 = j + grp * n1;
    int2 *p0 = src + P(i0), *p1 = src + P(i0 + n2), *p2 = src + P(i0 + 2 * n2), *p3 = src + P(i0 + 3 * n2);
    const int2 xa = *p0, xb = *p1, xc = *p2, xd = *p3;
    const int Rx = qadd<SAT>(xa.x, xc.x), Ry = qadd<SAT>(xa.y, xc.y), Sx = qsub<SAT>(xa.x, xc.x), Sy = qsub<SAT>(xa.y, xc.y);
    const int Tx = qadd<SAT>(xb.x, xd.x), Ty = qadd<SAT>(xb.y, xd.y), Ux = qsub<SAT>(xb.x, xd.x), Uy = qsub<SAT>(xb.y, xd.y);
    *p0 = make_int2(((Rx + Tx) >> 1) >> 1, ((Ry + Ty) >> 1) >> 1);
    *p1 = cmul(tw[2 * ic], make_int2((Rx - Tx) >> 1, (Ry - Ty) >> 1));
    *p2 = cmul(tw[ic], make_int2((Sx + Uy) >> 1, (Sy - Ux) >> 1));        // SHSAX(S, T)
    *p3 = cmul(tw[3 * ic], make_int2((Sx - Uy) >> 1, (Sy + Ux) >> 1));    // SHASX(S, T)
}

// last stage, butterfly b in [0, N/4) on elements 4b..4b+3
template <bool SAT = true>
QHD void last(int2 *src, int b)
{
    int2 *w = src + P(4 * b);                                  // 4b..4b+3 never straddle a skew boundary
    const int2 xa = w[0], xb = w[1], xc = w[2], xd = w[3];
    const int Rx = qadd<SAT>(xa.x, xc.x), Ry = qadd<SAT>(xa.y, xc.y), Sx = qsub<SAT>(xa.x, xc.x), Sy = qsub<SAT>(xa.y, xc.y);
    const int Tx = qadd<SAT>(xb.x, xd.x), Ty = qadd<SAT>(xb.y, xd.y), Ux = qsub<SAT>(xb.x, xd.x), Uy = qsub<SAT>(xb.y, xd.y);
    w[0] = make_int2((Rx + Tx) >> 1, (Ry + Ty) >> 1);
    w[1] = make_int2((Rx - Tx) >> 1, (Ry - Ty) >> 1);
    w[2] = make_int2((Sx + Uy) >> 1, (Sy - Ux) >> 1);          // SHSAX(S, U)
    w[3] = make_int2((Sx - Uy) >> 1, (Sy + Ux) >> 1);          // SHASX(S, U)
}

// ---- the same butterflies on registers (in place: x0..x3 = elements i, i+n2, i+2 n2, i+3 n2) ----------------------
// A thread that holds 16 elements can run two consecutive stages between two shared-memory round trips.
QHD void first_r(int2 &x0, int2 &x1, int2 &x2, int2 &x3, int2 t1, int2 t2, int2 t3)          // t1 = tw[ic], t2 = tw[2ic], t3 = tw[3ic]
{
    int2 xa = x0, xb = x1, xc = x2, xd = x3;
    xa.x >>= 2; xa.y >>= 2; xb.x >>= 2; xb.y >>= 2; xc.x >>= 2; xc.y >>= 2; xd.x >>= 2; xd.y >>= 2;
    const int Rx = xa.x + xc.x, Ry = xa.y + xc.y, Sx = xa.x - xc.x, Sy = xa.y - xc.y;
    const int Tx = xb.x + xd.x, Ty = xb.y + xd.y, Ux = xb.x - xd.x, Uy = xb.y - xd.y;
    x0 = make_int2((Rx + Tx) >> 1, (Ry + Ty) >> 1);
    x1 = cmul(t2, make_int2(Rx - Tx, Ry - Ty));
    x2 = cmul(t1, make_int2(Sx + Uy, Sy - Ux));
    x3 = cmul(t3, make_int2(Sx - Uy, Sy + Ux));
}
QHD void first_real_r(int2 &x0, int2 &x1, int2 &x2, int2 &x3, int2 t1, int2 t2, int2 t3)     // t1 = tw[ic], t2 = tw[2ic], t3 = tw[3ic]
{
    const int xa = x0.x >> 2, xb = x1.x >> 2, xc = x2.x >> 2, xd = x3.x >> 2;
    const int Rx = xa + xc, Sx = xa - xc, Tx = xb + xd, Ux = xb - xd;
    x0 = make_int2((Rx + Tx) >> 1, 0);
    x1 = cmul(t2, make_int2(Rx - Tx, 0));
    x2 = cmul(t1, make_int2(Sx, -Ux));
    x3 = cmul(t3, make_int2(Sx, Ux));
}
template <bool SAT = true>
QHD void middle_r(int2 &x0, int2 &x1, int2 &x2, int2 &x3, int2 t1, int2 t2, int2 t3)
{
    const int2 xa = x0, xb = x1, xc = x2, xd = x3;
    const int Rx = qadd<SAT>(xa.x, xc.x), Ry = qadd<SAT>(xa.y, xc.y), Sx = qsub<SAT>(xa.x, xc.x), Sy = qsub<SAT>(xa.y, xc.y);
    const int Tx = qadd<SAT>(xb.x, xd.x), Ty = qadd<SAT>(xb.y, xd.y), Ux = qsub<SAT>(xb.x, xd.x), Uy = qsub<SAT>(xb.y, xd.y);
    x0 = make_int2(((Rx + Tx) >> 1) >> 1, ((Ry + Ty) >> 1) >> 1);
    x1 = cmul(t2, make_int2((Rx - Tx) >> 1, (Ry - Ty) >> 1));
    x2 = cmul(t1, make_int2((Sx + Uy) >> 1, (Sy - Ux) >> 1));
    x3 = cmul(t3, make_int2((Sx - Uy) >> 1, (Sy + Ux) >> 1));
}
template <bool SAT = true>
QHD void last_r(int2 &x0, int2 &x1, int2 &x2, int2 &x3)
{
    const int2 xa = x0, xb = x1, xc = x2, xd = x3;
    const int Rx = qadd<SAT>(xa.x, xc.x), Ry = qadd<SAT>(xa.y, xc.y), Sx = qsub<SAT>(xa.x, xc.x), Sy = qsub<SAT>(xa.y, xc.y);
    const int Tx = qadd<SAT>(xb.x, xd.x), Ty = qadd<SAT>(xb.y, xd.y), Ux = qsub<SAT>(xb.x, xd.x), Uy = qsub<SAT>(xb.y, xd.y);
    x0 = make_int2((Rx + Tx) >> 1, (Ry + Ty) >> 1);
    x1 = make_int2((Rx - Tx) >> 1, (Ry - Ty) >> 1);
    x2 = make_int2((Sx + Uy) >> 1, (Sy - Ux) >> 1);
    x3 = make_int2((Sx - Uy) >> 1, (Sy + Ux) >> 1);
}

QHD uint32_t bitrev(uint32_t i, int bits)
{
    uint32_t r = 0;
    for (int b = 0; b < bits; b++) r |= ((i >> b) & 1u) << (bits - 1 - b);
    return r;
}

}  // namespace q15fft
#undef QHD
