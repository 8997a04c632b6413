// fft_f32_lanes.cuh — 256-point complex f32 FFT spread over the 32 lanes of a warp, 8 points per lane.
//
// Replaces arm_cfft_f32(&arm_cfft_sR_f32_len256, ...) at RDSP_convolutional.h:291,309: unnormalised
// forward DFT e^{-j 2 pi k n / N}, natural order in and out.
//
// Decomposition 256 = 8 x 8 x 4 (decimation in frequency), two shared-memory exchanges:
//   A : lane n1 holds x[n1 + 32 n2], n2 < 8.  Radix-8 over n2, twiddle W256^(n1 k2)   -> y[k2][n1]
//   B1: lane (k2, m1) takes y[k2][m1 + 4 m2], m2 < 8.  Radix-8 over m2, twiddle W32^(m1 q2) -> z[q2][k2][m1]
//   B2: pair p = lane + 32 h (k2 = p & 7, q2 = p >> 3) takes z[q2][k2][0..3], radix-4 over m1
//       -> X[64 q1 + p], stored in register j = 2 q1 + h, i.e. X[lane + 32 j].
// The output distribution equals the input distribution, so forward FFT, spectral product and inverse
// FFT chain in registers without an intermediate exchange.
//
// The phases are written as host+device functions of (lane, registers, exchange buffer) so that the
// exact arithmetic can be unit-tested on the CPU (tests/host/test_device_math.cu); the device wrapper
// below inserts the warp barriers.
#pragma once
#include <cuda_runtime.h>

#define FFT256_LDY 36              // padded row length of the exchange buffer (float2 units)
#define FFT256_BUF (8 * FFT256_LDY)

#define HD __host__ __device__ __forceinline__

// Complex arithmetic on the packed f32x2 instructions of sm_100 (FADD2 / FMUL2 / FFMA2: two IEEE f32 operations in one
// issue slot; half swaps, per-half negation and scalar broadcast are operand modifiers, so the swapped / negated pairs
// below cost nothing).  The host versions perform the same roundings (an explicit fmaf where the device fuses), which
// keeps tests/host/test_device_math.cu a test of the device arithmetic.
#ifndef RDSP_FFT_PACKED
#define RDSP_FFT_PACKED 1
#endif
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000 && RDSP_FFT_PACKED
HD float2 p_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
HD float2 p_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
HD float2 p_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#else
HD float2 p_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
HD float2 p_mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
HD float2 p_fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
HD float2 c_add(float2 a, float2 b) { return p_add(a, b); }
HD float2 c_sub(float2 a, float2 b) { return p_add(a, make_float2(-b.x, -b.y)); }
HD float2 c_mul_mj(float2 a) { return make_float2(a.y, -a.x); }                                               // a * (-j)
// a * b = (a.x b.x - a.y b.y, a.y b.x + a.x b.y)
HD float2 c_mul(float2 a, float2 b) { return p_fma(make_float2(a.y, a.x), make_float2(-b.y, b.y), p_mul(a, make_float2(b.x, b.x))); }
// conj(a * b) = (a.x b.x - a.y b.y, -a.y b.x - a.x b.y)
HD float2 c_mul_conj_result(float2 a, float2 b) { return p_fma(make_float2(a.y, a.x), make_float2(-b.y, -b.y), p_mul(a, make_float2(b.x, -b.x))); }
// a * conj(w) = (a.x w.x + a.y w.y, a.y w.x - a.x w.y)
HD float2 c_mulconj(float2 a, float2 w) { return p_fma(make_float2(a.y, -a.x), make_float2(w.y, w.y), p_mul(a, make_float2(w.x, w.x))); }

HD void dft4(float2 &c0, float2 &c1, float2 &c2, float2 &c3)
{
    const float2 s0 = c_add(c0, c2), s1 = c_sub(c0, c2), s2 = c_add(c1, c3), s3 = c_mul_mj(c_sub(c1, c3));
    c0 = c_add(s0, s2); c1 = c_add(s1, s3); c2 = c_sub(s0, s2); c3 = c_sub(s1, s3);
}

// forward 8-point DFT, natural order in and out
HD void dft8(float2 v[8])
{
    const float r = 0.70710678118654752440f;
    float2 a0 = c_add(v[0], v[4]), a1 = c_add(v[1], v[5]), a2 = c_add(v[2], v[6]), a3 = c_add(v[3], v[7]);
    float2 b0 = c_sub(v[0], v[4]);
    const float2 d1 = c_sub(v[1], v[5]), d3 = c_sub(v[3], v[7]);
    float2 b1 = p_mul(p_add(d1, make_float2(d1.y, -d1.x)), make_float2(r, r));                       // * W8^1
    float2 b2 = c_mul_mj(c_sub(v[2], v[6]));                                                         // * W8^2
    float2 b3 = p_mul(p_add(make_float2(d3.y, -d3.x), make_float2(-d3.x, -d3.y)), make_float2(r, r)); // * W8^3
    dft4(a0, a1, a2, a3);
    dft4(b0, b1, b2, b3);
    v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
    v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// tw[k] = (cos, sin)(2 pi k / 256)
HD void fft256_phaseA(int lane, float2 v[8], float2 *buf, const float2 *tw)
{
    dft8(v);
#pragma unroll
    for (int k2 = 1; k2 < 8; k2++) v[k2] = c_mulconj(v[k2], tw[(lane * k2) & 255]);
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) buf[k2 * FFT256_LDY + lane] = v[k2];
}
HD void fft256_phaseB1_load(int lane, float2 v[8], const float2 *buf)
{
    const int k2 = lane >> 2, m1 = lane & 3;
#pragma unroll
    for (int m2 = 0; m2 < 8; m2++) v[m2] = buf[k2 * FFT256_LDY + m1 + 4 * m2];
}
HD void fft256_phaseB1_store(int lane, float2 v[8], float2 *buf, const float2 *tw)
{
    const int m1 = lane & 3;
    dft8(v);
#pragma unroll
    for (int q2 = 1; q2 < 8; q2++) v[q2] = c_mulconj(v[q2], tw[8 * m1 * q2]);
#pragma unroll
    for (int q2 = 0; q2 < 8; q2++) buf[q2 * FFT256_LDY + lane + ((lane >> 4) << 1)] = v[q2];   // upper half of a row 16 bytes further: see phase B2      // lane = k2*4 + m1
}
HD void fft256_phaseB2(int lane, float2 v[8], const float2 *buf)
{
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int p = lane + 32 * h, k2 = p & 7, q2 = p >> 3;
        // 32-byte reads of eight lanes, 32 bytes apart: k2 and k2 + 4 would meet on the same banks (the ncu capture of r02b
        // showed a third of the kernel's shared-memory wavefronts were conflicts), so rows are stored with their upper half
        // skewed by 16 bytes
        const float2 *z = buf + q2 * FFT256_LDY + k2 * 4 + ((k2 >> 2) << 1);
        float2 z0 = z[0], z1 = z[1], z2 = z[2], z3 = z[3];
        dft4(z0, z1, z2, z3);
        v[0 + h] = z0; v[2 + h] = z1; v[4 + h] = z2; v[6 + h] = z3;
    }
}

// The same transform with the lane's 14 twiddles held in registers (twA[k] = W256^(lane (k+1)), twB[q] = W32^(m1 (q+1))):
// they are loop invariants of a kernel that transforms block after block, and the table reads were 30 % of the shared-memory
// wavefronts of a transform.
HD void fft256_load_twiddles(int lane, const float2 *tw, float2 twA[7], float2 twB[7])
{
    const int m1 = lane & 3;
#pragma unroll
    for (int k = 1; k < 8; k++) { twA[k - 1] = tw[(lane * k) & 255]; twB[k - 1] = tw[8 * m1 * k]; }
}
HD void fft256_phaseA_r(int lane, float2 v[8], float2 *buf, const float2 twA[7])
{
    dft8(v);
#pragma unroll
    for (int k2 = 1; k2 < 8; k2++) v[k2] = c_mulconj(v[k2], twA[k2 - 1]);
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) buf[k2 * FFT256_LDY + lane] = v[k2];
}
HD void fft256_phaseB1_store_r(int lane, float2 v[8], float2 *buf, const float2 twB[7])
{
    dft8(v);
#pragma unroll
    for (int q2 = 1; q2 < 8; q2++) v[q2] = c_mulconj(v[q2], twB[q2 - 1]);
#pragma unroll
    for (int q2 = 0; q2 < 8; q2++) buf[q2 * FFT256_LDY + lane + ((lane >> 4) << 1)] = v[q2];   // upper half of a row 16 bytes further: see phase B2
}

#ifdef __CUDACC__
__device__ __forceinline__ void fft256_warp_r(int lane, float2 v[8], float2 *buf, const float2 twA[7], const float2 twB[7])
{
    fft256_phaseA_r(lane, v, buf, twA);
    __syncwarp();
    fft256_phaseB1_load(lane, v, buf);
    __syncwarp();
    fft256_phaseB1_store_r(lane, v, buf, twB);
    __syncwarp();
    fft256_phaseB2(lane, v, buf);
    __syncwarp();
}
// v[j] = x[lane + 32 j] on entry, X[lane + 32 j] on exit.  buf: FFT256_BUF float2 of warp-private smem.
__device__ __forceinline__ void fft256_warp(int lane, float2 v[8], float2 *buf, const float2 *tw)
{
    fft256_phaseA(lane, v, buf, tw);
    __syncwarp();
    fft256_phaseB1_load(lane, v, buf);
    __syncwarp();
    fft256_phaseB1_store(lane, v, buf, tw);
    __syncwarp();
    fft256_phaseB2(lane, v, buf);
    __syncwarp();
}
#endif
#undef HD
