// k_front.cu — K0+K1+K2: input gain / IQ balance, Hilbert FIR pair, sideband sum or AM envelope,
// audio band-pass FIR.  All q15, bit-exact against oracle/rdsp_oracle.c:stage_frontend.
//
// Replaces the AudioSDRpreProcessor -> AudioSDR front half of the graph wired at
// RadioDSP_SDR_RX.ino:71-72,81-82 (library absent from the reference tree; arithmetic conventions
// are arm_fir_fast_q15 / AudioMixer4 as stated in SURVEY.md A.4).
//
// Mapping: one warp per channel, 8 channels per CTA.  A 128-sample block is 4 samples per lane; the
// 512-byte IQ block is one 128-bit load per lane.  Each q15 delay line lives in shared memory as
// [128 history | 128 current] int16; a lane produces 4 consecutive outputs from a register sliding
// window (8 taps x 4 outputs = 32 IMAD per 3 LDS.64 + 2 broadcast LDS.128 of taps).  The 32-bit
// accumulator wraps (unsigned arithmetic) exactly like the CMSIS fast FIR.
#include "rdsp_common.cuh"
#include "kernels.h"

namespace {

constexpr int WARPS = 8;

__device__ __forceinline__ int32_t mix_gain(int32_t x, int32_t mult)
{
    if (mult == 65536) return x;
    long long v = ((long long)mult * (long long)x) >> 16;
    v = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
    return (int32_t)v;
}

// buf: [0,128) history, [128,256) current block.  Outputs n = 4*lane + j, j < 4.
__device__ __forceinline__ void fir129_x4(const int16_t *buf, const int32_t *taps, int lane, int32_t y[4])
{
    const uint2 *b2 = reinterpret_cast<const uint2 *>(buf);
    uint32_t acc[4] = {0u, 0u, 0u, 0u};
#pragma unroll 2
    for (int c = 0; c < 16; c++) {
        const int q = 30 + lane - 2 * c;               // (128 + 4*lane - 8c - 8) / 4
        const uint2 w0 = b2[q], w1 = b2[q + 1], w2 = b2[q + 2];
        int32_t s[12];
        s[0] = lo16(w0.x); s[1] = hi16(w0.x); s[2]  = lo16(w0.y); s[3]  = hi16(w0.y);
        s[4] = lo16(w1.x); s[5] = hi16(w1.x); s[6]  = lo16(w1.y); s[7]  = hi16(w1.y);
        s[8] = lo16(w2.x); s[9] = hi16(w2.x); s[10] = lo16(w2.y); s[11] = hi16(w2.y);
        const int4 t0 = *reinterpret_cast<const int4 *>(taps + 8 * c);
        const int4 t1 = *reinterpret_cast<const int4 *>(taps + 8 * c + 4);
        const int32_t tp[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int kk = 0; kk < 8; kk++)
                acc[j] += (uint32_t)(tp[kk] * s[8 + j - kk]);
    }
    {   // tap 128 multiplies x[n-128] = history sample 4*lane + j
        const uint2 w = b2[lane];
        const int32_t t = taps[128];
        acc[0] += (uint32_t)(t * lo16(w.x));
        acc[1] += (uint32_t)(t * hi16(w.x));
        acc[2] += (uint32_t)(t * lo16(w.y));
        acc[3] += (uint32_t)(t * hi16(w.y));
    }
#pragma unroll
    for (int j = 0; j < 4; j++) y[j] = sat16(((int32_t)acc[j]) >> 15);
}

__global__ void __launch_bounds__(WARPS * 32) k_front(FrontArgs a)
{
    __shared__ __align__(16) int32_t s_taps[15 * RDSP_TAPS_PAD];
    __shared__ __align__(16) int16_t s_buf[WARPS][3][256];

    for (int i = threadIdx.x; i < 15 * RDSP_TAPS_PAD; i += WARPS * 32) s_taps[i] = a.taps[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ch = blockIdx.x * WARPS + warp;
    if (ch >= a.C) return;

    const RdspChanParams p = a.par[ch];
    const int32_t *tapA = s_taps + (0 + p.demod) * RDSP_TAPS_PAD;
    const int32_t *tapB = s_taps + (RDSP_N_DEMOD + p.demod) * RDSP_TAPS_PAD;
    const int32_t *tapM = s_taps + (2 * RDSP_N_DEMOD + p.filter) * RDSP_TAPS_PAD;
    int16_t *bI = s_buf[warp][0], *bQ = s_buf[warp][1], *bD = s_buf[warp][2];
    uint2 *bI2 = reinterpret_cast<uint2 *>(bI), *bQ2 = reinterpret_cast<uint2 *>(bQ), *bD2 = reinterpret_cast<uint2 *>(bD);

    // delay lines: hist[ch][3][128] int16
    uint2 *hrow = reinterpret_cast<uint2 *>(a.hist + (size_t)ch * 3 * RDSP_BLK);
    bI2[lane] = hrow[lane];
    bQ2[lane] = hrow[32 + lane];
    bD2[lane] = hrow[64 + lane];

    const bool usb = (p.demod == 1 || p.demod == 3);
    const bool am = (p.demod == 4);

    for (int t = 0; t < a.T; t++) {
        const size_t cb = (size_t)t * a.C + ch;                    // channel-block index
        const int4 v = ld_stream16(a.iq + cb * 2 * RDSP_BLK + lane * 8);
        const uint32_t w[4] = {(uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w};
        int32_t xi[4], xq[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            xi[j] = mix_gain(lo16(w[j]), p.mult_i);
            xq[j] = mix_gain(hi16(w[j]), p.mult_q);
        }
        bI2[32 + lane] = make_uint2(mk16(xi[0], xi[1]), mk16(xi[2], xi[3]));
        bQ2[32 + lane] = make_uint2(mk16(xq[0], xq[1]), mk16(xq[2], xq[3]));
        __syncwarp();

        int32_t ya[4], yb[4], d[4], m[4];
        fir129_x4(bI, tapA, lane, ya);
        fir129_x4(bQ, tapB, lane, yb);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (am) {
                uint32_t e = sqrt_u32_approx((uint32_t)(ya[j] * ya[j]) + (uint32_t)(yb[j] * yb[j]));
                d[j] = (int32_t)min(e, 32767u);
            } else {
                d[j] = usb ? sat16(ya[j] - yb[j]) : sat16(ya[j] + yb[j]);
            }
        }
        bD2[32 + lane] = make_uint2(mk16(d[0], d[1]), mk16(d[2], d[3]));
        __syncwarp();
        fir129_x4(bD, tapM, lane, m);

        if (a.out_mono)
            st_stream8(a.out_mono + cb * RDSP_BLK + lane * 4, make_int2((int)mk16(m[0], m[1]), (int)mk16(m[2], m[3])));
        if (a.out_stereo)
            st_stream16(a.out_stereo + cb * 2 * RDSP_BLK + lane * 8,
                        make_int4((int)mk16(m[0], m[0]), (int)mk16(m[1], m[1]), (int)mk16(m[2], m[2]), (int)mk16(m[3], m[3])));
        if (a.dbg) {
            float4 *dp = reinterpret_cast<float4 *>(a.dbg + cb * 2 * RDSP_BLK + lane * 8);
            const float f0 = (float)m[0] / 32768.0f, f1 = (float)m[1] / 32768.0f;
            const float f2 = (float)m[2] / 32768.0f, f3 = (float)m[3] / 32768.0f;
            dp[0] = make_float4(f0, f0, f1, f1);
            dp[1] = make_float4(f2, f2, f3, f3);
        }
        __syncwarp();
        // current block becomes history
        bI2[lane] = bI2[32 + lane];
        bQ2[lane] = bQ2[32 + lane];
        bD2[lane] = bD2[32 + lane];
        __syncwarp();
    }
    hrow[lane] = bI2[lane];
    hrow[32 + lane] = bQ2[lane];
    hrow[64 + lane] = bD2[lane];
}

}  // namespace

void launch_front(const FrontArgs &a, cudaStream_t st)
{
    const int grid = (a.C + WARPS - 1) / WARPS;
    k_front<<<grid, WARPS * 32, 0, st>>>(a);
}
