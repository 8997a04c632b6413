// k_spec256.cu — K9: the 256-point q15 IQ spectrum (RF panadapter / S-meter source).
//
// Replaces AudioAnalyzeFFT256IQ::update (analyze_fft256iq.cpp:65-118): pack previous + current block as 256
// complex q15 (:38-48,78-79), Hann window (v*w)>>15 (:50-63), radix-4 q15 FFT (:82), |.|^2 (:88-89),
// sum (+)= magsq / naverage with integer division at every update (:90,96), and after naverage updates
// output[255 - (i ^ 128)] = sqrt_uint32_approx(sum[i]) (:99-113).  All integer, bit-exact.  Its input is the
// high-passed IQ produced by k_biquad (the graph wiring of RadioDSP_SDR_RX.ino:75-78).
//
// Mapping: a half-warp per channel, 16 elements per lane, the four radix-4 stages as TWO register passes with one
// shared-memory exchange between them (the butterflies are those of fft_q15.cuh, so the result is bit-identical to
// the stage-by-stage form this kernel had before; 82 -> 71 us per 8-block launch of 8192 channels):
//   pass A  lane l holds x[d3][d2] = element 64 d3 + 16 d2 + l: window, stage 1 over d3, stage 2 over d2;
//   pass B  lane l holds y[d1][d0] = element 16 l + 4 d1 + d0: stage 3 over d1, last stage over d0, |.|^2, sums.
// The 256 running sums of a channel live in registers (16 per lane, indexed by ELEMENT = bit-reversed bin) and the
// previous block in 8 registers across the blocks of one call; the averaging counter is uniform over channels and
// kept on the host.  Element e sits at e + 2 (e >> 4) in the exchange buffer: the 16-byte reads of pass B (lanes
// 144 bytes apart) and the 8-byte writes of pass A (lanes 8 bytes apart) are both conflict free.
#include "rdsp_common.cuh"
#include "fft_q15.cuh"
#include "kernels.h"

namespace {

constexpr int WARPS = 4;                      // 8 channels per CTA
constexpr int XBUF = 256 + 32;                // int2 per channel in the exchange buffer

__device__ __forceinline__ int PX(int e) { return e + 2 * (e >> 4); }

#ifndef RDSP_SPEC256_MINB
#define RDSP_SPEC256_MINB 4      // 128 registers (112 B of spills): 71 -> 74.5 us alone, but the STEP gains 2.6 % — a fourth CTA per SM is co-residency for the kernels running beside it (r02 A/B, 3 / 4 / 5: 510.6 / 497.5 / 510.5 us per step)
#endif
__global__ void __launch_bounds__(WARPS * 32, RDSP_SPEC256_MINB) k_spec256(Spec256Args a)
{
    __shared__ __align__(16) int2 s_x[WARPS * 2][XBUF];
    __shared__ int16_t s_win[256];
    __shared__ __align__(8) int2 s_tw[192];                       // twiddle k*16 of the 4096-table, k < 192
    __shared__ uint16_t s_guess[34];

    for (int i = threadIdx.x; i < 256; i += WARPS * 32) s_win[i] = a.win[i];
    for (int i = threadIdx.x; i < 192; i += WARPS * 32) s_tw[i] = a.tw[16 * i];
    if (threadIdx.x < 33) s_guess[threadIdx.x] = c_sqrt_guess[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l = lane & 15, hw = lane >> 4;
    const int lc = (blockIdx.x * WARPS + warp) * 2 + hw;
    const bool active = lc < a.n;
    const int ch = a.ch0 + (active ? lc : a.n - 1);                // idle half-warps shadow the last channel, stores masked
    int2 *xb = s_x[warp * 2 + hw];

    // sums by element 16 l + k (pass B order); previous block and window by element 64 d3 + 16 d2 + l (pass A order)
    uint32_t sum[16], pw[8];
#pragma unroll
    for (int k = 0; k < 16; k++) sum[k] = a.sum[(size_t)ch * 256 + (__brev((unsigned)(16 * l + k)) >> 24)];
    {
        const uint32_t *pr = reinterpret_cast<const uint32_t *>(a.prev + (size_t)ch * 2 * RDSP_BLK);
#pragma unroll
        for (int q = 0; q < 8; q++) pw[q] = pr[16 * q + l];        // sample 64 d3 + 16 d2 + l of the previous block, q = 4 d3 + d2
    }
    uint32_t wp[8];                                               // window of elements 16 q + l (low half) and 16 (8 + q) + l (high half)
#pragma unroll
    for (int q = 0; q < 8; q++) wp[q] = mk16(s_win[16 * q + l], s_win[16 * (8 + q) + l]);
    int have_prev = a.tick_in->have_prev, count = a.tick_in->count;

    for (int t = 0; t < a.T; t++) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.iq + ((size_t)t * a.C + ch) * 2 * RDSP_BLK);
        uint32_t cw[8];                                    // (I | Q << 16) for samples 16 q + l of this block
#pragma unroll
        for (int q = 0; q < 8; q++) cw[q] = src[16 * q + l];
        if (have_prev) {
            // ---- pass A: elements 64 d3 + 16 d2 + l; d3 = 0,1 previous block, d3 = 2,3 this block
            int2 x[4][4];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                // (v * w) >> 15 is stored back into an int16 by the reference; with 0 <= w <= 32767 (make_hann_q15) it
                // lies in [-32767, 32766], so the narrowing changes nothing and is not spelled out
                const int32_t w0 = lo16(wp[q]), w1 = hi16(wp[q]);
                x[q >> 2][q & 3] = make_int2((lo16(pw[q]) * w0) >> 15, (hi16(pw[q]) * w0) >> 15);
                x[2 + (q >> 2)][q & 3] = make_int2((lo16(cw[q]) * w1) >> 15, (hi16(cw[q]) * w1) >> 15);
            }
#pragma unroll
            for (int d2 = 0; d2 < 4; d2++) {                        // stage 1: span 64, butterfly i = 16 d2 + l, twiddle step 1
                const int ic = 16 * d2 + l;
                q15fft::first_r(x[0][d2], x[1][d2], x[2][d2], x[3][d2], s_tw[ic], s_tw[2 * ic], s_tw[3 * ic]);
            }
            {
                const int ic = 4 * l;                               // stage 2: span 16, j = l, twiddle step 4
                const int2 t1 = s_tw[ic], t2 = s_tw[2 * ic], t3 = s_tw[3 * ic];
#pragma unroll
                for (int d3 = 0; d3 < 4; d3++) q15fft::middle_r(x[d3][0], x[d3][1], x[d3][2], x[d3][3], t1, t2, t3);
            }
#pragma unroll
            for (int d3 = 0; d3 < 4; d3++)
#pragma unroll
                for (int d2 = 0; d2 < 4; d2++) xb[PX(64 * d3 + 16 * d2 + l)] = x[d3][d2];
            __syncwarp();
            // ---- pass B: elements 16 l + 4 d1 + d0
            int2 y[4][4];
            {
                const int4 *yp = reinterpret_cast<const int4 *>(&xb[PX(16 * l)]);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int4 v = yp[k];
                    y[(2 * k) >> 2][(2 * k) & 3] = make_int2(v.x, v.y);
                    y[(2 * k + 1) >> 2][(2 * k + 1) & 3] = make_int2(v.z, v.w);
                }
            }
            __syncwarp();                                            // the buffer may be rewritten by the next block
#pragma unroll
            for (int d0 = 0; d0 < 4; d0++) {                        // stage 3: span 4, j = d0, twiddle step 16
                q15fft::middle_r(y[0][d0], y[1][d0], y[2][d0], y[3][d0], a.tw3[d0][0], a.tw3[d0][1], a.tw3[d0][2]);
            }
#pragma unroll
            for (int d1 = 0; d1 < 4; d1++) q15fft::last_r(y[d1][0], y[d1][1], y[d1][2], y[d1][3]);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int2 v = y[k >> 2][k & 3];                    // element 16 l + k = bin bitrev8(16 l + k)
                const uint32_t magsq = (uint32_t)(v.x * v.x) + (uint32_t)(v.y * v.y);
                const uint32_t q = (uint32_t)(((unsigned long long)magsq * a.div_magic) >> a.div_shift);   // magsq / naverage, exact
                sum[k] = (count == 0) ? q : sum[k] + q;
            }
            if (++count == a.naverage) {
                count = 0;
                if (active) {
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        const int i = (int)(__brev((unsigned)(16 * l + k)) >> 24);
                        a.output[(size_t)ch * 256 + (255 - (i ^ 128))] = (uint16_t)sqrt_u32_approx_fast(sum[k], s_guess);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 8; q++) pw[q] = cw[q];
        have_prev = 1;
    }

    if (blockIdx.x == 0 && threadIdx.x == 0) { a.tick_out->have_prev = have_prev; a.tick_out->count = count; }
    if (active) {
#pragma unroll
        for (int k = 0; k < 16; k++) a.sum[(size_t)ch * 256 + (__brev((unsigned)(16 * l + k)) >> 24)] = sum[k];
        uint32_t *pr = reinterpret_cast<uint32_t *>(a.prev + (size_t)ch * 2 * RDSP_BLK);
#pragma unroll
        for (int q = 0; q < 8; q++) pr[16 * q + l] = pw[q];
    }
}

}  // namespace

void launch_spec256(const Spec256Args &a, cudaStream_t st)
{
    RDSP_CARVEOUT_ONCE(k_spec256);
    const int cpb = WARPS * 2;
    if (a.n > 0) k_spec256<<<(a.n + cpb - 1) / cpb, WARPS * 32, 0, st>>>(a);
}
