// k_spec256.cu — K9: the 256-point q15 IQ spectrum (RF panadapter / S-meter source).
//
// Replaces AudioAnalyzeFFT256IQ::update (analyze_fft256iq.cpp:65-118): pack previous + current block as 256
// complex q15 (:38-48,78-79), Hann window (v*w)>>15 (:50-63), radix-4 q15 FFT (:82), |.|^2 (:88-89),
// sum (+)= magsq / naverage with integer division at every update (:90,96), and after naverage updates
// output[255 - (i ^ 128)] = sqrt_uint32_approx(sum[i]) (:99-113).  All integer, bit-exact.  Its input is the
// high-passed IQ produced by k_biquad (the graph wiring of RadioDSP_SDR_RX.ino:75-78).
//
// Mapping: one warp per channel, two radix-4 butterflies per lane per stage, in place in shared memory.
// The 256 running sums of a channel live in registers (8 per lane) and the previous block in 4 registers
// across the blocks of one call; the averaging counter is uniform over channels and kept on the host.
#include "rdsp_common.cuh"
#include "fft_q15.cuh"
#include "kernels.h"

namespace {

constexpr int WARPS = 8;

__global__ void __launch_bounds__(WARPS * 32, 3) k_spec256(Spec256Args a)
{
    __shared__ __align__(8) int2 s_fft[WARPS][256 + 64];          // unpacked (re, im), skewed (fft_q15.cuh)
    __shared__ int16_t s_win[256];
    __shared__ __align__(8) int2 s_tw[192];                       // twiddle k*16 of the 4096-table, k < 192

    for (int i = threadIdx.x; i < 256; i += WARPS * 32) s_win[i] = a.win[i];
    for (int i = threadIdx.x; i < 192; i += WARPS * 32) s_tw[i] = a.tw[16 * i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lc = blockIdx.x * WARPS + warp;
    if (lc >= a.n) return;
    const int ch = a.ch0 + lc;

    // register j holds bin bitrev8(lane + 32 j): the FFT leaves bin i at element bitrev(i), so walking the ELEMENTS
    // lane + 32 j keeps the shared-memory reads of the |.|^2 loop contiguous (the sums are only loaded / stored once per call)
    uint32_t sum[8], pw[4];
#pragma unroll
    for (int j = 0; j < 8; j++) sum[j] = a.sum[(size_t)ch * 256 + (__brev((unsigned)(lane + 32 * j)) >> 24)];
    {
        const uint32_t *pr = reinterpret_cast<const uint32_t *>(a.prev + (size_t)ch * 2 * RDSP_BLK);
#pragma unroll
        for (int j = 0; j < 4; j++) pw[j] = pr[lane + 32 * j];
    }
    int32_t w0[4], w1[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { w0[j] = s_win[lane + 32 * j]; w1[j] = s_win[128 + lane + 32 * j]; }
    int have_prev = a.have_prev, count = a.count;
    int2 *fb = s_fft[warp];

    for (int t = 0; t < a.T; t++) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.iq + ((size_t)t * a.C + ch) * 2 * RDSP_BLK);
        uint32_t cw[4];                                    // (I | Q << 16) for samples lane + 32 j
#pragma unroll
        for (int j = 0; j < 4; j++) cw[j] = src[lane + 32 * j];
        if (have_prev) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int n = lane + 32 * j;
                // (v * w) >> 15 stored back into an int16 by the reference: keep the low 16 bits, sign-extended
                fb[q15fft::P(n)] = make_int2((int16_t)((lo16(pw[j]) * w0[j]) >> 15), (int16_t)((hi16(pw[j]) * w0[j]) >> 15));
                fb[q15fft::P(128 + n)] = make_int2((int16_t)((lo16(cw[j]) * w1[j]) >> 15), (int16_t)((hi16(cw[j]) * w1[j]) >> 15));
            }
            __syncwarp();
            q15fft::first(fb, s_tw, 256, 1, lane);                 // twiddle steps in units of the 256-point table
            q15fft::first(fb, s_tw, 256, 1, lane + 32);
            __syncwarp();
            q15fft::middle(fb, s_tw, 64, 16, 4, lane);
            q15fft::middle(fb, s_tw, 64, 16, 4, lane + 32);
            __syncwarp();
            q15fft::middle(fb, s_tw, 16, 4, 16, lane);
            q15fft::middle(fb, s_tw, 16, 4, 16, lane + 32);
            __syncwarp();
            q15fft::last(fb, lane);
            q15fft::last(fb, lane + 32);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int2 w = fb[q15fft::P(lane + 32 * j)];           // = bin bitrev8(lane + 32 j)
                const uint32_t magsq = (uint32_t)(w.x * w.x) + (uint32_t)(w.y * w.y);
                const uint32_t q = (uint32_t)(((unsigned long long)magsq * a.div_magic) >> a.div_shift);   // magsq / naverage, exact
                sum[j] = (count == 0) ? q : sum[j] + q;
            }
            if (++count == a.naverage) {
                count = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int i = (int)(__brev((unsigned)(lane + 32 * j)) >> 24);
                    a.output[(size_t)ch * 256 + (255 - (i ^ 128))] = (uint16_t)sqrt_u32_approx(sum[j]);
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 4; j++) pw[j] = cw[j];
        have_prev = 1;
    }

#pragma unroll
    for (int j = 0; j < 8; j++) a.sum[(size_t)ch * 256 + (__brev((unsigned)(lane + 32 * j)) >> 24)] = sum[j];
    uint32_t *pr = reinterpret_cast<uint32_t *>(a.prev + (size_t)ch * 2 * RDSP_BLK);
#pragma unroll
    for (int j = 0; j < 4; j++) pr[lane + 32 * j] = pw[j];
}

}  // namespace

void launch_spec256(const Spec256Args &a, cudaStream_t st)
{
    RDSP_CARVEOUT_ONCE(k_spec256);
    if (a.n > 0) k_spec256<<<(a.n + WARPS - 1) / WARPS, WARPS * 32, 0, st>>>(a);
}
