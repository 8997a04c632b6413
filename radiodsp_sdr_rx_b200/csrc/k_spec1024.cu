// k_spec1024.cu — K10: 1024-point q15 spectrum of the audio output L (the "audio scope").
//
// Replaces AudioAnalyzeFFT1024 (RadioDSP_SDR_RX.ino:58,87,147-148, read at RDSP_display.h:219; Teensy Audio
// library, semantics per SURVEY.md Appendix A.3): collect 8 blocks, real samples with zero imaginary part,
// Hann-1024 window (v*w)>>15, arm_cfft_radix4_q15 (1024), output[i] = sqrt_uint32_approx(re^2+im^2) for
// i < 512, then keep the last 4 blocks (50 % overlap => a new spectrum every 4 ticks from tick 7 on).
// All integer, bit-exact.
//
// Mapping: one warp per channel, 8 butterflies per lane per stage, in place in shared memory.  The last 8
// blocks of L live in an HBM ring [C][8][128] indexed by tick mod 8; ticks that do not complete a frame only
// append their block (no ring read).
#include "rdsp_common.cuh"
#include "fft_q15.cuh"
#include "kernels.h"

namespace {

constexpr int WARPS = 2;

__global__ void __launch_bounds__(WARPS * 32) k_spec1024(Spec1024Args a)
{
    __shared__ __align__(8) int2 s_fft[WARPS][1024 + 256];        // unpacked (re, im), skewed (fft_q15.cuh)
    __shared__ __align__(16) int16_t s_ring[WARPS][8][RDSP_BLK];
    __shared__ int16_t s_win[1024];
    __shared__ __align__(8) int2 s_tw[768];                       // twiddle k*4 of the 4096-table, k < 768
    __shared__ __align__(16) uint16_t s_o[WARPS][512];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ch = blockIdx.x * WARPS + warp;
    if (a.any_fft) {
        for (int i = threadIdx.x; i < 1024; i += WARPS * 32) s_win[i] = a.win[i];
        for (int i = threadIdx.x; i < 768; i += WARPS * 32) s_tw[i] = a.tw[4 * i];
        __syncthreads();
    }
    if (ch >= a.C) return;

    uint2 *gring = reinterpret_cast<uint2 *>(a.ring + (size_t)ch * 8 * RDSP_BLK);    // 32 uint2 per slot

    if (!a.any_fft) {                                   // append-only call
        for (int t = 0; t < a.T; t++) {
            const int slot = (int)((a.tick0 + t) & 7ull);
            const int4 v = *reinterpret_cast<const int4 *>(a.audio + ((size_t)t * a.C + ch) * 2 * RDSP_BLK + lane * 8);
            gring[slot * 32 + lane] = make_uint2(((uint32_t)v.x & 0xFFFFu) | ((uint32_t)v.y << 16),
                                                 ((uint32_t)v.z & 0xFFFFu) | ((uint32_t)v.w << 16));
        }
        return;
    }

    uint2 *ring2 = reinterpret_cast<uint2 *>(&s_ring[warp][0][0]);
#pragma unroll
    for (int s = 0; s < 8; s++) ring2[s * 32 + lane] = gring[s * 32 + lane];
    __syncwarp();
    int2 *fb = s_fft[warp];

    for (int t = 0; t < a.T; t++) {
        const unsigned long long tick = a.tick0 + t;
        const int slot = (int)(tick & 7ull);
        const int4 v = *reinterpret_cast<const int4 *>(a.audio + ((size_t)t * a.C + ch) * 2 * RDSP_BLK + lane * 8);
        const uint2 blk = make_uint2(((uint32_t)v.x & 0xFFFFu) | ((uint32_t)v.y << 16),
                                     ((uint32_t)v.z & 0xFFFFu) | ((uint32_t)v.w << 16));
        ring2[slot * 32 + lane] = blk;
        gring[slot * 32 + lane] = blk;
        __syncwarp();
        if (tick >= 7ull && ((tick - 7ull) & 3ull) == 0ull) {
            // frame = blocks tick-7 .. tick
#pragma unroll 4
            for (int j = 0; j < 32; j++) {
                const int i = lane + 32 * j;                       // frame sample index
                const int b = i >> 7, n = i & 127;
                const int32_t smp = s_ring[warp][(int)((tick + 1 + b) & 7ull)][n];
                fb[q15fft::P(i)] = make_int2((int16_t)((smp * (int32_t)s_win[i]) >> 15), 0);     // imaginary part 0
            }
            __syncwarp();
#pragma unroll 2
            for (int r = 0; r < 8; r++) q15fft::first(fb, s_tw, 1024, 1, lane + 32 * r);      // steps in units of the 1024-table
            __syncwarp();
#pragma unroll 2
            for (int r = 0; r < 8; r++) q15fft::middle(fb, s_tw, 256, 64, 4, lane + 32 * r);
            __syncwarp();
#pragma unroll 2
            for (int r = 0; r < 8; r++) q15fft::middle(fb, s_tw, 64, 16, 16, lane + 32 * r);
            __syncwarp();
#pragma unroll 2
            for (int r = 0; r < 8; r++) q15fft::middle(fb, s_tw, 16, 4, 64, lane + 32 * r);
            __syncwarp();
#pragma unroll 2
            for (int r = 0; r < 8; r++) q15fft::last(fb, lane + 32 * r);
            __syncwarp();
            // bins 0..511 sit at the EVEN elements (bin = bitrev10(element)); walk the elements, stage the u16 results
            // in natural order, store them coalesced
#pragma unroll 4
            for (int j = 0; j < 16; j++) {
                const int e = 2 * (lane + 32 * j);
                const int2 w = fb[q15fft::P(e)];
                const uint32_t magsq = (uint32_t)(w.x * w.x) + (uint32_t)(w.y * w.y);
                s_o[warp][__brev((unsigned)e) >> 22] = (uint16_t)sqrt_u32_approx(magsq);
            }
            __syncwarp();
            {
                const uint4 *so = reinterpret_cast<const uint4 *>(s_o[warp]);
                uint4 *go = reinterpret_cast<uint4 *>(a.output + (size_t)ch * 512);
                go[lane] = so[lane];
                go[lane + 32] = so[lane + 32];
            }
            __syncwarp();
        }
    }
}

}  // namespace

void launch_spec1024(const Spec1024Args &a, cudaStream_t st)
{
    k_spec1024<<<(a.C + WARPS - 1) / WARPS, WARPS * 32, 0, st>>>(a);
}
