// k_spec1024.cu — K10: 1024-point q15 spectrum of the audio output L (the "audio scope").
//
// Replaces AudioAnalyzeFFT1024 (RadioDSP_SDR_RX.ino:58,87,147-148, read at RDSP_display.h:219; Teensy Audio
// library, semantics per SURVEY.md Appendix A.3): collect 8 blocks, real samples with zero imaginary part,
// Hann-1024 window (v*w)>>15, arm_cfft_radix4_q15 (1024), output[i] = sqrt_uint32_approx(re^2+im^2) for
// i < 512, then keep the last 4 blocks (50 % overlap => a new spectrum every 4 ticks from tick 7 on).
// All integer, bit-exact.
//
// Mapping: one 64-thread CTA per channel.  The five radix-4 stages run as THREE passes: a thread holds 16 elements in
// registers and runs two consecutive stages on them (stages 1+2 on the elements that differ in index digits 4,3 — read
// straight from the HBM ring and windowed, the frame is never staged; stages 3+4 on digits 2,1), then the last stage and
// the magnitudes.  Two shared-memory round trips and barriers instead of five (the per-butterfly arithmetic is the one of
// fft_q15.cuh, so the result is bit-identical); element i sits at i + 4 (i >> 6), which keeps all three access patterns
// conflict free.  Twiddles and the window come through L1.  The last 8 blocks of L live in an HBM ring [C][8][128]
// indexed by tick mod 8; ticks that do not complete a frame only append their block.  In one-block calls (the sketch's calling
// pattern) the kernels that emit the audio append the row themselves (k_fftfilt / k_nlms, `ring`) and this kernel is launched
// as k_spec1024<true>; on the three ticks out of four that complete no frame the host (which mirrors the cadence) launches ONE
// CTA of it, which advances the tick counter and returns (r02f: a non-frame tick 100 -> 90 us, the mean over four ticks 103 -> 100.7 us).
#include "rdsp_common.cuh"
#include "fft_q15.cuh"
#include "kernels.h"

namespace {

constexpr int NT = 64;
__device__ __forceinline__ int PP(int i) { return i + 4 * (i >> 6); }

// stages 1..5 and the magnitudes of one frame whose windowed samples sit in x[d4][d3] = element 256 d4 + 64 d3 + tid
template <bool SAT>
__device__ __forceinline__ void fft1024_passes(int2 (&x)[4][4], int2 *s_fft, const Spec1024Args &a, const uint16_t *s_guess, int tid, uint16_t (&v)[8])
{
#pragma unroll
    for (int d3 = 0; d3 < 4; d3++) {                            // stage 1: span 256, twiddle step 4, j = 64 d3 + tid
        const int ic = 4 * (64 * d3 + tid);
        q15fft::first_real_r(x[0][d3], x[1][d3], x[2][d3], x[3][d3], a.tw[ic], a.tw[2 * ic], a.tw[3 * ic]);
    }
    {
        const int ic = 16 * tid;                                // stage 2: span 64, twiddle step 16, j = tid
        const int2 t1 = a.tw[ic], t2 = a.tw[2 * ic], t3 = a.tw[3 * ic];
#pragma unroll
        for (int d4 = 0; d4 < 4; d4++) q15fft::middle_r<SAT>(x[d4][0], x[d4][1], x[d4][2], x[d4][3], t1, t2, t3);
    }
#pragma unroll
    for (int d4 = 0; d4 < 4; d4++)
#pragma unroll
        for (int d3 = 0; d3 < 4; d3++) s_fft[PP(256 * d4 + 64 * d3 + tid)] = x[d4][d3];
    __syncthreads();
    // ---- pass 2: stages 3 + 4 on y[d2][d1] = element base + 16 d2 + 4 d1, base = 64 (tid >> 2) + (tid & 3)
    {
        const int base = 64 * (tid >> 2) + (tid & 3), d0 = tid & 3;
        int2 y[4][4];
#pragma unroll
        for (int d2 = 0; d2 < 4; d2++)
#pragma unroll
            for (int d1 = 0; d1 < 4; d1++) y[d2][d1] = s_fft[PP(base + 16 * d2 + 4 * d1)];
#pragma unroll
        for (int d1 = 0; d1 < 4; d1++) {                        // stage 3: span 16, twiddle step 64, j = 4 d1 + d0
            const int ic = 64 * (4 * d1 + d0);
            q15fft::middle_r<SAT>(y[0][d1], y[1][d1], y[2][d1], y[3][d1], a.tw[ic], a.tw[2 * ic], a.tw[3 * ic]);
        }
        {
            const int ic = 256 * d0;                            // stage 4: span 4, twiddle step 256, j = d0
            const int2 t1 = a.tw[ic], t2 = a.tw[2 * ic], t3 = a.tw[3 * ic];
#pragma unroll
            for (int d2 = 0; d2 < 4; d2++) q15fft::middle_r<SAT>(y[d2][0], y[d2][1], y[d2][2], y[d2][3], t1, t2, t3);
        }
#pragma unroll
        for (int d2 = 0; d2 < 4; d2++)
#pragma unroll
            for (int d1 = 0; d1 < 4; d1++) s_fft[PP(base + 16 * d2 + 4 * d1)] = y[d2][d1];
    }
    __syncthreads();
    // ---- pass 3: last stage on elements 4b .. 4b+3, b = tid + 64 r; bins 0..511 are the EVEN elements
    // (bin = bitrev10(element)): magnitudes straight from the registers
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int e = 4 * (tid + NT * r);
        const int4 lo = *reinterpret_cast<const int4 *>(&s_fft[PP(e)]), hi = *reinterpret_cast<const int4 *>(&s_fft[PP(e) + 2]);
        int2 w0 = make_int2(lo.x, lo.y), w1 = make_int2(lo.z, lo.w), w2 = make_int2(hi.x, hi.y), w3 = make_int2(hi.z, hi.w);
        q15fft::last_r<SAT>(w0, w1, w2, w3);
        v[2 * r] = (uint16_t)sqrt_u32_approx_fast((uint32_t)(w0.x * w0.x) + (uint32_t)(w0.y * w0.y), s_guess);
        v[2 * r + 1] = (uint16_t)sqrt_u32_approx_fast((uint32_t)(w2.x * w2.x) + (uint32_t)(w2.y * w2.y), s_guess);
    }
}

#ifndef RDSP_SPEC1024_MINB
#define RDSP_SPEC1024_MINB 16    // 64 registers, no spills (r02b A/B on one box, cfg5 step: 10 / 12 / 14 / 16 CTAs per SM = 417.7 / 416.7 / 414.3 / 413.6 us; alone 96.9 / 94.4 / 92.1 / 91.6 us)
#endif
// APPENDED: one-block calls, where the kernels that emitted the audio appended the row to the ring themselves (`ring` of FftFiltArgs /
// NlmsArgs): nothing to do on the three ticks out of four that complete no frame
template <bool APPENDED>
__global__ void __launch_bounds__(NT, RDSP_SPEC1024_MINB) k_spec1024(Spec1024Args a)
{
    __shared__ __align__(16) int2 s_fft[1024 + 64];               // unpacked (re, im), element i at PP(i)
    __shared__ uint16_t s_guess[34];
    __shared__ int s_amax[2];
    if (threadIdx.x < 33) s_guess[threadIdx.x] = c_sqrt_guess[threadIdx.x];

    const int tid = threadIdx.x;
    const int ch = a.list ? a.list[blockIdx.x] : a.ch0 + blockIdx.x;
    int16_t *ring = a.ring + (size_t)ch * 8 * RDSP_BLK;
    uint2 *gring = reinterpret_cast<uint2 *>(ring);               // 32 uint2 (4 samples each) per slot

    const unsigned long long tick0 = a.tick_in->tick;
    if (blockIdx.x == 0 && tid == 0) a.tick_out->tick = tick0 + (unsigned long long)a.T;
    pdl_wait_predecessor();                                       // the audio rows are the predecessor's output
    // (before the early return, not after: whatever follows this kernel on the stream — the join, the copy out — is ordered behind
    // THIS grid only, and a programmatic dependent that returned without waiting could complete before the kernel in front of it)
    if (APPENDED && !(tick0 >= 7ull && ((tick0 - 7ull) & 3ull) == 0ull)) return;
    for (int t = 0; t < a.T; t++) {
        const unsigned long long tick = tick0 + t;
        const int slot = (int)(tick & 7ull);
        if (tid < 32 && !APPENDED) {                                // append L of this block: 4 frames per lane
            if (a.audio_mono) {
                gring[slot * 32 + tid] = *reinterpret_cast<const uint2 *>(a.audio + ((size_t)t * a.C + ch) * RDSP_BLK + tid * 4);
            } else {
                const int4 v = *reinterpret_cast<const int4 *>(a.audio + ((size_t)t * a.C + ch) * 2 * RDSP_BLK + tid * 8);
                gring[slot * 32 + tid] = make_uint2(((uint32_t)v.x & 0xFFFFu) | ((uint32_t)v.y << 16),
                                                    ((uint32_t)v.z & 0xFFFFu) | ((uint32_t)v.w << 16));
            }
        }
        if (!(tick >= 7ull && ((tick - 7ull) & 3ull) == 0ull)) continue;       // uniform over the CTA
        __syncthreads();                                            // the appended block is visible to the whole CTA

        // ---- pass 1: stages 1 + 2 on x[d4][d3] = element 256 d4 + 64 d3 + tid, frame = blocks tick-7 .. tick, windowed
        int2 x[4][4];
#pragma unroll
        for (int d4 = 0; d4 < 4; d4++)
#pragma unroll
            for (int d3 = 0; d3 < 4; d3++) {
                const int i = 256 * d4 + 64 * d3 + tid;
                const int b = i >> 7, n = i & 127;
                const int32_t smp = ring[(int)((tick + 1 + b) & 7ull) * RDSP_BLK + n];
                // the int16 store of the reference changes nothing: 0 <= window <= 32767 (make_hann_q15) keeps the value in int16
                x[d4][d3] = make_int2((smp * (int32_t)__ldg(a.win + i)) >> 15, 0);                 // imaginary part 0
            }
        // a frame that provably cannot saturate (fft_q15.cuh) takes the butterflies without the min / max pairs
        int amax = 0;
#pragma unroll
        for (int d4 = 0; d4 < 4; d4++)
#pragma unroll
            for (int d3 = 0; d3 < 4; d3++) amax = max(amax, abs(x[d4][d3].x));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        if ((tid & 31) == 0) s_amax[tid >> 5] = amax;
        __syncthreads();
        const bool no_sat = max(s_amax[0], s_amax[1]) <= q15fft::NO_SAT_BOUND_1024_REAL;
        uint16_t v[8];
        if (no_sat) fft1024_passes<false>(x, s_fft, a, s_guess, tid, v);
        else fft1024_passes<true>(x, s_fft, a, s_guess, tid, v);
        __syncthreads();
        // stage the u16 results in natural order on top of the (now dead) frame, store them coalesced
        uint16_t *s_o = reinterpret_cast<uint16_t *>(s_fft);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const unsigned e = 4u * (unsigned)(tid + NT * r);
            s_o[__brev(e) >> 22] = v[2 * r];
            s_o[__brev(e + 2u) >> 22] = v[2 * r + 1];
        }
        __syncthreads();
        reinterpret_cast<uint4 *>(a.output + (size_t)ch * 512)[tid] = reinterpret_cast<const uint4 *>(s_o)[tid];
        __syncthreads();
    }
}

}  // namespace

void launch_spec1024(const Spec1024Args &a, cudaStream_t st)
{
    RDSP_CARVEOUT_ONCE(k_spec1024<false>); RDSP_CARVEOUT_ONCE(k_spec1024<true>);
    if (a.n <= 0) return;
    if (a.appended) rdsp_launch(k_spec1024<true>, a.n, NT, 0, st, a.pdl != 0, a);
    else rdsp_launch(k_spec1024<false>, a.n, NT, 0, st, a.pdl != 0, a);
}
