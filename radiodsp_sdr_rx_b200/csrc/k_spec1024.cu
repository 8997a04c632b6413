// k_spec1024.cu — K10: 1024-point q15 spectrum of the audio output L (the "audio scope").
//
// Replaces AudioAnalyzeFFT1024 (RadioDSP_SDR_RX.ino:58,87,147-148, read at RDSP_display.h:219; Teensy Audio
// library, semantics per SURVEY.md Appendix A.3): collect 8 blocks, real samples with zero imaginary part,
// Hann-1024 window (v*w)>>15, arm_cfft_radix4_q15 (1024), output[i] = sqrt_uint32_approx(re^2+im^2) for
// i < 512, then keep the last 4 blocks (50 % overlap => a new spectrum every 4 ticks from tick 7 on).
// All integer, bit-exact.
//
// Mapping: one 64-thread CTA per channel, 4 radix-4 butterflies per thread per stage on unpacked (re, im) pairs in
// shared memory (fft_q15.cuh).  The 10 KB frame is the only per-channel shared memory, so ~20 channels are
// resident per SM; twiddles and the window come through L1.  The last 8 blocks of L live in an HBM ring
// [C][8][128] indexed by tick mod 8; ticks that do not complete a frame only append their block.
#include "rdsp_common.cuh"
#include "fft_q15.cuh"
#include "kernels.h"

namespace {

constexpr int NT = 64;

__global__ void __launch_bounds__(NT) k_spec1024(Spec1024Args a)
{
    __shared__ __align__(16) int2 s_fft[1024 + 256];              // unpacked (re, im), skewed (fft_q15.cuh)
    __shared__ uint16_t s_guess[34];
    if (threadIdx.x < 33) s_guess[threadIdx.x] = c_sqrt_guess[threadIdx.x];

    const int tid = threadIdx.x;
    const int ch = a.ch0 + blockIdx.x;
    int16_t *ring = a.ring + (size_t)ch * 8 * RDSP_BLK;
    uint2 *gring = reinterpret_cast<uint2 *>(ring);               // 32 uint2 (4 samples each) per slot

    for (int t = 0; t < a.T; t++) {
        const unsigned long long tick = a.tick0 + t;
        const int slot = (int)(tick & 7ull);
        if (tid < 32) {                                             // append L of this block: 4 frames (16 bytes) per lane
            const int4 v = *reinterpret_cast<const int4 *>(a.audio + ((size_t)t * a.C + ch) * 2 * RDSP_BLK + tid * 8);
            gring[slot * 32 + tid] = make_uint2(((uint32_t)v.x & 0xFFFFu) | ((uint32_t)v.y << 16),
                                                ((uint32_t)v.z & 0xFFFFu) | ((uint32_t)v.w << 16));
        }
        if (!(tick >= 7ull && ((tick - 7ull) & 3ull) == 0ull)) continue;       // uniform over the CTA
        __syncthreads();                                            // the appended block is visible to the whole CTA

        // frame = blocks tick-7 .. tick, windowed
#pragma unroll 4
        for (int j = 0; j < 16; j++) {
            const int i = tid + NT * j;
            const int b = i >> 7, n = i & 127;
            const int32_t smp = ring[(int)((tick + 1 + b) & 7ull) * RDSP_BLK + n];
            s_fft[q15fft::P(i)] = make_int2((int16_t)((smp * (int32_t)__ldg(a.win + i)) >> 15), 0);    // imaginary part 0
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; r++) q15fft::first_real(s_fft, a.tw, 1024, 4, tid + NT * r);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; r++) q15fft::middle(s_fft, a.tw, 256, 64, 16, tid + NT * r);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; r++) q15fft::middle(s_fft, a.tw, 64, 16, 64, tid + NT * r);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; r++) q15fft::middle(s_fft, a.tw, 16, 4, 256, tid + NT * r);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; r++) q15fft::last(s_fft, tid + NT * r);
        __syncthreads();

        // bins 0..511 sit at the EVEN elements (bin = bitrev10(element)); walk the elements, stage the u16 results in
        // natural order on top of the (now dead) frame, store them coalesced
        uint16_t v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int2 w = s_fft[q15fft::P(2 * (tid + NT * j))];
            v[j] = (uint16_t)sqrt_u32_approx_fast((uint32_t)(w.x * w.x) + (uint32_t)(w.y * w.y), s_guess);
        }
        __syncthreads();
        uint16_t *s_o = reinterpret_cast<uint16_t *>(s_fft);
#pragma unroll
        for (int j = 0; j < 8; j++) s_o[__brev((unsigned)(2 * (tid + NT * j))) >> 22] = v[j];
        __syncthreads();
        reinterpret_cast<uint4 *>(a.output + (size_t)ch * 512)[tid] = reinterpret_cast<const uint4 *>(s_o)[tid];
        __syncthreads();
    }
}

}  // namespace

void launch_spec1024(const Spec1024Args &a, cudaStream_t st)
{
    RDSP_CARVEOUT_ONCE(k_spec1024);
    if (a.n > 0) k_spec1024<<<a.n, NT, 0, st>>>(a);
}
