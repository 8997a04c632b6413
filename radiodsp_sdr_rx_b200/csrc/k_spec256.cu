// k_spec256.cu — a11 + K9: DC-cleaning high-pass biquads on I and Q, then the 256-point q15 IQ spectrum.
//
// Replaces biquad1/biquad2 (RadioDSP_SDR_RX.ino:59-60,75-78,155-156; AudioFilterBiquad::update as shipped,
// SURVEY.md Appendix G.2) and AudioAnalyzeFFT256IQ::update (analyze_fft256iq.cpp:65-118): pack previous +
// current block as 256 complex q15 (:38-48,78-79), Hann window (v*w)>>15 (:50-63), radix-4 q15 FFT (:82),
// |.|^2 (:88-89), sum (+)= magsq / naverage with integer division at every update (:90,96), and after
// naverage updates output[255 - (i ^ 128)] = sqrt_uint32_approx(sum[i]) (:99-113).  All integer, bit-exact.
//
// Mapping: 16 channels per CTA.  The biquad is a saturating recurrence with error feedback (sequential in
// time), so the 32 I/Q streams of the CTA are walked by the 32 lanes of warp 0 while the block sits in
// shared memory; the FFT is one warp per channel, two radix-4 butterflies per lane per stage, in place in
// shared memory.  The 256 running sums of a channel live in registers (8 per lane) across the blocks of
// one call; the averaging counter is uniform over channels and kept on the host.
#include "rdsp_common.cuh"
#include "fft_q15.cuh"
#include "kernels.h"

namespace {

constexpr int CH = 16;                 // channels (= warps) per CTA
constexpr int SW = 66;                 // words per biquad stream row (64 + pad)

__device__ __forceinline__ int32_t smlaw(int32_t c, int32_t x16, int32_t acc)
{
    return (int32_t)((uint32_t)acc + (uint32_t)(int32_t)(((long long)c * (long long)x16) >> 16));
}

__global__ void __launch_bounds__(CH * 32) k_spec256(Spec256Args a)
{
    __shared__ uint32_t s_cur[CH][2][SW];      // current block: I stream, Q stream, two samples per word
    __shared__ uint32_t s_prev[CH][128];       // previous post-biquad block, (I | Q << 16)
    __shared__ uint32_t s_fft[CH][256];
    __shared__ int16_t s_win[256];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ch = blockIdx.x * CH + warp;
    const bool valid = ch < a.C;
    for (int i = threadIdx.x; i < 256; i += CH * 32) s_win[i] = a.win[i];

    // biquad state of stream `lane` (warp 0 only): channel blockIdx*16 + lane/2, I/Q = lane & 1
    uint32_t bprev = 0, aprev = 0;
    int32_t bsum = 0;
    const int bch = blockIdx.x * CH + (lane >> 1);
    if (warp == 0 && bch < a.C) {
        const int32_t *st = a.bq_state + ((size_t)bch * 2 + (lane & 1)) * 4;
        bprev = (uint32_t)st[0]; aprev = (uint32_t)st[1]; bsum = st[2];
    }

    uint32_t sum[8];
    if (valid) {
#pragma unroll
        for (int j = 0; j < 8; j++) sum[j] = a.sum[(size_t)ch * 256 + lane + 32 * j];
        const uint32_t *pr = reinterpret_cast<const uint32_t *>(a.prev + (size_t)ch * 2 * RDSP_BLK);
#pragma unroll
        for (int j = 0; j < 4; j++) s_prev[warp][lane + 32 * j] = pr[lane + 32 * j];
    }
    int have_prev = a.have_prev, count = a.count;
    uint32_t *fb = s_fft[warp];

    for (int t = 0; t < a.T; t++) {
        if (valid) {
            const int4 v = ld_stream16(a.iq + ((size_t)t * a.C + ch) * 2 * RDSP_BLK + lane * 8);
            const uint32_t w0 = (uint32_t)v.x, w1 = (uint32_t)v.y, w2 = (uint32_t)v.z, w3 = (uint32_t)v.w;
            s_cur[warp][0][2 * lane] = (w0 & 0xFFFFu) | (w1 << 16);
            s_cur[warp][0][2 * lane + 1] = (w2 & 0xFFFFu) | (w3 << 16);
            s_cur[warp][1][2 * lane] = (w0 >> 16) | (w1 & 0xFFFF0000u);
            s_cur[warp][1][2 * lane + 1] = (w2 >> 16) | (w3 & 0xFFFF0000u);
        }
        __syncthreads();

        if (warp == 0 && bch < a.C) {          // 32 streams, one per lane
            uint32_t *d = s_cur[lane >> 1][lane & 1];
            for (int i = 0; i < 64; i++) {
                const uint32_t in2 = d[i];
                bsum = smlaw(a.b0, lo16(in2), bsum);
                bsum = smlaw(a.b1, hi16(bprev), bsum);
                bsum = smlaw(a.b2, lo16(bprev), bsum);
                bsum = smlaw(a.a1, hi16(aprev), bsum);
                bsum = smlaw(a.a2, lo16(aprev), bsum);
                const int32_t o_lo = sat16(bsum >> 14);
                bsum &= 0x3FFF;
                bsum = smlaw(a.b0, hi16(in2), bsum);
                bsum = smlaw(a.b1, lo16(in2), bsum);
                bsum = smlaw(a.b2, hi16(bprev), bsum);
                bsum = smlaw(a.a1, o_lo, bsum);
                bsum = smlaw(a.a2, hi16(aprev), bsum);
                const int32_t o_hi = sat16(bsum >> 14);
                bsum &= 0x3FFF;
                bprev = in2;
                aprev = mk16(o_lo, o_hi);
                d[i] = aprev;
            }
        }
        __syncthreads();

        if (valid) {
            // current block as (I | Q << 16) words for samples lane + 32 j
            uint32_t cw[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int n = lane + 32 * j;
                const uint32_t wi = s_cur[warp][0][n >> 1], wq = s_cur[warp][1][n >> 1];
                const uint32_t si = (n & 1) ? (wi >> 16) : (wi & 0xFFFFu);
                const uint32_t sq = (n & 1) ? (wq >> 16) : (wq & 0xFFFFu);
                cw[j] = si | (sq << 16);
            }
            if (have_prev) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int n = lane + 32 * j;
                    const uint32_t pwd = s_prev[warp][n];
                    const int32_t w0 = s_win[n], w1 = s_win[128 + n];
                    fb[n] = mk16((lo16(pwd) * w0) >> 15, (hi16(pwd) * w0) >> 15);
                    fb[128 + n] = mk16((lo16(cw[j]) * w1) >> 15, (hi16(cw[j]) * w1) >> 15);
                }
                __syncwarp();
                q15fft::first(fb, a.tw, 256, 16, lane);
                q15fft::first(fb, a.tw, 256, 16, lane + 32);
                __syncwarp();
                q15fft::middle(fb, a.tw, 64, 16, 64, lane);
                q15fft::middle(fb, a.tw, 64, 16, 64, lane + 32);
                __syncwarp();
                q15fft::middle(fb, a.tw, 16, 4, 256, lane);
                q15fft::middle(fb, a.tw, 16, 4, 256, lane + 32);
                __syncwarp();
                q15fft::last(fb, lane);
                q15fft::last(fb, lane + 32);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int i = lane + 32 * j;
                    const uint32_t w = fb[__brev((unsigned)i) >> 24];
                    const uint32_t magsq = (uint32_t)(lo16(w) * lo16(w)) + (uint32_t)(hi16(w) * hi16(w));
                    const uint32_t q = (uint32_t)(((unsigned long long)magsq * a.div_magic) >> a.div_shift);   // magsq / naverage, exact
                    sum[j] = (count == 0) ? q : sum[j] + q;
                }
                if (++count == a.naverage) {
                    count = 0;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int i = lane + 32 * j;
                        a.output[(size_t)ch * 256 + (255 - (i ^ 128))] = (uint16_t)sqrt_u32_approx(sum[j]);
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 4; j++) s_prev[warp][lane + 32 * j] = cw[j];
        }
        have_prev = 1;
        __syncthreads();
    }

    if (valid) {
#pragma unroll
        for (int j = 0; j < 8; j++) a.sum[(size_t)ch * 256 + lane + 32 * j] = sum[j];
        uint32_t *pr = reinterpret_cast<uint32_t *>(a.prev + (size_t)ch * 2 * RDSP_BLK);
#pragma unroll
        for (int j = 0; j < 4; j++) pr[lane + 32 * j] = s_prev[warp][lane + 32 * j];
    }
    if (warp == 0 && bch < a.C) {
        int32_t *st = a.bq_state + ((size_t)bch * 2 + (lane & 1)) * 4;
        st[0] = (int32_t)bprev; st[1] = (int32_t)aprev; st[2] = bsum;
    }
}

}  // namespace

void launch_spec256(const Spec256Args &a, cudaStream_t st)
{
    k_spec256<<<(a.C + CH - 1) / CH, CH * 32, 0, st>>>(a);
}
