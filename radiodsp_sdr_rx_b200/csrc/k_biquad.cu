// k_biquad.cu — a11: the two DC-cleaning high-pass biquads in front of the IQ spectrum.
//
// Replaces biquad1 / biquad2 = AudioFilterBiquad::setHighpass(0, 500, 0.5) on the I and the Q stream
// (RadioDSP_SDR_RX.ino:59-60,75-78,155-156).  The update loop is the one shipped in the reference's firmware
// (SURVEY.md Appendix G.2): Q30 coefficients, 32x16 MACs that keep the top 32 bits (SMLAWB/SMLAWT), 14-bit
// error feedback carried in `sum`, output SSAT16(sum >> 14).  Integer, bit-exact.
//
// A saturating recurrence with error feedback is sequential in time, so one THREAD owns one stream
// (channel x {I,Q}) with b/a history and the error accumulator in registers; a warp owns 16 channels (every
// lane walks).  The kernel sits on the spectrum branch of the graph, which runs beside the (longer) audio branch,
// so what counts is not its latency but how few issue slots it takes from the kernels running next to it.
// Rows are staged with double-buffered 16-byte cp.async copies (chunk-rotated per row, conflict free), the I
// and Q lanes of a channel re-interleave their outputs with one shuffle, and results leave as coalesced
// 16-byte stores.
//   SMLAWx(c, x) = (c * x) >> 16 is computed as __mulhi(c, x << 16): one IMAD.HI.
#include "rdsp_common.cuh"
#include "kernels.h"

namespace {

constexpr int R = 16;                                  // channels per warp: all 32 lanes walk a stream

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// acc + ((c * (int16)x) >> 16), x already positioned in the top half of a word
__device__ __forceinline__ int32_t mlaw(int32_t c, uint32_t x_top, int32_t acc)
{
    return (int32_t)((uint32_t)acc + (uint32_t)__mulhi(c, (int32_t)x_top));
}

__global__ void __launch_bounds__(32) k_biquad(BiquadArgs a)
{
    __shared__ __align__(16) unsigned char s_in[2][R][512];
    __shared__ __align__(16) unsigned char s_out[R][512];

    const int lane = threadIdx.x;
    const int row = (lane >> 1) & (R - 1), iq = lane & 1;   // stream = (channel row, I or Q); lanes >= 2R mirror
    const int ch0 = a.ch0 + blockIdx.x * R, ch_end = a.ch0 + a.n;
    const int myc = ch0 + row;
    const bool walker = lane < 2 * R && myc < ch_end;

    uint32_t bprev = 0, aprev = 0;                          // two previous inputs / outputs: lo = older, hi = newer
    int32_t sum = 0;
    if (walker) {
        const int32_t *st = a.state + ((size_t)myc * 2 + iq) * 4;
        bprev = (uint32_t)st[0]; aprev = (uint32_t)st[1]; sum = st[2];
    }
    const int32_t b0 = a.b0, b1 = a.b1, b2 = a.b2, a1 = a.a1, a2 = a.a2;
    const int32_t a1h = a1 >> 16, a1l = a1 & 0xFFFF;       // a1 = 65536 a1h + a1l, a1l unsigned
    const unsigned sh = iq ? 0u : 16u;                      // brings this stream's half-word to the top

    auto issue_load = [&](int t, int buf) {
#pragma unroll
        for (int r = 0; r < R; r++)
            if (ch0 + r < ch_end)
                cp_async16(&s_in[buf][r][((lane + r) & 31) * 16],
                           reinterpret_cast<const unsigned char *>(a.iq + ((size_t)t * a.C + ch0 + r) * 2 * RDSP_BLK) + lane * 16);
        cp_async_commit();
    };

    issue_load(0, 0);
    for (int t = 0; t < a.T; t++) {
        const int buf = t & 1;
        if (t + 1 < a.T) { issue_load(t + 1, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncwarp();

        if (lane < 2 * R) {
            // 4 frames (16 bytes) per iteration: two passes of the 2-sample update loop
#pragma unroll 2
            for (int j = 0; j < 32; j++) {
                const uint4 v = *reinterpret_cast<const uint4 *>(&s_in[buf][row][((j + row) & 31) * 16]);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};     // frame = (I | Q << 16)
                uint32_t o[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t x0 = (w[2 * h] << sh) & 0xFFFF0000u;        // earlier sample, in the top half
                    const uint32_t x1 = (w[2 * h + 1] << sh) & 0xFFFF0000u;    // later sample
                    // The accumulator wraps mod 2^32, so the ten MACs of a sample pair may be added in any order: everything
                    // that does not depend on the newest output is summed OFF the recurrence (pre0, pre1), and what stays
                    // on it per sample is a1 * (previous output) -> add -> >> 14 -> saturate.  That one product is done as
                    // two full-rate IMADs ((a1 o) >> 16 = a1h o + ((a1l o) >> 16), exact) instead of a quarter-rate IMAD.HI.
                    const uint32_t pre0 = (uint32_t)__mulhi(b0, (int32_t)x0) + (uint32_t)__mulhi(b1, (int32_t)(bprev & 0xFFFF0000u)) +
                                          (uint32_t)__mulhi(b2, (int32_t)(bprev << 16)) + (uint32_t)__mulhi(a2, (int32_t)(aprev << 16));
                    const uint32_t pre1 = (uint32_t)__mulhi(b0, (int32_t)x1) + (uint32_t)__mulhi(b1, (int32_t)x0) +
                                          (uint32_t)__mulhi(b2, (int32_t)(bprev & 0xFFFF0000u)) + (uint32_t)__mulhi(a2, (int32_t)(aprev & 0xFFFF0000u));
                    const int32_t o_prev = (int32_t)aprev >> 16;               // the newest output so far
                    sum = (int32_t)((uint32_t)sum + pre0 + (uint32_t)(a1h * o_prev + ((a1l * o_prev) >> 16)));
                    const int32_t o_lo = sat16(sum >> 14);
                    sum &= 0x3FFF;
                    sum = (int32_t)((uint32_t)sum + pre1 + (uint32_t)(a1h * o_lo + ((a1l * o_lo) >> 16)));
                    const int32_t o_hi = sat16(sum >> 14);
                    sum &= 0x3FFF;
                    bprev = (x0 >> 16) | x1;
                    aprev = mk16(o_lo, o_hi);
                    o[h] = aprev;
                }
                // re-interleave: the I lane has (I0|I1), (I2|I3); the Q lane has (Q0|Q1), (Q2|Q3)
                const uint32_t p0 = __shfl_xor_sync(0xffffffffu, o[0], 1), p1 = __shfl_xor_sync(0xffffffffu, o[1], 1);
                uint2 outw;
                if (iq == 0) outw = make_uint2((o[0] & 0xFFFFu) | (p0 << 16), (o[0] >> 16) | (p0 & 0xFFFF0000u));      // frames 0, 1
                else         outw = make_uint2((p1 & 0xFFFFu) | (o[1] << 16), (p1 >> 16) | (o[1] & 0xFFFF0000u));      // frames 2, 3
                *reinterpret_cast<uint2 *>(&s_out[row][((j + row) & 31) * 16 + iq * 8]) = outw;
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < R; r++)
            if (ch0 + r < ch_end)
                *reinterpret_cast<int4 *>(reinterpret_cast<unsigned char *>(a.out + ((size_t)t * a.C + ch0 + r) * 2 * RDSP_BLK) + lane * 16) =
                    *reinterpret_cast<const int4 *>(&s_out[r][((lane + r) & 31) * 16]);
        __syncwarp();
    }
    if (walker) {
        int32_t *st = a.state + ((size_t)myc * 2 + iq) * 4;
        st[0] = (int32_t)bprev; st[1] = (int32_t)aprev; st[2] = sum;
    }
}

}  // namespace

void launch_biquad(const BiquadArgs &a, cudaStream_t st)
{
    RDSP_CARVEOUT_ONCE(k_biquad);
    if (a.n > 0) k_biquad<<<(a.n + R - 1) / R, 32, 0, st>>>(a);
}
