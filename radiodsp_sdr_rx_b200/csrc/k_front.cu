// k_front.cu — K0+K1+K2: input gain / IQ balance, Hilbert FIR pair, sideband sum or AM envelope,
// audio band-pass FIR.  All q15, bit-exact against oracle/rdsp_oracle.c:stage_frontend.
//
// Replaces the AudioSDRpreProcessor -> AudioSDR front half of the graph wired at
// RadioDSP_SDR_RX.ino:71-72,81-82 (library absent from the reference tree; arithmetic conventions
// are arm_fir_fast_q15 / AudioMixer4 as stated in SURVEY.md A.4).
//
// Three 129-tap FIRs per channel = 49.5 k multiply-accumulates per 128-sample block.  Measured on B200
// (tools/ubench_pipes.cu): IMAD issues at the FFMA rate (123 lanes/clk/SM; an 8 x 16 register tile of this FIR
// sustains 115), shifts / permutes / min-max at half of it, and shared memory delivers one 128-byte wavefront per
// clock per SM.  A tile that re-loads its window as int32 needs one wavefront per four IMAD warp-instructions and
// is shared-memory bound at a third of the IMAD rate (r01 profiles), so:
//   * delay lines and taps stay PACKED int16 in shared memory ([128 history | 128 current] per line);
//   * a half-warp owns a channel, a lane owns 8 consecutive outputs and walks the taps 16 at a time: one chunk is
//     3 LDS.128 of samples + 2 broadcast LDS.128 of taps (1 wavefront per 8 IMAD instructions), 40 sign-extending
//     permutes on the otherwise idle ALU pipe, and 128 IMADs from registers;
//   * lanes are 16 bytes apart, so the sample loads are conflict free without padding;
//   * the 32-bit accumulators wrap (unsigned arithmetic) exactly like the CMSIS fast FIR; SSAT(acc >> 15, 16).
#include "rdsp_common.cuh"
#include "kernels.h"

namespace {

constexpr int MAXWARPS = 14;                  // up to 28 channels per CTA (chosen per launch, see launch_front)
constexpr int TROW = 136;                     // int16 per tap row (129 padded to a multiple of 8)

__device__ __forceinline__ int32_t mix_gain(int32_t x, int32_t mult)
{
    if (mult == 65536) return x;
    long long v = ((long long)mult * (long long)x) >> 16;
    v = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
    return (int32_t)v;
}

__device__ __forceinline__ void unpack8(const int4 v, int32_t *d)
{
    d[0] = lo16((uint32_t)v.x); d[1] = hi16((uint32_t)v.x); d[2] = lo16((uint32_t)v.y); d[3] = hi16((uint32_t)v.y);
    d[4] = lo16((uint32_t)v.z); d[5] = hi16((uint32_t)v.z); d[6] = lo16((uint32_t)v.w); d[7] = hi16((uint32_t)v.w);
}
__device__ __forceinline__ int4 pack8(const int32_t *v)
{
    return make_int4((int)mk16(v[0], v[1]), (int)mk16(v[2], v[3]), (int)mk16(v[4], v[5]), (int)mk16(v[6], v[7]));
}

// buf: int16 delay line, samples [0,128) history, [128,256) current.  Outputs n = 8*l16 + j, j < 8.
__device__ __forceinline__ void fir129_x8(const int16_t *buf, const int16_t *taps, int l16, int32_t y[8])
{
    uint32_t acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0u;
    const int4 *b4 = reinterpret_cast<const int4 *>(buf);                  // 8 samples per int4
    const int4 *t4 = reinterpret_cast<const int4 *>(taps);
#pragma unroll 2
    for (int c = 0; c < 8; c++) {
        const int q = 14 + l16 - 2 * c;                    // (128 + 8*l16 - 16c - 16) / 8
        int32_t s[24], tp[16];
        unpack8(b4[q], s); unpack8(b4[q + 1], s + 8); unpack8(b4[q + 2], s + 16);
        unpack8(t4[2 * c], tp); unpack8(t4[2 * c + 1], tp + 8);
#pragma unroll
        for (int kk = 0; kk < 16; kk++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] += (uint32_t)(tp[kk] * s[16 + j - kk]);
    }
    {   // tap 128 multiplies x[n-128] = history sample 8*l16 + j
        int32_t s[8];
        unpack8(b4[l16], s);
        const int32_t t = taps[128];
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] += (uint32_t)(t * s[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) y[j] = sat16(((int32_t)acc[j]) >> 15);
}

__global__ void __launch_bounds__(MAXWARPS * 32, 2) k_front(FrontArgs a)
{
    extern __shared__ __align__(16) int16_t s_dyn16[];
    int16_t *s_taps = s_dyn16;                                             // [15][TROW]
    int16_t (*s_buf)[3][256] = reinterpret_cast<int16_t (*)[3][256]>(s_dyn16 + 15 * TROW);   // [2*warps][3][256]

    for (int i = threadIdx.x; i < 15 * TROW; i += blockDim.x) {
        const int r = i / TROW, k = i % TROW;
        s_taps[i] = k < RDSP_NTAPS ? (int16_t)a.taps[r * RDSP_TAPS_PAD + k] : (int16_t)0;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = lane >> 4, l16 = lane & 15;
    const int slot = warp * 2 + half;
    const int chq = blockIdx.x * (blockDim.x >> 4) + slot;
    const bool active = chq < a.C;
    const int ch = active ? chq : a.C - 1;                   // idle half-warps shadow the last channel, stores masked

    const RdspChanParams p = a.par[ch];
    const int td = p.demod >= RDSP_N_DEMOD ? RDSP_DEMOD_AM_ : p.demod;     // (SAM is only built in k_front_tc; the host rejects it here)
    const int16_t *tapA = s_taps + (0 + td) * TROW;
    const int16_t *tapB = s_taps + (RDSP_N_DEMOD + td) * TROW;
    const int16_t *tapM = s_taps + (2 * RDSP_N_DEMOD + p.filter) * TROW;
    int16_t *bI = s_buf[slot][0], *bQ = s_buf[slot][1], *bD = s_buf[slot][2];
    int4 *bI4 = reinterpret_cast<int4 *>(bI), *bQ4 = reinterpret_cast<int4 *>(bQ), *bD4 = reinterpret_cast<int4 *>(bD);

    // delay lines: hist[ch][3][128] int16, 16 bytes (8 samples) per lane and line
    int4 *hrow = reinterpret_cast<int4 *>(a.hist + (size_t)ch * 3 * RDSP_BLK);
    bI4[l16] = hrow[l16];
    bQ4[l16] = hrow[16 + l16];
    bD4[l16] = hrow[32 + l16];

    const bool usb = (p.demod == 1 || p.demod == 3);
    const bool am = (p.demod == 4);

    // 8 frames = 32 bytes per lane and block; the next block is fetched while this one is filtered
    const int4 *src0 = reinterpret_cast<const int4 *>(a.iq + (size_t)ch * 2 * RDSP_BLK) + 2 * l16;
    const size_t blk_stride = (size_t)a.C * 2 * RDSP_BLK / 8;         // int4 units between consecutive blocks
    int4 nv0 = ld_stream16(src0), nv1 = ld_stream16(src0 + 1);
    for (int t = 0; t < a.T; t++) {
        const size_t cb = (size_t)t * a.C + ch;                    // channel-block index
        const int4 v0 = nv0, v1 = nv1;
        if (t + 1 < a.T) { nv0 = ld_stream16(src0 + (size_t)(t + 1) * blk_stride); nv1 = ld_stream16(src0 + (size_t)(t + 1) * blk_stride + 1); }
        const uint32_t w[8] = {(uint32_t)v0.x, (uint32_t)v0.y, (uint32_t)v0.z, (uint32_t)v0.w,
                               (uint32_t)v1.x, (uint32_t)v1.y, (uint32_t)v1.z, (uint32_t)v1.w};
        int32_t xi[8], xq[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            xi[j] = mix_gain(lo16(w[j]), p.mult_i);
            xq[j] = mix_gain(hi16(w[j]), p.mult_q);
        }
        const int4 pi = pack8(xi), pq = pack8(xq);
        bI4[16 + l16] = pi;
        bQ4[16 + l16] = pq;
        __syncwarp();

        int32_t ya[8], yb[8], d[8], m[8];
        fir129_x8(bI, tapA, l16, ya);
        fir129_x8(bQ, tapB, l16, yb);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (am) {
                uint32_t e = sqrt_u32_approx((uint32_t)(ya[j] * ya[j]) + (uint32_t)(yb[j] * yb[j]));
                d[j] = (int32_t)min(e, 32767u);
            } else {
                d[j] = usb ? sat16(ya[j] - yb[j]) : sat16(ya[j] + yb[j]);
            }
        }
        const int4 pd = pack8(d);
        bD4[16 + l16] = pd;
        __syncwarp();
        fir129_x8(bD, tapM, l16, m);

        if (active) {
            if (a.out_mono) st_stream16(a.out_mono + cb * RDSP_BLK + 8 * l16, pack8(m));
            if (a.out_stereo) {
                int16_t *dst = a.out_stereo + (cb * RDSP_BLK + 8 * l16) * 2;
                st_stream16(dst, make_int4((int)mk16(m[0], m[0]), (int)mk16(m[1], m[1]), (int)mk16(m[2], m[2]), (int)mk16(m[3], m[3])));
                st_stream16(dst + 8, make_int4((int)mk16(m[4], m[4]), (int)mk16(m[5], m[5]), (int)mk16(m[6], m[6]), (int)mk16(m[7], m[7])));
            }
            if (a.dbg) {
                float4 *dp = reinterpret_cast<float4 *>(a.dbg + (cb * RDSP_BLK + 8 * l16) * 2);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float f0 = (float)m[2 * q] / 32768.0f, f1 = (float)m[2 * q + 1] / 32768.0f;
                    dp[q] = make_float4(f0, f0, f1, f1);
                }
            }
        }
        __syncwarp();
        // current block becomes history (each lane moves its own 8 samples of the three lines)
        bI4[l16] = pi;
        bQ4[l16] = pq;
        bD4[l16] = pd;
        __syncwarp();
    }
    if (active) {
        hrow[l16] = bI4[l16];
        hrow[16 + l16] = bQ4[l16];
        hrow[32 + l16] = bD4[l16];
    }
}

}  // namespace

// Pick warps per CTA (w) and resident CTAs per SM (r) so that the grid is a whole number of full waves: every
// channel costs the same, so a ragged last wave is pure loss (8192 channels on 148 SMs = 55.35 per SM).
void launch_front(const FrontArgs &a, cudaStream_t st)
{
    // per (kernel, device) set-up, see launch_front_tc
    static std::atomic<int> n_sm_dev[RDSP_MAX_DEVICES], smem_dev[RDSP_MAX_DEVICES];
    const int dev = rdsp_current_device();
    int n_sm = n_sm_dev[dev].load(std::memory_order_acquire);
    if (!n_sm) {
        int v = 0;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaFuncSetAttribute(k_front, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
        smem_dev[dev].store(v, std::memory_order_release);
        n_sm_dev[dev].store(n_sm, std::memory_order_release);
    }
    const size_t smem_max = (size_t)smem_dev[dev].load(std::memory_order_acquire);
    auto need = [](int w) { return (size_t)(15 * TROW + 2 * w * 3 * 256) * sizeof(int16_t); };
    int best_w = 4, best_r = 4;
    double best = -1.0;
    for (int w = 2; w <= MAXWARPS; w++)
        for (int r = 1; r <= 14; r++) {
            if (w * r > 28) continue;                                     // 72 registers per thread (launch bounds)
            if ((need(w) + 1024) * r > smem_max + 1024) continue;
            const long ctas = (a.C + 2 * w - 1) / (2 * w);
            const long waves = (ctas + (long)n_sm * r - 1) / ((long)n_sm * r);
            double eff = (double)a.C / ((double)waves * n_sm * r * 2 * w);
            eff *= (w * r >= 20) ? 1.0 : 0.4 + 0.6 * (w * r) / 20.0;       // too few warps cannot cover LDS / IMAD latency
            eff += 1e-4 * w * r - 1e-5 * r;                               // ties: the fuller SM, then fewer and larger CTAs
            if (eff > best) { best = eff; best_w = w; best_r = r; }
        }
    // ask for enough shared memory that exactly best_r CTAs fit on an SM
    size_t smem = need(best_w);
    const size_t floor_r = (smem_max + 1024) / (best_r + 1) + 16 - 1024;   // more than an (r+1)-th of the SM
    if (smem < floor_r && floor_r <= smem_max) smem = floor_r;
    const int cpb = 2 * best_w;
    k_front<<<(a.C + cpb - 1) / cpb, best_w * 32, smem, st>>>(a);
}
