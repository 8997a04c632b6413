// k_nlms_direct.cu — K3 / K6, DIRECT per-sample form of the 96-tap normalised LMS (kept as the cross-check and
// fallback of k_nlms.cu; selected with RDSP_NLMS_IMPL=direct).
//
// Replaces LMS_NoiseReduction() + arm_lms_norm_f32 (RDSP_noise_reduction.h:66-80; CMSIS semantics per
// SURVEY.md A.1).  The FIR input is the current block, the desired signal is the block 128 samples
// earlier (the de-correlation ring of RDSP_noise_reduction.h:71-79; on the very first call it is the
// same block).  K6 emits the estimate y (x1.1, L = R, RDSP_convolutional.h:332-336),
// K3 emits the error d - y.
//
// The recurrence is sequential in time (coefficients at sample n depend on the error at n-1), so the
// parallelism is across channels and across taps: G lanes per channel, W = 96/G taps per lane held in
// registers together with a W-deep circular window of the delayed input (static indices through
// unrolling by W).  Per sample: W FMAs (dot) + log2(G) xor-shuffles + W FMAs (update); the normaliser
// 1/(energy + eps) is computed off the critical path.  Channels per warp = 32/G.
//
// The per-channel state in HBM is coefficients (384 B) + previous block (512 B) + energy: the CMSIS
// state buffer (last 95 inputs) and x0 are a suffix of the previous block, so they are not stored twice.
#include "rdsp_common.cuh"
#include "kernels.h"

namespace {

constexpr int NWARPS = 2;
constexpr int XS = 257;

__device__ __forceinline__ void st4(float *p, float4 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; }
__device__ __forceinline__ float4 ld4(const float *p) { return make_float4(p[0], p[1], p[2], p[3]); }
constexpr float LMS_EPS = 0.000000119209289f;

template <int G>
__global__ void __launch_bounds__(NWARPS * 32) k_nlms_direct(NlmsArgs a)
{
    constexpr int W = RDSP_LMS_NTAPS / G;        // taps per lane
    constexpr int CPW = 32 / G;                  // channels per warp
    // [0,128) previous block / outputs, [128,256) current.  Row stride 257: the 32/G channels of a warp and the G
    // lanes of a channel (offsets -W*g, W a multiple of 4) then hit 32 different banks on every access.
    __shared__ float s_x[NWARPS * CPW][XS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane % G;                      // lane within the channel group
    const int slot = warp * CPW + lane / G;
    const int li = (blockIdx.x * NWARPS + warp) * CPW + lane / G;
    const bool active = li < a.n_list;
    const int ch = active ? (a.list ? a.list[li] : li) : 0;
    float *xb = s_x[slot];

    float c[W], xw[W];
    float energy = 0.0f, mu = 0.0f;
    bool first = false, peak = false;
    if (active) {
        const RdspChanParams p = a.par[ch];
        mu = a.mode ? p.mu_dnr : p.mu_notch;
        peak = !a.mode && p.als_peak != 0;
        const float *cf = a.coeff + (size_t)ch * RDSP_LMS_NTAPS;
#pragma unroll
        for (int i = 0; i < W; i++) c[i] = cf[95 - W * g - i];        // register i <-> delay W*g + i
        const float4 *pv = reinterpret_cast<const float4 *>(a.prev + (size_t)ch * RDSP_BLK);
        for (int i = g; i < 32; i += G) st4(xb + 4 * i, pv[i]);
        energy = a.energy[ch];
        first = a.first[ch] != 0;
    } else {
#pragma unroll
        for (int i = 0; i < W; i++) c[i] = 0.0f;
        for (int i = g; i < 32; i += G) st4(xb + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
    }

    for (int t = 0; t < a.T; t++) {
        const size_t cb = (size_t)t * a.C + ch;
        // ---- stage the current block into xb[128..255]
        if (active) {
            if (a.in_f32) {
                const float4 *src = reinterpret_cast<const float4 *>(a.in_f32 + cb * RDSP_BLK);
                for (int i = g; i < 32; i += G) st4(xb + 128 + 4 * i, src[i]);
            } else {
                const int4 *src = reinterpret_cast<const int4 *>(a.in_q15 + cb * RDSP_BLK);
                for (int i = g; i < 16; i += G) {
                    const int4 v = src[i];
                    st4(xb + 128 + 8 * i, make_float4((float)lo16(v.x) / 32768.0f, (float)hi16(v.x) / 32768.0f,
                                                      (float)lo16(v.y) / 32768.0f, (float)hi16(v.y) / 32768.0f));
                    st4(xb + 128 + 8 * i + 4, make_float4((float)lo16(v.z) / 32768.0f, (float)hi16(v.z) / 32768.0f,
                                                          (float)lo16(v.w) / 32768.0f, (float)hi16(v.w) / 32768.0f));
                }
            }
        } else {
            for (int i = g; i < 32; i += G) st4(xb + 128 + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        __syncwarp();

        // ---- window of this lane before sample 0: delays 1..W-1 relative to x[0 - W*g]
#pragma unroll
        for (int i = 1; i < W; i++) xw[(W - i) % W] = xb[128 - W * g - i];
        const bool same_block_ref = first && t == 0;

        for (int n0 = 0; n0 < RDSP_BLK; n0 += W) {
#pragma unroll
            for (int u = 0; u < W; u++) {
                const int n = n0 + u;
                if (n < RDSP_BLK) {
                    xw[u] = xb[128 + n - W * g];              // newest sample of this lane's window
                    const float xn = xb[128 + n];             // in
                    const float x0 = xb[32 + n];              // x[n-96], leaves the window
                    const float d = same_block_ref ? xn : xb[n];
                    energy = __fsub_rn(energy, __fmul_rn(x0, x0));
                    energy = __fadd_rn(energy, __fmul_rn(xn, xn));
                    const float inv = __frcp_rn(fmaxf(energy + LMS_EPS, LMS_EPS));   // never divide by <= 0
                    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
                    for (int i = 0; i < W; i++) {
                        if (i & 1) acc1 = fmaf(c[i], xw[(u - i + W) % W], acc1);
                        else acc0 = fmaf(c[i], xw[(u - i + W) % W], acc0);
                    }
                    float sum = acc0 + acc1;
#pragma unroll
                    for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                    const float e = d - sum;
                    const float w = (e * mu) * inv;
#pragma unroll
                    for (int i = 0; i < W; i++) c[i] = fmaf(w, xw[(u - i + W) % W], c[i]);
                    if (g == 0) xb[n] = (a.mode || peak) ? sum : e;    // slot n of the previous block is dead after d was read
                }
            }
        }
        __syncwarp();

        // ---- emit the block (outputs sit in xb[0..127])
        if (active) {
            if (a.mode == 0) {
                float4 *dst = reinterpret_cast<float4 *>(a.out_f32 + cb * RDSP_BLK);
                for (int i = g; i < 32; i += G) dst[i] = ld4(xb + 4 * i);
            } else {
                int4 *dst = reinterpret_cast<int4 *>(a.out_stereo + cb * 2 * RDSP_BLK);
                float4 *dbg = a.dbg ? reinterpret_cast<float4 *>(a.dbg + cb * 2 * RDSP_BLK) : nullptr;
                for (int i = g; i < 32; i += G) {
                    const float4 y = ld4(xb + 4 * i);
                    const float f0 = (float)((double)y.x * 1.1), f1 = (float)((double)y.y * 1.1);
                    const float f2 = (float)((double)y.z * 1.1), f3 = (float)((double)y.w * 1.1);
                    const int32_t q0 = f32_to_q15(f0), q1 = f32_to_q15(f1), q2 = f32_to_q15(f2), q3 = f32_to_q15(f3);
                    if (a.out_mono) reinterpret_cast<int2 *>(a.out_mono + cb * RDSP_BLK)[i] = make_int2((int)mk16(q0, q1), (int)mk16(q2, q3));
                    else dst[i] = make_int4((int)mk16(q0, q0), (int)mk16(q1, q1), (int)mk16(q2, q2), (int)mk16(q3, q3));
                    if (dbg) {
                        dbg[2 * i] = make_float4(f0, f0, f1, f1);
                        dbg[2 * i + 1] = make_float4(f2, f2, f3, f3);
                    }
                }
            }
        }
        __syncwarp();
        for (int i = g; i < 32; i += G) st4(xb + 4 * i, ld4(xb + 128 + 4 * i));   // current block becomes the previous one
        __syncwarp();
    }

    if (active) {
        float *cf = a.coeff + (size_t)ch * RDSP_LMS_NTAPS;
#pragma unroll
        for (int i = 0; i < W; i++) cf[95 - W * g - i] = c[i];
        float4 *pv = reinterpret_cast<float4 *>(a.prev + (size_t)ch * RDSP_BLK);
        for (int i = g; i < 32; i += G) pv[i] = ld4(xb + 4 * i);
        if (g == 0) {
            a.energy[ch] = energy;
            a.first[ch] = 0;
        }
    }
}

}  // namespace

void launch_nlms_direct(const NlmsArgs &a, cudaStream_t st)
{
    if (a.n_list <= 0) return;
    const int G = a.n_list >= 4096 ? 4 : 8;
    const int cpb = NWARPS * (32 / G);
    const int grid = (a.n_list + cpb - 1) / cpb;
    if (G == 4) k_nlms_direct<4><<<grid, NWARPS * 32, 0, st>>>(a);
    else k_nlms_direct<8><<<grid, NWARPS * 32, 0, st>>>(a);
}
