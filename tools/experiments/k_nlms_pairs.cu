// EXPERIMENT, NOT PART OF THE PRODUCT LIBRARY (not in the Makefile).  r02: k_nlms with two channels packed into the halves
// of the f32x2 instructions, G lanes per channel pair.  Correct (passed the whole GPU parity suite) and 2.5 x fewer warp
// instructions per channel than the shipped kernel, but SLOWER at the channel counts of BASELINE.json, because the launch
// then has fewer warps than the GPU has schedulers and a lone warp issues one instruction per 3.0 clk (ncu, r02_nlms_g4:
// issue active 33 %, stall_wait 1.18 + short_scoreboard 0.32 + no_instruction 0.22 per issue; FFMA2 occupies the FMA pipe
// for 2.3 clk, tools/ubench_lat.cu):
//     us per 8-block launch            shipped (r01 form)   pairs G = 4   pairs G = 8
//     cfg5 DNR   (6554 channels)            139                191           181
//     cfg5 notch (2048 channels)             64                171            99
//     cfg4a DNR  (8192 channels)            181                191           183
//     cfg3 notch (16384 channels)           272                305..389      322
//     cfg5 step                             520                627..637      577..585
// Kept as the record of the measurement (profiles/r02_nlms_pairs.md); see DESIGN.md section 4.
// k_nlms.cu — K3 (ALS auto-notch) and K6 (DNR): 96-tap normalised LMS, one 128-sample block per tick.
//
// Replaces LMS_NoiseReduction() + arm_lms_norm_f32 (RDSP_noise_reduction.h:66-80; CMSIS semantics per
// SURVEY.md A.1).  The FIR input is the current block, the desired signal is the block 128 samples
// earlier (the de-correlation ring of RDSP_noise_reduction.h:71-79; on the very first call it is the
// same block).  K6 emits the estimate y (x1.1, L = R, RDSP_convolutional.h:332-336),
// K3 emits the error d - y.
//
// The textbook recurrence (per sample: 96-tap dot -> error -> 96-tap update) is one long dependent chain per
// channel, and with only thousands of channels a B200 cannot hide it.  The kernel therefore evaluates the SAME
// recurrence four samples at a time with the tap-sized work taken out of the chain (exact algebra, no
// approximation; only the f32 summation order changes):
//
//     c[n+j] = c[n] + sum_{i<j} g[i] x[n+i]          (g = mu e / (energy + eps), x[m] = the 96-sample window at m)
//     y[n+j] = c[n+j]' x[n+j] = p[j] + sum_{i<j} g[i] R[i][j],   p[j] = c[n]' x[n+j],   R[i][j] = x[n+i]' x[n+j]
//
//   * p[0..3] are four independent 96-tap dot products against the coefficients at the start of the group;
//   * R[i][j] = s_{j-i}(n+j) comes from three lag-autocorrelations: anchored exactly on the register window every
//     fourth group, then slid over the samples (2 FMAs per lag and sample);
//   * what remains sequential is a scalar chain of one subtract, one multiply and one FMA per sample;
//   * the coefficient update c += sum_j g[j] x[n+j] is four independent FMAs per tap.
//
// Mapping (r02): TWO CHANNELS PER THREAD, packed into the two halves of the f32x2 instructions of sm_100
// (FFMA2 / FADD2 / FMUL2: two independent IEEE f32 operations per issue slot).  Every value of the algorithm — taps,
// window, dot products, energies, lag sums, the sequential chain — is a float2 (channel A, channel B), so the scalar
// part of the recurrence, which used to cost as many issue slots as the taps, is shared by two channels as well, and a
// window register is an aligned operand for every tap without a second, shifted copy.  G = 4 lanes share a channel
// pair (24 taps + a 32-slot circular window per lane, static indices through unrolling), so a warp carries 16 channels
// where the r01 kernel carried 4: 2.5 x fewer warp instructions per channel.  The kernel is latency bound by design
// at the channel counts of BASELINE.json (cfg5: 410 + 128 warps for 592 schedulers): one warp per CTA spreads them over
// the SMs, and eight independent accumulators per dot product give a lone warp the instruction-level parallelism to
// issue nearly every cycle.  ONE form serves every channel count, and the two halves of a packed instruction never
// mix, so a channel's bits depend neither on its partner nor on the list it was launched in.
//
// The block of a channel pair sits in shared memory as float2 rows ([0,128) previous block / outputs, [128,256) current);
// input rows arrive by cp.async one block ahead.  State in HBM per channel: coefficients (384 B) + previous block
// (512 B) + energy: the CMSIS state buffer (last 95 inputs), x0 and the lag sums are functions of the previous block.
#include "rdsp_common.cuh"
#include "kernels.h"
#include <cstdlib>

namespace {

constexpr int D = 4;                         // samples per group
constexpr int ROW = 258;                     // float2 per pair row: the lanes of a quarter-warp (8 pairs, same g) hit 8 distinct 16-byte bank groups
constexpr int ANCHOR = 4;                    // groups between exact re-anchorings of the lag sums
constexpr float LMS_EPS = 0.000000119209289f;
// G lanes per channel pair: W taps per lane, a circular window of S slots (slot = lane-relative sample index mod S; W + D are
// needed, a power of two divides the block, so the unrolled body has no conditional tail), PPW channel pairs per warp
template <int G> struct Geo {
    static constexpr int W = RDSP_LMS_NTAPS / G, S = G == 4 ? 32 : 16, PPW = 32 / G;
    static constexpr int SMEM_X = PPW * ROW * (int)sizeof(float2);              // per warp: the pair rows ...
    static constexpr int SMEM_RAW = PPW * 2 * RDSP_BLK * (int)sizeof(float);    // ... and the staging rows
    static_assert(W % 4 == 0 && S % 4 == 0 && S >= W + D && RDSP_BLK % S == 0, "window geometry");
};

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 shfl2(float2 v, int o)
{
    return make_float2(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o));
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int G, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) k_nlms(NlmsArgs a)
{
    constexpr int W = Geo<G>::W, S = Geo<G>::S, PPW = Geo<G>::PPW, SMEM_X = Geo<G>::SMEM_X;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    float2 (*s_x)[ROW] = reinterpret_cast<float2 (*)[ROW]>(s_dyn);                                     // [NWARPS * PPW][ROW]: (channel A, channel B) per sample
    float (*s_raw)[2][RDSP_BLK] = reinterpret_cast<float (*)[2][RDSP_BLK]>(s_dyn + SMEM_X * NWARPS);   // input rows of the next block as they arrive

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / PPW, pr = lane % PPW;                     // lane of the pair / pair of the warp
    const int li = (blockIdx.x * NWARPS + warp) * PPW + pr;        // pair index in the launch list
    const bool actA = 2 * li < a.n_list, actB = 2 * li + 1 < a.n_list;
    const int chA = actA ? (a.list ? a.list[2 * li] : 2 * li) : 0;
    const int chB = actB ? (a.list ? a.list[2 * li + 1] : 2 * li + 1) : chA;
    float *xb = reinterpret_cast<float *>(s_x[warp * PPW + pr]);   // float index 2 k = sample k of channel A, 2 k + 1 = channel B
    float *rawA = s_raw[warp * PPW + pr][0], *rawB = s_raw[warp * PPW + pr][1];

    float2 cp[W];                                // tap register i <-> delay W g + i
    float2 E[S];                                 // window: E[m mod S] = x[m - W g]
    auto w = [&](int m) -> float2 { return E[((m % S) + S) % S]; };
    float2 energy = f2(0.f, 0.f), mu = f2(0.f, 0.f);
    bool firstA = false, firstB = false, peakA = false, peakB = false;
    {
        // coefficient rows: delays W g .. W g + W - 1 are the coefficients 95 - W g down to 96 - W - W g (CMSIS order: index 0 = oldest)
        float ca[W], cb[W];
#pragma unroll
        for (int q = 0; q < W / 4; q++) {
            const float4 va = actA ? ld4(a.coeff + (size_t)chA * RDSP_LMS_NTAPS + (RDSP_LMS_NTAPS - W) - W * g + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 vb = actB ? ld4(a.coeff + (size_t)chB * RDSP_LMS_NTAPS + (RDSP_LMS_NTAPS - W) - W * g + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
            ca[4 * q] = va.x; ca[4 * q + 1] = va.y; ca[4 * q + 2] = va.z; ca[4 * q + 3] = va.w;
            cb[4 * q] = vb.x; cb[4 * q + 1] = vb.y; cb[4 * q + 2] = vb.z; cb[4 * q + 3] = vb.w;
        }
#pragma unroll
        for (int i = 0; i < W; i++) cp[i] = f2(ca[W - 1 - i], cb[W - 1 - i]);
        // previous block -> xb[0..127]
        for (int k = g; k < 32; k += G) {
            const float4 va = actA ? ld4(a.prev + (size_t)chA * RDSP_BLK + 4 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 vb = actB ? ld4(a.prev + (size_t)chB * RDSP_BLK + 4 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
            st4(xb + 8 * k, make_float4(va.x, vb.x, va.y, vb.y));
            st4(xb + 8 * k + 4, make_float4(va.z, vb.z, va.w, vb.w));
        }
        if (actA) {
            const RdspChanParams p = a.par[chA];
            mu.x = a.mode ? p.mu_dnr : p.mu_notch;
            peakA = !a.mode && p.als_peak != 0;                                // ALS "peak": the notch stage emits the estimate
            energy.x = a.energy[chA];
            firstA = a.first[chA] != 0;
        }
        if (actB) {
            const RdspChanParams p = a.par[chB];
            mu.y = a.mode ? p.mu_dnr : p.mu_notch;
            peakB = !a.mode && p.als_peak != 0;
            energy.y = a.energy[chB];
            firstB = a.first[chB] != 0;
        }
    }
    const bool estA = a.mode || peakA, estB = a.mode || peakB;     // emit the estimate (else the error)

    // the input rows of block t travel into s_raw while block t - 1 is processed (16-byte pieces g, g + G, ...)
    auto fetch = [&](int t) {
        if (t >= a.T) return;
        const size_t ra = ((size_t)t * a.C + chA) * RDSP_BLK, rb = ((size_t)t * a.C + chB) * RDSP_BLK;
        if (a.in_f32) {
            for (int k = g; k < 32; k += G) {
                if (actA) cp_async16(rawA + 4 * k, a.in_f32 + ra + 4 * k);
                if (actB) cp_async16(rawB + 4 * k, a.in_f32 + rb + 4 * k);
            }
        } else {
            for (int k = g; k < 16; k += G) {
                if (actA) cp_async16(rawA + 4 * k, a.in_q15 + ra + 8 * k);
                if (actB) cp_async16(rawB + 4 * k, a.in_q15 + rb + 8 * k);
            }
        }
    };
    fetch(0);
    __syncwarp();

    for (int t = 0; t < a.T; t++) {
        const size_t cbA = (size_t)t * a.C + chA, cbB = (size_t)t * a.C + chB;
        // ---- stage the current block into xb[128..255]
        cp_async_wait_all();
        __syncwarp();
        if (a.in_f32) {
            for (int k = g; k < 32; k += G) {
                const float4 va = actA ? ld4(rawA + 4 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 vb = actB ? ld4(rawB + 4 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
                st4(xb + 256 + 8 * k, make_float4(va.x, vb.x, va.y, vb.y));
                st4(xb + 256 + 8 * k + 4, make_float4(va.z, vb.z, va.w, vb.w));
            }
        } else {
            for (int k = g; k < 16; k += G) {
                const int4 va = actA ? *reinterpret_cast<const int4 *>(rawA + 4 * k) : make_int4(0, 0, 0, 0);
                const int4 vb = actB ? *reinterpret_cast<const int4 *>(rawB + 4 * k) : make_int4(0, 0, 0, 0);
                const float q = 1.0f / 32768.0f;                                   // exact scaling (arm_q15_to_float)
                st4(xb + 256 + 16 * k, make_float4((float)lo16(va.x) * q, (float)lo16(vb.x) * q, (float)hi16(va.x) * q, (float)hi16(vb.x) * q));
                st4(xb + 256 + 16 * k + 4, make_float4((float)lo16(va.y) * q, (float)lo16(vb.y) * q, (float)hi16(va.y) * q, (float)hi16(vb.y) * q));
                st4(xb + 256 + 16 * k + 8, make_float4((float)lo16(va.z) * q, (float)lo16(vb.z) * q, (float)hi16(va.z) * q, (float)hi16(vb.z) * q));
                st4(xb + 256 + 16 * k + 12, make_float4((float)lo16(va.w) * q, (float)lo16(vb.w) * q, (float)hi16(va.w) * q, (float)hi16(vb.w) * q));
            }
        }
        __syncwarp();
        fetch(t + 1);                            // in flight while this block is processed

        // ---- lane-relative window u[m] = x[m - W g]; slots m mod S.  Before sample 0: m = -S .. -1 (all slots)
#pragma unroll
        for (int q = 0; q < S / 2; q++) {
            const float4 v = ld4(xb + 2 * (128 - W * g - S + 2 * q));           // m = -S + 2q, -S + 2q + 1
            E[2 * q] = f2(v.x, v.y); E[2 * q + 1] = f2(v.z, v.w);
        }
        float2 xnp[4], xop[4];                   // x[-4..-1], x[-100..-97]
        {
            const float4 a0 = ld4(xb + 2 * 124), a1 = ld4(xb + 2 * 126), b0 = ld4(xb + 2 * 28), b1 = ld4(xb + 2 * 30);
            xnp[0] = f2(a0.x, a0.y); xnp[1] = f2(a0.z, a0.w); xnp[2] = f2(a1.x, a1.y); xnp[3] = f2(a1.z, a1.w);
            xop[0] = f2(b0.x, b0.y); xop[1] = f2(b0.z, b0.w); xop[2] = f2(b1.x, b1.y); xop[3] = f2(b1.z, b1.w);
        }
        const bool sameA = firstA && t == 0, sameB = firstB && t == 0;       // desired = the same block on the very first call

        float2 s1 = f2(0.f, 0.f), s2 = s1, s3 = s1;
        for (int n0 = 0; n0 < RDSP_BLK; n0 += S) {
#pragma unroll
            for (int gq = 0; gq < S / 4; gq++) {
                const int n = n0 + 4 * gq;
                {
                    const int sb = 4 * gq;                                  // slot of u[n] (n0 is a multiple of S)
                    // ---- lag sums s_l(n-1) = x[n-1-l]' x[n-1], l = 1..3, anchored EXACTLY on the window every ANCHOR
                    // groups (the window still holds m = n-W-4 .. n-1) and slid over the samples in between.  A running
                    // sum carried for long would lose all its digits when the signal drops by orders of magnitude
                    // inside the window, exactly where 1/(energy + eps) amplifies every error.
                    if (gq % ANCHOR == 0) {
                        float2 t1 = f2(0.f, 0.f), t2 = t1, t3 = t1, u1 = t1, u2 = t1, u3 = t1;
#pragma unroll
                        for (int i = 0; i < W; i += 2) {
                            const float2 uk = w(sb - 1 - i), uj = w(sb - 2 - i);
                            t1 = __ffma2_rn(uj, uk, t1);
                            t2 = __ffma2_rn(w(sb - 3 - i), uk, t2);
                            t3 = __ffma2_rn(w(sb - 4 - i), uk, t3);
                            u1 = __ffma2_rn(w(sb - 3 - i), uj, u1);
                            u2 = __ffma2_rn(w(sb - 4 - i), uj, u2);
                            u3 = __ffma2_rn(w(sb - 5 - i), uj, u3);
                        }
                        s1 = __fadd2_rn(t1, u1); s2 = __fadd2_rn(t2, u2); s3 = __fadd2_rn(t3, u3);
#pragma unroll
                        for (int o = PPW; o < 32; o <<= 1) {
                            s1 = __fadd2_rn(s1, shfl2(s1, o));
                            s2 = __fadd2_rn(s2, shfl2(s2, o));
                            s3 = __fadd2_rn(s3, shfl2(s3, o));
                        }
                    }
                    // ---- loads: the four new window samples of this lane, the newest / oldest / desired samples of the group
                    float2 xn[4], xo[4], dd[4];
                    {
                        const float4 u0 = ld4(xb + 2 * (128 + n - W * g)), u1 = ld4(xb + 2 * (130 + n - W * g));
                        E[sb] = f2(u0.x, u0.y); E[sb + 1] = f2(u0.z, u0.w); E[sb + 2] = f2(u1.x, u1.y); E[sb + 3] = f2(u1.z, u1.w);
                        const float4 n0v = ld4(xb + 2 * (128 + n)), n1v = ld4(xb + 2 * (130 + n));      // in[n .. n+3]
                        const float4 o0v = ld4(xb + 2 * (32 + n)), o1v = ld4(xb + 2 * (34 + n));        // x[n-96 .. n-93]
                        const float4 d0v = ld4(xb + 2 * n), d1v = ld4(xb + 2 * (n + 2));                // the block before
                        xn[0] = f2(n0v.x, n0v.y); xn[1] = f2(n0v.z, n0v.w); xn[2] = f2(n1v.x, n1v.y); xn[3] = f2(n1v.z, n1v.w);
                        xo[0] = f2(o0v.x, o0v.y); xo[1] = f2(o0v.z, o0v.w); xo[2] = f2(o1v.x, o1v.y); xo[3] = f2(o1v.z, o1v.w);
                        dd[0] = f2(d0v.x, d0v.y); dd[1] = f2(d0v.z, d0v.w); dd[2] = f2(d1v.x, d1v.y); dd[3] = f2(d1v.z, d1v.w);
#pragma unroll
                        for (int j = 0; j < 4; j++) dd[j] = f2(sameA ? xn[j].x : dd[j].x, sameB ? xn[j].y : dd[j].y);
                    }

                    // ---- p[j] = c' x[n+j] with the coefficients at the start of the group: even and odd taps in accumulators
                    // of their own (eight independent chains keep a lone warp issuing), added at the end
                    float2 p[4];
                    {
                        float2 pe[4], po[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) { pe[j] = f2(0.f, 0.f); po[j] = f2(0.f, 0.f); }
#pragma unroll
                        for (int i = 0; i < W; i += 2) {
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                pe[j] = __ffma2_rn(cp[i], w(sb + j - i), pe[j]);
                                po[j] = __ffma2_rn(cp[i + 1], w(sb + j - i - 1), po[j]);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) p[j] = __fadd2_rn(pe[j], po[j]);
#pragma unroll
                        for (int o = PPW; o < 32; o <<= 1) {
#pragma unroll
                            for (int j = 0; j < 4; j++) p[j] = __fadd2_rn(p[j], shfl2(p[j], o));
                        }
                    }

                    // ---- scalars that do not depend on the error: energy, normaliser, lag sums
                    float2 qn[4], r1[4], r2[4], r3[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        energy = __fadd2_rn(energy, neg2(__fmul2_rn(xo[j], xo[j])));
                        energy = __fadd2_rn(energy, __fmul2_rn(xn[j], xn[j]));
                        // energy is a running difference: never divide by <= 0.  MUFU.RCP (relative error 2^-23; the
                        // reference divides, which no other f32 evaluation order reproduces bit for bit anyway)
                        const float2 den = __fadd2_rn(energy, f2(LMS_EPS, LMS_EPS));
                        qn[j] = __fmul2_rn(mu, f2(rcp_approx(fmaxf(den.x, LMS_EPS)), rcp_approx(fmaxf(den.y, LMS_EPS))));
                    }
                    {
                        // recent / old samples around the group: index 4 + j <-> sample n + j
                        const float2 xr[8] = {xnp[0], xnp[1], xnp[2], xnp[3], xn[0], xn[1], xn[2], xn[3]};   // x[n-4 .. n+3]
                        const float2 xq[8] = {xop[0], xop[1], xop[2], xop[3], xo[0], xo[1], xo[2], xo[3]};   // x[n-100 .. n-93]
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            // s_l(b) = s_l(b-1) + x[b-l] x[b] - x[b-l-96] x[b-96],  b = n + j
                            const float2 no = neg2(xo[j]);
                            s1 = __ffma2_rn(xr[4 + j - 1], xn[j], s1); s1 = __ffma2_rn(xq[4 + j - 1], no, s1);
                            s2 = __ffma2_rn(xr[4 + j - 2], xn[j], s2); s2 = __ffma2_rn(xq[4 + j - 2], no, s2);
                            s3 = __ffma2_rn(xr[4 + j - 3], xn[j], s3); s3 = __ffma2_rn(xq[4 + j - 3], no, s3);
                            r1[j] = s1; r2[j] = s2; r3[j] = s3;
                        }
                    }

                    // ---- the sequential part: one subtract, one multiply, one FMA per sample (two channels per instruction)
                    float2 y[4], e[4], gj[4];
                    y[0] = p[0];
                    e[0] = __fadd2_rn(dd[0], neg2(y[0])); gj[0] = __fmul2_rn(e[0], qn[0]);
                    y[1] = __ffma2_rn(gj[0], r1[1], p[1]);
                    e[1] = __fadd2_rn(dd[1], neg2(y[1])); gj[1] = __fmul2_rn(e[1], qn[1]);
                    y[2] = __ffma2_rn(gj[1], r1[2], __ffma2_rn(gj[0], r2[2], p[2]));
                    e[2] = __fadd2_rn(dd[2], neg2(y[2])); gj[2] = __fmul2_rn(e[2], qn[2]);
                    y[3] = __ffma2_rn(gj[2], r1[3], __ffma2_rn(gj[1], r2[3], __ffma2_rn(gj[0], r3[3], p[3])));
                    e[3] = __fadd2_rn(dd[3], neg2(y[3])); gj[3] = __fmul2_rn(e[3], qn[3]);

                    // slots n .. n+3 of the previous block are dead once dd was read.  Every lane of the pair holds the same bits (the
                    // xor-butterfly sums commute) and stores them: no branch, so the unrolled groups stay ONE basic block and the
                    // scheduler can run the loads / lag sums / energies of the next group under the chain of this one
                    st4(xb + 2 * n, make_float4(estA ? y[0].x : e[0].x, estB ? y[0].y : e[0].y, estA ? y[1].x : e[1].x, estB ? y[1].y : e[1].y));
                    st4(xb + 2 * n + 4, make_float4(estA ? y[2].x : e[2].x, estB ? y[2].y : e[2].y, estA ? y[3].x : e[3].x, estB ? y[3].y : e[3].y));

                    // ---- coefficient update c += sum_j g[j] x[n+j]
#pragma unroll
                    for (int i = 0; i < W; i++) {
#pragma unroll
                        for (int j = 0; j < 4; j++) cp[i] = __ffma2_rn(gj[j], w(sb + j - i), cp[i]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) { xnp[j] = xn[j]; xop[j] = xo[j]; }
                }
            }
        }
        __syncwarp();

        // ---- emit the block (outputs sit in xb[0..127]); lane g of a pair writes pieces g, g + G, ... of both rows
        if (a.mode == 0) {
            for (int k = g; k < 32; k += G) {
                const float4 v0 = ld4(xb + 8 * k), v1 = ld4(xb + 8 * k + 4);
                if (actA) st4(a.out_f32 + cbA * RDSP_BLK + 4 * k, make_float4(v0.x, v0.z, v1.x, v1.z));
                if (actB) st4(a.out_f32 + cbB * RDSP_BLK + 4 * k, make_float4(v0.y, v0.w, v1.y, v1.w));
            }
        } else {
            for (int k = g; k < 32; k += G) {
                const float4 v0 = ld4(xb + 8 * k), v1 = ld4(xb + 8 * k + 4);
                const float ya[4] = {v0.x, v0.z, v1.x, v1.z}, yb[4] = {v0.y, v0.w, v1.y, v1.w};
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (!(h ? actB : actA)) continue;
                    const float *yv = h ? yb : ya;
                    const size_t cb = h ? cbB : cbA;
                    float f[4];
                    int32_t q[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) { f[j] = (float)((double)yv[j] * 1.1); q[j] = f32_to_q15(f[j]); }   // RDSP_convolutional.h:334: double multiply
                    if (a.out_mono)
                        *reinterpret_cast<int2 *>(a.out_mono + cb * RDSP_BLK + 4 * k) = make_int2((int)mk16(q[0], q[1]), (int)mk16(q[2], q[3]));
                    else
                        *reinterpret_cast<int4 *>(a.out_stereo + cb * 2 * RDSP_BLK + 8 * k) =
                            make_int4((int)mk16(q[0], q[0]), (int)mk16(q[1], q[1]), (int)mk16(q[2], q[2]), (int)mk16(q[3], q[3]));
                    if (a.dbg) {
                        float4 *dbg = reinterpret_cast<float4 *>(a.dbg + cb * 2 * RDSP_BLK + 8 * k);
                        dbg[0] = make_float4(f[0], f[0], f[1], f[1]);
                        dbg[1] = make_float4(f[2], f[2], f[3], f[3]);
                    }
                }
            }
        }
        __syncwarp();
        for (int k = g; k < 64; k += G) st4(xb + 4 * k, ld4(xb + 256 + 4 * k));       // current block becomes the previous one
        // Safety net, outside the reference's arithmetic: when the running energy has lost its digits the recurrence can
        // run away to inf / NaN (it does in the reference too, and its coefficients then stay NaN for ever because
        // Init_LMS_NR never clears them).  A channel whose filter went non-finite restarts from zero coefficients.
        {
            float2 chk = energy;
#pragma unroll
            for (int i = 0; i < W; i++) chk = __fadd2_rn(chk, cp[i]);
            bool badA = !isfinite(chk.x), badB = !isfinite(chk.y);
#pragma unroll
            for (int o = PPW; o < 32; o <<= 1) {
                badA |= (__shfl_xor_sync(0xffffffffu, (int)badA, o) != 0);
                badB |= (__shfl_xor_sync(0xffffffffu, (int)badB, o) != 0);
            }
            __syncwarp();
            if (badA || badB) {
#pragma unroll
                for (int i = 0; i < W; i++) cp[i] = f2(badA ? 0.f : cp[i].x, badB ? 0.f : cp[i].y);
                energy = f2(badA ? 0.f : energy.x, badB ? 0.f : energy.y);
                for (int k = g; k < 128; k += G) {                                // like Init_LMS_NR: history cleared too
                    if (badA) xb[2 * k] = 0.f;
                    if (badB) xb[2 * k + 1] = 0.f;
                }
            }
        }
        __syncwarp();
    }

    {
        float ca[W], cb[W];
#pragma unroll
        for (int i = 0; i < W; i++) { ca[W - 1 - i] = cp[i].x; cb[W - 1 - i] = cp[i].y; }
#pragma unroll
        for (int q = 0; q < W / 4; q++) {
            if (actA) st4(a.coeff + (size_t)chA * RDSP_LMS_NTAPS + (RDSP_LMS_NTAPS - W) - W * g + 4 * q, make_float4(ca[4 * q], ca[4 * q + 1], ca[4 * q + 2], ca[4 * q + 3]));
            if (actB) st4(a.coeff + (size_t)chB * RDSP_LMS_NTAPS + (RDSP_LMS_NTAPS - W) - W * g + 4 * q, make_float4(cb[4 * q], cb[4 * q + 1], cb[4 * q + 2], cb[4 * q + 3]));
        }
        for (int k = g; k < 32; k += G) {
            const float4 v0 = ld4(xb + 8 * k), v1 = ld4(xb + 8 * k + 4);
            if (actA) st4(a.prev + (size_t)chA * RDSP_BLK + 4 * k, make_float4(v0.x, v0.z, v1.x, v1.z));
            if (actB) st4(a.prev + (size_t)chB * RDSP_BLK + 4 * k, make_float4(v0.y, v0.w, v1.y, v1.w));
        }
        if (g == 0) {
            if (actA) { a.energy[chA] = energy.x; a.first[chA] = 0; }
            if (actB) { a.energy[chB] = energy.y; a.first[chB] = 0; }
        }
    }
}

}  // namespace

void launch_nlms_direct(const NlmsArgs &a, cudaStream_t st);

template <int G, int NW>
static void launch_form(const NlmsArgs &a, int grid, cudaStream_t st)
{
    constexpr int smem = NW * (Geo<G>::SMEM_X + Geo<G>::SMEM_RAW);
    RDSP_CARVEOUT_ONCE((k_nlms<G, NW>));
    static std::atomic<bool> done[RDSP_MAX_DEVICES];                  // > 48 KB of dynamic shared memory: opt in per device
    const int dev = rdsp_current_device();
    if (smem > 48 * 1024 && !done[dev].load(std::memory_order_acquire)) {
        cudaFuncSetAttribute(k_nlms<G, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        done[dev].store(true, std::memory_order_release);
    }
    k_nlms<G, NW><<<grid, NW * 32, smem, st>>>(a);
}

void launch_nlms(const NlmsArgs &a, cudaStream_t st)
{
    if (a.n_list <= 0) return;
    if (a.direct) { launch_nlms_direct(a, st); return; }
    int G = 4, nw = 1;
    if (const char *env = getenv("RDSP_NLMS_G")) G = atoi(env) == 8 ? 8 : 4;           // experiments only
    if (const char *env = getenv("RDSP_NLMS_WARPS")) nw = atoi(env) == 2 ? 2 : 1;      // experiments only
    const int cpw = 2 * (32 / G);                                                     // channels per warp
    const int warps = (a.n_list + cpw - 1) / cpw;
    const int grid = (warps + nw - 1) / nw;
    if (G == 4) { if (nw == 2) launch_form<4, 2>(a, grid, st); else launch_form<4, 1>(a, grid, st); }
    else        { if (nw == 2) launch_form<8, 2>(a, grid, st); else launch_form<8, 1>(a, grid, st); }
}
