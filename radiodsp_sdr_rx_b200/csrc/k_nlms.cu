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
//   * p[0..3] are four independent 96-tap dot products against the coefficients at the start of the group
//     (registers, G lanes per channel, xor-shuffle reductions that pipeline);
//   * R[i][j] = s_{j-i}(n+j) comes from three lag-autocorrelations.  They depend on the input only, so they are
//     computed for the whole block BEFORE the main loop, spread over the eight virtual lanes of the channel instead of
//     replicated in every lane (r02: a third of the loop's scalar instructions): virtual lane k takes samples
//     16k .. 16k+15, anchors its sums exactly on six 16-sample chunk sums (no running sum lives longer than 16
//     samples: a sum carried for long would lose all its digits when the signal drops by orders of magnitude inside
//     the window, exactly where 1/(energy + eps) amplifies every error), slides them over its samples (2 FMAs per lag
//     and sample) and leaves the six values a group needs in shared memory;
//   * what remains sequential is a scalar chain of one subtract, one multiply and one FMA per sample;
//   * the coefficient update c += sum_j g[j] x[n+j] is four independent FMAs per tap.
//
// The 96 taps of a channel are cut into EIGHT segments of 12 ("virtual lanes"), each with a circular window of the
// delayed input in registers (static indices through unrolling).  Two forms: G = 8 lanes per channel hold one segment
// each; G = 4 lanes hold two (segments g and g + 4) and add their two partial sums first — which is exactly the first
// stage (xor 4) of the 8-lane butterfly, so BOTH FORMS PERFORM THE SAME ROUNDINGS IN THE SAME ORDER and give the same
// bits.  The launcher picks the form by list size for speed only (4 lanes: 30 % fewer instructions per sample, better
// with many channels; 8 lanes: half the dependent chain per group, better with few); what a channel computes depends
// neither on the form nor on the list it was launched in.  The block sits in shared memory with a row stride that keeps
// every 16-byte access of a quarter-warp on distinct banks.
//
// Both tap-sized loops run on the packed f32x2 FMA of sm_100 (FFMA2 = two IEEE f32 FMAs in one issue slot): taps are
// held as pairs (c[i+1], c[i]) and the window twice, as even pairs (w[2q], w[2q+1]) and as odd pairs (w[2q+1], w[2q+2]),
// so that every (x[k-1], x[k]) a pair of taps meets is an aligned register pair.  Each lane of a pair is exactly the
// scalar FMA sequence of the unpacked form (even taps in one lane, odd taps in the other), so results do not change.
//
// (r02b, measured and not kept: the odd pairs assembled from the even ones with register moves instead of the shifted
// shared-memory copy — 4 wavefronts per group less, but the unrolled loop loses its register rotation: 134 moves per 16
// samples, 496 -> 601 instructions.)
//
// The per-channel state in HBM is coefficients (384 B) + previous block (512 B) + energy: the CMSIS state
// buffer (last 95 inputs), x0 and the lag sums are functions of the previous block, so they are not stored.
#include "rdsp_common.cuh"
#include "kernels.h"
#include <cstdlib>

namespace {

// warps per CTA: 2 for the scalar forms, 4 for the packed one (measured inside the cfg5 step: 1 warp 0.542 ms, 2 warps
// 0.519, 3 warps 0.549, 4 warps 0.504, 5 warps 0.551 — the kernel alone takes the same time with each; the scalar form
// of cfg3 LOSES 9 % with 4)
constexpr int NWARPS_SCALAR = 2, NWARPS_PACKED = 4;
constexpr int XS = 260;                      // floats per row: 16-byte aligned, rows 4 banks apart
constexpr int D = 4;                         // samples per group
#ifndef RDSP_NLMS_MINB8
#define RDSP_NLMS_MINB8 1
#endif
constexpr int RS = 196;                      // floats per row of lag sums: [32 groups][4] + [32 groups][2], rows 4 banks apart
constexpr float LMS_EPS = 0.000000119209289f;

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

// RING: one-block calls, where the DNR appends the audio rows it emits to the ring of the audio spectrum (NlmsArgs::ring)
template <int G, bool PACKED, bool RING>
__global__ void __launch_bounds__((PACKED ? NWARPS_PACKED : NWARPS_SCALAR) * 32, G == 4 ? 7 : RDSP_NLMS_MINB8) k_nlms(NlmsArgs a)   // 4-lane form: 7 CTAs per SM hold the 16 384 channels of cfg3 in one wave
{
    constexpr int V = 8 / G;                     // virtual lanes (tap segments) per lane: 1 or 2
    constexpr int W = RDSP_LMS_NTAPS / 8;        // taps per segment
    constexpr int S = W + D;                     // circular window of a segment (slots = segment-relative sample index mod S)
    constexpr int CPW = 32 / G;                  // channels per warp
    static_assert(W % 4 == 0 && (G == 4 || G == 8) && !(PACKED && V != 1), "forms: 8 lanes (scalar or packed), 4 lanes (scalar)");
    constexpr int NWARPS = PACKED ? NWARPS_PACKED : NWARPS_SCALAR;
    __shared__ __align__(16) float s_x[NWARPS * CPW][XS];     // [0,128) previous block / outputs, [128,256) current
    __shared__ __align__(16) float s_x1[PACKED ? NWARPS * CPW : 1][XS];   // the same samples one to the left: s_x1[i] = x[i + 1]
    __shared__ __align__(16) float s_r[NWARPS * CPW][RS];     // lag sums of the block: group q = 4k + qq -> (r1[1], r1[2], r1[3], r2[2]) at 4 (8 qq + k), (r2[3], r3[3]) at 128 + 2 (8 qq + k)

    pdl_release_successor();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane % G;                      // lane within the channel group
    const int li = (blockIdx.x * NWARPS + warp) * CPW + lane / G;
    const bool active = li < a.n_list;
    const int ch = active ? (a.list ? a.list[li] : li) : 0;
    float *xb = s_x[warp * CPW + lane / G];
    float *xs = s_x1[PACKED ? warp * CPW + lane / G : 0];     // odd window pairs load from here as aligned 16-byte quads
    float *rr = s_r[warp * CPW + lane / G];

    constexpr int HP = S / 2;                    // window pairs
    float2 cpv[V][W / 2];                        // segment v: (c[2r+1], c[2r])
    float2 Ev[V][HP] = {}, O[HP] = {};           // Ev[v][q] = (w[2q], w[2q+1]), O[q] = (w[2q+1], w[(2q+2) % S]) (packed form, one segment)
    float2 (&cp)[W / 2] = cpv[0];
    float2 (&E)[HP] = Ev[0];
    auto wv = [&](int v, int m) -> float { m = ((m % S) + S) % S; return (m & 1) ? Ev[v][m / 2].y : Ev[v][m / 2].x; };
    auto w1 = [&](int m) -> float { return wv(0, m); };
    auto w2 = [&](int m) -> float2 { m = ((m % S) + S) % S; return (m & 1) ? O[m / 2] : E[m / 2]; };      // (w[m], w[m+1])
    auto seg = [&](int v) -> int { return g + G * v; };           // the virtual lane of segment v of this lane: delays W seg .. W seg + W - 1
    float energy = 0.0f, mu = 0.0f;
    bool first = false, peak = false;
    if (active) {
        const RdspChanParams p = a.par[ch];
        mu = a.mode ? p.mu_dnr : p.mu_notch;
        peak = !a.mode && p.als_peak != 0;                               // ALS "peak": the notch stage emits the estimate
        const float *cf = a.coeff + (size_t)ch * RDSP_LMS_NTAPS;
#pragma unroll
        for (int v = 0; v < V; v++)
#pragma unroll
            for (int r = 0; r < W / 2; r++)                            // tap register i of segment v <-> delay W * seg(v) + i
                cpv[v][r] = *reinterpret_cast<const float2 *>(cf + 94 - W * seg(v) - 2 * r);
        const float4 *pv = reinterpret_cast<const float4 *>(a.prev + (size_t)ch * RDSP_BLK);
        for (int i = g; i < 32; i += G) st4(xb + 4 * i, pv[i]);
        energy = a.energy[ch];
        first = a.first[ch] != 0;
    } else {
#pragma unroll
        for (int v = 0; v < V; v++)
#pragma unroll
            for (int r = 0; r < W / 2; r++) cpv[v][r] = make_float2(0.f, 0.f);
        for (int i = g; i < 32; i += G) st4(xb + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    __syncwarp();
    // shifted copy of [from, from + 128): xs[from + i] = xb[from + i + 1], and xs[from - 1] = xb[from]
    auto shift_copy = [&](int from) {
        if constexpr (PACKED) {
            for (int i = g; i < 32; i += G) {
                const float4 v = ld4(xb + from + 4 * i);
                const float nxt = (from + 4 * i + 4 < 256) ? xb[from + 4 * i + 4] : 0.0f;
                st4(xs + from + 4 * i, make_float4(v.y, v.z, v.w, nxt));
                if (i == 0 && from > 0) xs[from - 1] = v.x;
            }
        }
    };
    shift_copy(0);                               // xs[127] is completed when the first block is staged
    __syncwarp();

    // the input row of the next block travels while the current one is processed (16-byte pieces g, g + G, ... of the row)
    constexpr int NPF = 32 / G;
    int4 nx[NPF];
    auto fetch = [&](int t) {
        if (!active || t >= a.T) return;
        const size_t rb = ((size_t)t * a.C + ch) * RDSP_BLK;
        if (a.in_f32) {
            const int4 *src = reinterpret_cast<const int4 *>(a.in_f32 + rb);
#pragma unroll
            for (int k = 0; k < NPF; k++) nx[k] = src[g + G * k];
        } else {
            const int4 *src = reinterpret_cast<const int4 *>(a.in_q15 + rb);
#pragma unroll
            for (int k = 0; k < NPF / 2; k++) nx[k] = src[g + G * k];
        }
    };
    pdl_wait_predecessor();                      // own state is loaded; from here on: the predecessor's output
    fetch(0);

    const bool st_y = g == 0 && (a.mode || peak), st_e = g == 0 && !(a.mode || peak);     // lane 0 of a channel emits: estimate or error
    for (int t = 0; t < a.T; t++) {
        const size_t cb = (size_t)t * a.C + ch;
        // ---- stage the current block into xb[128..255]
        if (active) {
            if (a.in_f32) {
#pragma unroll
                for (int k = 0; k < NPF; k++)
                    st4(xb + 128 + 4 * (g + G * k), make_float4(__int_as_float(nx[k].x), __int_as_float(nx[k].y), __int_as_float(nx[k].z), __int_as_float(nx[k].w)));
            } else {
#pragma unroll
                for (int k = 0; k < NPF / 2; k++) {
                    const int4 v = nx[k];
                    const int i = g + G * k;
                    st4(xb + 128 + 8 * i, make_float4((float)lo16(v.x) / 32768.0f, (float)hi16(v.x) / 32768.0f,
                                                      (float)lo16(v.y) / 32768.0f, (float)hi16(v.y) / 32768.0f));
                    st4(xb + 128 + 8 * i + 4, make_float4((float)lo16(v.z) / 32768.0f, (float)hi16(v.z) / 32768.0f,
                                                          (float)lo16(v.w) / 32768.0f, (float)hi16(v.w) / 32768.0f));
                }
            }
            fetch(t + 1);                        // in flight while this block is processed
        } else {
            for (int i = g; i < 32; i += G) st4(xb + 128 + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        __syncwarp();
        shift_copy(128);

        // ---- lag sums r_l(b) = x[b-l]' x[b] (96 products, l = 1..3) of every sample of the block.  Virtual lane k owns
        // samples b = 16k .. 16k+15: xnw[i] = x[16k-4+i] (entering products), xow[i] = x[16k-100+i] (leaving products).
        {
            float xnw[V][20], xow[V][20];
            auto load_chunk = [&](int v) {
                const int k = seg(v);
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    const float4 n4 = ld4(xb + 124 + 16 * k + 4 * i), o4 = ld4(xb + 28 + 16 * k + 4 * i);
                    xnw[v][4 * i] = n4.x; xnw[v][4 * i + 1] = n4.y; xnw[v][4 * i + 2] = n4.z; xnw[v][4 * i + 3] = n4.w;
                    xow[v][4 * i] = o4.x; xow[v][4 * i + 1] = o4.y; xow[v][4 * i + 2] = o4.z; xow[v][4 * i + 3] = o4.w;
                }
            };
#pragma unroll
            for (int v = 0; v < V; v++) {
                const int k = seg(v);
                load_chunk(v);
                // chunk sums B_l(c) = sum over the 16 samples m of chunk c of x[m-l] x[m]: the entering samples are chunk k,
                // the leaving ones chunk k - 6.  Slots 0..13 of a lag <-> chunks -6..7.
                float bn1 = 0.f, bn2 = 0.f, bn3 = 0.f, bo1 = 0.f, bo2 = 0.f, bo3 = 0.f;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    bn1 = fmaf(xnw[v][3 + j], xnw[v][4 + j], bn1); bn2 = fmaf(xnw[v][2 + j], xnw[v][4 + j], bn2); bn3 = fmaf(xnw[v][1 + j], xnw[v][4 + j], bn3);
                    bo1 = fmaf(xow[v][3 + j], xow[v][4 + j], bo1); bo2 = fmaf(xow[v][2 + j], xow[v][4 + j], bo2); bo3 = fmaf(xow[v][1 + j], xow[v][4 + j], bo3);
                }
                rr[k + 6] = bn1; rr[14 + k + 6] = bn2; rr[28 + k + 6] = bn3;
                if (k < 6) { rr[k] = bo1; rr[14 + k] = bo2; rr[28 + k] = bo3; }      // (k = 6, 7: chunks 0, 1 — lanes 0, 1 write them)
            }
            __syncwarp();
            // r_l(16k - 1) = B_l(k-6) + ... + B_l(k-1), summed in this order
            float sl[V][3];
#pragma unroll
            for (int v = 0; v < V; v++) {
                const int k = seg(v);
#pragma unroll
                for (int l = 0; l < 3; l++) {
                    float acc = rr[14 * l + k];
#pragma unroll
                    for (int i = 1; i < 6; i++) acc = __fadd_rn(acc, rr[14 * l + k + i]);
                    sl[v][l] = acc;
                }
            }
            __syncwarp();
            // slide: r_l(b) = r_l(b-1) + x[b-l] x[b] - x[b-l-96] x[b-96]; a group of 4 samples n..n+3 needs
            // r1(n+1), r1(n+2), r1(n+3), r2(n+2) | r2(n+3), r3(n+3)
#pragma unroll
            for (int v = 0; v < V; v++) {
                const int k = seg(v);
                if (V > 1) load_chunk(v);                  // two chunks per lane: loaded again rather than kept in 80 registers
                float s1 = sl[v][0], s2 = sl[v][1], s3 = sl[v][2];
#pragma unroll
                for (int qq = 0; qq < 4; qq++) {
                    float r1[4], r2[4], r3[4];
#pragma unroll
                    for (int jj = 0; jj < 4; jj++) {
                        const int j = 4 * qq + jj;
                        s1 = fmaf(xnw[v][3 + j], xnw[v][4 + j], s1); s1 = fmaf(-xow[v][3 + j], xow[v][4 + j], s1);
                        s2 = fmaf(xnw[v][2 + j], xnw[v][4 + j], s2); s2 = fmaf(-xow[v][2 + j], xow[v][4 + j], s2);
                        s3 = fmaf(xnw[v][1 + j], xnw[v][4 + j], s3); s3 = fmaf(-xow[v][1 + j], xow[v][4 + j], s3);
                        r1[jj] = s1; r2[jj] = s2; r3[jj] = s3;
                    }
                    // group 4k + qq sits at slot 8 qq + k: the eight lanes of a channel store 16 bytes apart (conflict free; chunk
                    // after chunk, 64 bytes apart, was a 4-way bank conflict)
                    st4(rr + 4 * (8 * qq + k), make_float4(r1[1], r1[2], r1[3], r2[2]));
                    *reinterpret_cast<float2 *>(rr + 128 + 2 * (8 * qq + k)) = make_float2(r2[3], r3[3]);
                }
            }
        }
        __syncwarp();

        // ---- segment-relative window u[m] = x[m - W*seg]; slots m mod S.  Before sample 0: m = -S .. -1 (all slots)
#pragma unroll
        for (int q = 0; q < S / 4; q++) {
#pragma unroll
            for (int vv = 0; vv < V; vv++) {
                const float4 v = ld4(xb + 128 - W * seg(vv) - S + 4 * q);   // m = -S + 4q .. -S + 4q + 3
                Ev[vv][2 * q] = make_float2(v.x, v.y); Ev[vv][2 * q + 1] = make_float2(v.z, v.w);
            }
            if (PACKED) {
                const float4 o = ld4(xs + 128 - W * g - S + 4 * q);         // m + 1: the last one (slot 0) is sample 0 already
                O[2 * q] = make_float2(o.x, o.y); O[2 * q + 1] = make_float2(o.z, o.w);
            }
        }
        const float *dref = (first && t == 0) ? xb + 128 : xb;      // desired signal: the previous block; on the very first call the same block

        for (int n0 = 0; n0 < RDSP_BLK; n0 += S) {
#pragma unroll
            for (int gq = 0; gq < S / 4; gq++) {
                const int n = n0 + 4 * gq;
                if (n < RDSP_BLK) {
                    const int sb = 4 * gq;                                  // slot of u[n] (n0 is a multiple of S)
                    // ---- loads
#pragma unroll
                    for (int v = 0; v < V; v++) {
                        const float4 un = ld4(xb + 128 + n - W * seg(v));
                        Ev[v][sb / 2] = make_float2(un.x, un.y); Ev[v][sb / 2 + 1] = make_float2(un.z, un.w);
                    }
                    // odd pairs (w[sb+1], w[sb+2]), (w[sb+3], w[sb+4]): the slot of w[sb+4] held w[sb-W], which no pair needs any more
                    if (PACKED) {
                        const float4 uo = ld4(xs + 128 + n - W * g);
                        O[sb / 2] = make_float2(uo.x, uo.y); O[sb / 2 + 1] = make_float2(uo.z, uo.w);
                    }
                    const float4 xn4 = ld4(xb + 128 + n);                   // in[n..n+3]
                    const float4 xo4 = ld4(xb + 32 + n);                    // x[n-96 .. n-93]
                    const float4 d4 = ld4(dref + n);                        // desired (always a load: a select costs four moves)
                    const int rslot = 8 * gq + n0 / S;                      // group n / 4 = 4 (n0 / 16) + gq -> slot 8 gq + n0 / 16
                    const float4 ra = ld4(rr + 4 * rslot);                  // r1(n+1), r1(n+2), r1(n+3), r2(n+2)
                    const float2 rb = *reinterpret_cast<const float2 *>(rr + 128 + 2 * rslot);   // r2(n+3), r3(n+3)
                    const float xn[4] = {xn4.x, xn4.y, xn4.z, xn4.w};
                    const float xo[4] = {xo4.x, xo4.y, xo4.z, xo4.w};
                    const float dd[4] = {d4.x, d4.y, d4.z, d4.w};

                    // ---- p[j] = c' x[n+j] with the coefficients at the start of the group: even taps and odd taps in
                    // accumulators of their own, added at the end (both forms sum in this order)
                    float p[4];
                    if (PACKED) {
                        // pairs over the samples (j, j+1): tap i is a broadcast scalar, (x[n+j-i], x[n+j+1-i]) an aligned pair
                        float2 pe01 = make_float2(0.f, 0.f), po01 = pe01, pe23 = pe01, po23 = pe01;
#pragma unroll
                        for (int i = 0; i < W; i++) {
                            const float ci = (i & 1) ? cp[i / 2].x : cp[i / 2].y;
                            const float2 cc = make_float2(ci, ci);
                            if (i & 1) { po01 = __ffma2_rn(cc, w2(sb - i), po01); po23 = __ffma2_rn(cc, w2(sb + 2 - i), po23); }
                            else { pe01 = __ffma2_rn(cc, w2(sb - i), pe01); pe23 = __ffma2_rn(cc, w2(sb + 2 - i), pe23); }
                        }
                        float2 p01 = __fadd2_rn(pe01, po01), p23 = __fadd2_rn(pe23, po23);
#pragma unroll
                        for (int o = G / 2; o > 0; o >>= 1) {
                            p01 = __fadd2_rn(p01, make_float2(__shfl_xor_sync(0xffffffffu, p01.x, o), __shfl_xor_sync(0xffffffffu, p01.y, o)));
                            p23 = __fadd2_rn(p23, make_float2(__shfl_xor_sync(0xffffffffu, p23.x, o), __shfl_xor_sync(0xffffffffu, p23.y, o)));
                        }
                        p[0] = p01.x; p[1] = p01.y; p[2] = p23.x; p[3] = p23.y;
                    } else {
                        float pv[V][4];
#pragma unroll
                        for (int v = 0; v < V; v++) {
                            float2 pa[4];                                   // .y: even taps, .x: odd taps
#pragma unroll
                            for (int j = 0; j < 4; j++) pa[j] = make_float2(0.f, 0.f);
#pragma unroll
                            for (int r = 0; r < W / 2; r++) {
#pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    pa[j].y = fmaf(cpv[v][r].y, wv(v, sb + j - 2 * r), pa[j].y);
                                    pa[j].x = fmaf(cpv[v][r].x, wv(v, sb + j - 2 * r - 1), pa[j].x);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 4; j++) pv[v][j] = pa[j].y + pa[j].x;
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) p[j] = V == 2 ? pv[0][j] + pv[V - 1][j] : pv[0][j];      // = the xor-4 stage of the 8-lane form
#pragma unroll
                        for (int o = G / 2 >= 4 ? 4 : 2; o > 0; o >>= 1) {
#pragma unroll
                            for (int j = 0; j < 4; j++) p[j] += __shfl_xor_sync(0xffffffffu, p[j], o);
                        }
                    }

                    // ---- scalars that do not depend on the error: energy, normaliser
                    float qn[4];
                    // squares, the + eps and the * mu two samples per instruction in the packed form (same roundings)
                    float xo2[4], xn2[4], en[4];
                    if (PACKED) {
                        const float2 a0 = __fmul2_rn(make_float2(xo4.x, xo4.y), make_float2(xo4.x, xo4.y));
                        const float2 a1 = __fmul2_rn(make_float2(xo4.z, xo4.w), make_float2(xo4.z, xo4.w));
                        const float2 b0 = __fmul2_rn(make_float2(xn4.x, xn4.y), make_float2(xn4.x, xn4.y));
                        const float2 b1 = __fmul2_rn(make_float2(xn4.z, xn4.w), make_float2(xn4.z, xn4.w));
                        xo2[0] = a0.x; xo2[1] = a0.y; xo2[2] = a1.x; xo2[3] = a1.y;
                        xn2[0] = b0.x; xn2[1] = b0.y; xn2[2] = b1.x; xn2[3] = b1.y;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; j++) { xo2[j] = __fmul_rn(xo[j], xo[j]); xn2[j] = __fmul_rn(xn[j], xn[j]); }
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        energy = __fsub_rn(energy, xo2[j]);
                        energy = __fadd_rn(energy, xn2[j]);
                        en[j] = energy;
                    }
                    // energy is a running difference: never divide by <= 0.  MUFU.RCP (relative error 2^-23; the
                    // reference divides, which this path never reproduced bit for bit anyway)
                    {
                        float den[4], rc[4];
                        if (PACKED) {
                            const float2 eps2 = make_float2(LMS_EPS, LMS_EPS);
                            const float2 d0 = __fadd2_rn(make_float2(en[0], en[1]), eps2), d1 = __fadd2_rn(make_float2(en[2], en[3]), eps2);
                            den[0] = d0.x; den[1] = d0.y; den[2] = d1.x; den[3] = d1.y;
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; j++) den[j] = en[j] + LMS_EPS;
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const float dj = fmaxf(den[j], LMS_EPS);
                            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc[j]) : "f"(dj));
                        }
                        if (PACKED) {
                            const float2 m2 = make_float2(mu, mu);
                            const float2 q0 = __fmul2_rn(m2, make_float2(rc[0], rc[1])), q1 = __fmul2_rn(m2, make_float2(rc[2], rc[3]));
                            qn[0] = q0.x; qn[1] = q0.y; qn[2] = q1.x; qn[3] = q1.y;
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; j++) qn[j] = mu * rc[j];
                        }
                    }

                    // ---- the sequential part: one subtract, one multiply, one FMA per sample
                    float y[4], e[4], gj[4];
                    y[0] = p[0];
                    e[0] = dd[0] - y[0]; gj[0] = e[0] * qn[0];
                    y[1] = fmaf(gj[0], ra.x, p[1]);
                    e[1] = dd[1] - y[1]; gj[1] = e[1] * qn[1];
                    y[2] = fmaf(gj[1], ra.y, fmaf(gj[0], ra.w, p[2]));
                    e[2] = dd[2] - y[2]; gj[2] = e[2] * qn[2];
                    y[3] = fmaf(gj[2], ra.z, fmaf(gj[1], rb.x, fmaf(gj[0], rb.y, p[3])));
                    e[3] = dd[3] - y[3]; gj[3] = e[3] * qn[3];

                    if (st_y) st4(xb + n, make_float4(y[0], y[1], y[2], y[3]));        // two predicated stores, no selects
                    if (st_e) st4(xb + n, make_float4(e[0], e[1], e[2], e[3]));

                    // ---- coefficient update c += sum_j g[j] x[n+j]
#pragma unroll
                    for (int v = 0; v < V; v++) {
#pragma unroll
                        for (int r = 0; r < W / 2; r++) {
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                if (PACKED) cp[r] = __ffma2_rn(make_float2(gj[j], gj[j]), w2(sb + j - 2 * r - 1), cp[r]);
                                else {
                                    cpv[v][r].y = fmaf(gj[j], wv(v, sb + j - 2 * r), cpv[v][r].y);
                                    cpv[v][r].x = fmaf(gj[j], wv(v, sb + j - 2 * r - 1), cpv[v][r].x);
                                }
                            }
                        }
                    }
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
                int2 *dst_mono = reinterpret_cast<int2 *>(a.out_mono + cb * RDSP_BLK);
                // one-block calls: the L row goes into the audio-spectrum ring from here (k_spec1024.cu, `appended`)
                int2 *dst_ring = RING ? reinterpret_cast<int2 *>(a.ring + ((size_t)ch * 8 + (size_t)((a.tick_in->tick + (unsigned long long)t) & 7ull)) * RDSP_BLK) : nullptr;
                float4 *dbg = a.dbg ? reinterpret_cast<float4 *>(a.dbg + cb * 2 * RDSP_BLK) : nullptr;
                for (int i = g; i < 32; i += G) {
                    const float4 yv = ld4(xb + 4 * i);
                    const float f0 = (float)((double)yv.x * 1.1), f1 = (float)((double)yv.y * 1.1);
                    const float f2 = (float)((double)yv.z * 1.1), f3 = (float)((double)yv.w * 1.1);
                    const int32_t q0 = f32_to_q15(f0), q1 = f32_to_q15(f1), q2 = f32_to_q15(f2), q3 = f32_to_q15(f3);
                    if (a.out_mono) dst_mono[i] = make_int2((int)mk16(q0, q1), (int)mk16(q2, q3));     // RDSP_AUDIO_MONO: L only
                    else dst[i] = make_int4((int)mk16(q0, q0), (int)mk16(q1, q1), (int)mk16(q2, q2), (int)mk16(q3, q3));
                    if (RING) dst_ring[i] = make_int2((int)mk16(q0, q1), (int)mk16(q2, q3));
                    if (dbg) {
                        dbg[2 * i] = make_float4(f0, f0, f1, f1);
                        dbg[2 * i + 1] = make_float4(f2, f2, f3, f3);
                    }
                }
            }
        }
        __syncwarp();
        for (int i = g; i < 32; i += G) {                                         // current block becomes the previous one
            st4(xb + 4 * i, ld4(xb + 128 + 4 * i));
            if (PACKED) st4(xs + 4 * i, ld4(xs + 128 + 4 * i));
        }
        // Safety net, outside the reference's arithmetic: when the running energy has lost its digits the recurrence can
        // run away to inf / NaN (it does in the reference too, and its coefficients then stay NaN for ever because
        // Init_LMS_NR never clears them).  A channel whose filter went non-finite restarts from zero coefficients.
        {
            float chk = energy;
#pragma unroll
            for (int v = 0; v < V; v++)
#pragma unroll
                for (int r = 0; r < W / 2; r++) chk += cpv[v][r].x + cpv[v][r].y;
            bool bad = !isfinite(chk);
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) bad |= (__shfl_xor_sync(0xffffffffu, (int)bad, o) != 0);
            if (bad) {
#pragma unroll
                for (int v = 0; v < V; v++)
#pragma unroll
                    for (int r = 0; r < W / 2; r++) cpv[v][r] = make_float2(0.f, 0.f);
                energy = 0.0f;
                for (int i = g; i < 32; i += G) {                                                    // like Init_LMS_NR: history cleared too
                    st4(xb + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
                    if (PACKED) st4(xs + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
                }
            }
        }
        __syncwarp();
    }

    if (active) {
        float *cf = a.coeff + (size_t)ch * RDSP_LMS_NTAPS;
#pragma unroll
        for (int v = 0; v < V; v++)
#pragma unroll
            for (int r = 0; r < W / 2; r++) *reinterpret_cast<float2 *>(cf + 94 - W * seg(v) - 2 * r) = cpv[v][r];
        float4 *pv = reinterpret_cast<float4 *>(a.prev + (size_t)ch * RDSP_BLK);
        for (int i = g; i < 32; i += G) pv[i] = ld4(xb + 4 * i);
        if (g == 0) {
            a.energy[ch] = energy;
            a.first[ch] = 0;
        }
    }
}

}  // namespace

void launch_nlms_direct(const NlmsArgs &a, cudaStream_t st);

void launch_nlms(const NlmsArgs &a, cudaStream_t st)
{
    if (a.n_list <= 0) return;
    if (a.direct) { launch_nlms_direct(a, st); return; }
    // 4 lanes per channel minimise instructions (the reductions are two shuffle stages, 30 % fewer instructions per
    // sample); 8 lanes halve the dependent chain of a group.  Measured alone (8 blocks per launch, us, G = 4 / G = 8):
    // 2048 channels 170 / 100, 6554 channels 167 / 165, 8192 channels 165 / 211, 16384 channels 324 / 336; inside the
    // cfg5 step (6554 DNR channels beside the other kernels) G = 8 is 4 % ahead, inside cfg4a (8192) G = 4 by 11 %.
    // The two forms give the same bits (the 4-lane one adds its two segment sums first = the xor-4 stage of the 8-lane
    // butterfly), so the switch is a pure speed choice and sits where the measurements cross.
    int G = a.n_list >= 7168 ? 4 : 8;
    if (const char *env = getenv("RDSP_NLMS_LANES")) G = atoi(env) == 4 ? 4 : 8;       // experiments only
    // The packed f32x2 form issues half the tap FMAs (FFMA2) for the same FMA-pipe time and a second window copy.  It
    // wins where other kernels compete for the issue slots (cfg5: 0.585 -> 0.537 ms per step although the kernel alone
    // takes the same 145 us) and loses where the NLMS has the GPU to itself (cfg3, 8192 notch channels: 270 -> 315 us),
    // so the caller says which situation this is.  Both forms perform the same roundings per lane: bit-identical output.
    // A handle with spectrum branches has other kernels competing for the issue slots while the NLMS runs.
    bool packed = a.contended != 0;
    if (const char *env = getenv("RDSP_NLMS_PACKED")) packed = env[0] == '1';          // experiments only
    const int nw = (G == 8 && packed) ? NWARPS_PACKED : NWARPS_SCALAR;
    const int cpb = nw * (32 / G);
    const int grid = (a.n_list + cpb - 1) / cpb;
    RDSP_CARVEOUT_ONCE((k_nlms<4, false, false>)); RDSP_CARVEOUT_ONCE((k_nlms<8, false, false>)); RDSP_CARVEOUT_ONCE((k_nlms<8, true, false>));
    RDSP_CARVEOUT_ONCE((k_nlms<4, false, true>)); RDSP_CARVEOUT_ONCE((k_nlms<8, false, true>)); RDSP_CARVEOUT_ONCE((k_nlms<8, true, true>));
    const bool ring = a.mode == 1 && a.ring != nullptr;                     // one-block calls: the DNR appends its rows to the audio-spectrum ring
    // (G = 4 packed measured slower than G = 4 scalar everywhere: cfg5 0.565 vs 0.530 ms for G = 8 packed, cfg4a 0.381 vs 0.333)
    if (ring) {
        if (G == 4) rdsp_launch(k_nlms<4, false, true>, grid, nw * 32, 0, st, a.pdl != 0, a);
        else if (packed) rdsp_launch(k_nlms<8, true, true>, grid, nw * 32, 0, st, a.pdl != 0, a);
        else rdsp_launch(k_nlms<8, false, true>, grid, nw * 32, 0, st, a.pdl != 0, a);
    } else {
        if (G == 4) rdsp_launch(k_nlms<4, false, false>, grid, nw * 32, 0, st, a.pdl != 0, a);
        else if (packed) rdsp_launch(k_nlms<8, true, false>, grid, nw * 32, 0, st, a.pdl != 0, a);
        else rdsp_launch(k_nlms<8, false, false>, grid, nw * 32, 0, st, a.pdl != 0, a);
    }
}
