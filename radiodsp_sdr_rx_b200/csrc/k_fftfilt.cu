// k_fftfilt.cu — K5 (+K8) + K7: FFT-256 overlap-save band-pass ("PBT") filter of the complex signal
// L + jR, optional spectral-subtraction NR, f32 -> q15.
//
// Replaces doConvolutionalProcessing(), RDSP_convolutional.h:228-353 up to the DNR call:
//   q15 -> f32 (:241-242), frame = [previous block | current block] (:256-285), forward FFT (:291),
//   product with the pre-computed mask (:301), inverse FFT (:309), keep samples 128..255 (:314-318),
//   f32 -> q15 with truncation + saturation (:346-347).
// With nr_kind == SPECTRAL the product is replaced by the spectral subtraction of the backup sketch
// (backup/RDSP_convolutional_spec.h:181-238, loop bounds restated as FFT_length): magnitudes, noise floor
// from the mean of bins 30..180, one-pole tracker, subtract / floor, then the bin is rebuilt from the new magnitude
// and the ORIGINAL phase exactly the way the sketch does it (:221-238): phi = atan2(im, re), re' = m * arm_cos_f32(phi),
// im' = m * arm_sin_f32(phi) — the CMSIS fast-math pair, a 512-entry sine table with linear interpolation whose
// 1.9e-5 absolute error is part of the reference's output, so the same table and the same interpolation run here.
//
// Mapping: one warp per channel, 8 points per lane, the whole forward FFT -> product -> inverse FFT chain
// stays in registers (fft_f32_lanes.cuh) with warp-private shared memory exchanges.  The previous input
// block is carried in registers across the blocks of one call and stored in HBM as the exact q15 values.
// Channels whose NLMS DNR follows hand their f32 L signal to k_nlms through a scratch row.
#include "rdsp_common.cuh"
#include "fft_f32_lanes.cuh"
#include "kernels.h"

namespace {

#ifndef RDSP_FFTFILT_WARPS
#define RDSP_FFTFILT_WARPS 4
#endif
constexpr int WARPS = RDSP_FFTFILT_WARPS;

// arm_sin_f32 / arm_cos_f32 (CMSIS-DSP fast math): quarter = 0 for the sine, 0.25 for the cosine.
// Every operation is a separately rounded f32 operation, in the order of the C source (no FMA contraction).
__device__ __forceinline__ float arm_trig_f32(float x, bool cosine, const float *tab)
{
    float in = __fmul_rn(x, 0.159154943092f);
    if (cosine) in = __fadd_rn(in, 0.25f);
    int n = (int)in;
    if (in < 0.0f) n--;
    in = __fsub_rn(in, (float)n);
    const float findex = __fmul_rn(512.0f, in);
    const unsigned whole = (unsigned)findex & 0xFFFFu;                  // (uint16_t)findex
    const float fract = __fsub_rn(findex, (float)whole);
    const float a = tab[whole & 0x1FFu], b = tab[(whole & 0x1FFu) + 1];
    return __fadd_rn(__fmul_rn(__fsub_rn(1.0f, fract), a), __fmul_rn(fract, b));
}

#ifndef RDSP_FFTFILT_MINB
#define RDSP_FFTFILT_MINB 5     // 96 registers (r02b A/B on one box, cfg5 step: 8 warps x 3 CTAs 444.5 us, 8 x 2 441.6, 6 x 3 437.8, 4 x 5 420.6)
#endif
#ifndef RDSP_FFT_TW_REGS
#define RDSP_FFT_TW_REGS 1      // the lane's 14 twiddles live in registers across the blocks of a call (86 -> 64 us per 8-block launch)
#endif
#ifndef RDSP_FFTFILT_WARPS
#define RDSP_FFTFILT_WARPS 4
#endif
// RING: one-block calls, where this kernel appends the audio rows it emits to the ring of the audio spectrum (FftFiltArgs::ring)
template <bool RING>
__global__ void __launch_bounds__(WARPS * 32, RDSP_FFTFILT_MINB) k_fftfilt(FftFiltArgs a)
{
    __shared__ float2 s_tw[256];
    __shared__ __align__(16) float2 s_buf[WARPS][FFT256_BUF];
    __shared__ float s_sin[513];
    pdl_release_successor();

    for (int i = threadIdx.x; i < 256; i += WARPS * 32) s_tw[i] = a.tw256[i];
    if (a.nr_stage)
        for (int i = threadIdx.x; i < 513; i += WARPS * 32) s_sin[i] = a.sin512[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lc = blockIdx.x * WARPS + warp;
    if (lc >= a.n) return;
    const int ch = a.list ? a.list[lc] : a.ch0 + lc;
    float2 *buf = s_buf[warp];

    const RdspChanParams p = a.par[ch];
    const int kind = a.nr_stage ? p.nr_kind : 0;       // 0 off, 1 LMS (follows in k_nlms), 2 spectral
    const bool spectral = (kind == 2);
    const bool to_dnr = (kind == 1);

    float2 mk[8];
    if (!spectral) {
        const float2 *mrow = a.masks + (size_t)p.mask_id * 256;
#pragma unroll
        for (int j = 0; j < 8; j++) mk[j] = mrow[lane + 32 * j];
    }

    // previous block: q15 words (L | R << 16) for samples lane + 32 j
    uint32_t pw[4];
    {
        const uint32_t *lrow = reinterpret_cast<const uint32_t *>(a.last + (size_t)ch * 2 * RDSP_BLK);
#pragma unroll
        for (int j = 0; j < 4; j++) pw[j] = lrow[lane + 32 * j];
    }
    float nfloor = a.nfloor[ch];
#if RDSP_FFT_TW_REGS
    float2 twA[7], twB[7];
    fft256_load_twiddles(lane, s_tw, twA, twB);
#define FFT256(v) fft256_warp_r(lane, v, buf, twA, twB)
#else
#define FFT256(v) fft256_warp(lane, v, buf, s_tw)
#endif

    pdl_wait_predecessor();                            // tables, mask, previous block are loaded; the rows are the predecessor's output
    for (int t = 0; t < a.T; t++) {
        const size_t cb = (size_t)t * a.C + ch;
        uint32_t cw[4];
        if (a.in_mono) {
            const uint16_t *src = reinterpret_cast<const uint16_t *>(a.in_mono + cb * RDSP_BLK);
#pragma unroll
            for (int j = 0; j < 4; j++) { const uint32_t s = src[lane + 32 * j]; cw[j] = s | (s << 16); }
        } else {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(a.in_stereo + cb * 2 * RDSP_BLK);
#pragma unroll
            for (int j = 0; j < 4; j++) cw[j] = src[lane + 32 * j];
        }
        float2 v[8];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float2 q15 = make_float2(1.0f / 32768.0f, 1.0f / 32768.0f);      // exact scaling, two lanes per instruction
            v[j] = p_mul(make_float2((float)lo16(pw[j]), (float)hi16(pw[j])), q15);
            v[4 + j] = p_mul(make_float2((float)lo16(cw[j]), (float)hi16(cw[j])), q15);
            pw[j] = cw[j];
        }

        FFT256(v);                   // v[j] = X[lane + 32 j]

        // the inverse transform is conj -> forward -> conj -> 1/N: the first conjugation rides on the product
        if (!spectral) {
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = c_mul_conj_result(v[j], mk[j]);
        } else {
            float mag[8], part = 0.0f;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                mag[j] = sqrtf(__fadd_rn(__fmul_rn(v[j].x, v[j].x), __fmul_rn(v[j].y, v[j].y)));
                const int k = lane + 32 * j;
                if (k >= 30 && k <= 180) part += mag[j];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            float th = part / 150.0f;
            th = (float)((double)th * ((double)p.nr_spec_level * 1.5));
            nfloor = __fadd_rn(nfloor, __fmul_rn(th - nfloor, 0.65f));
            nfloor = nfloor > 0.0f ? nfloor : 0.0f;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float nm = mag[j] <= nfloor ? (float)((double)mag[j] * 0.2) : __fsub_rn(mag[j], nfloor);
                const float phi = atan2f(v[j].y, v[j].x);
                // conjugated on the way (the inverse transform below is conj -> forward -> conj -> 1/N)
                v[j] = make_float2(__fmul_rn(nm, arm_trig_f32(phi, true, s_sin)), -__fmul_rn(nm, arm_trig_f32(phi, false, s_sin)));
            }
        }

        FFT256(v);

        // keep x[128 + lane + 32 h].  The lane stores its four samples where they are: every store instruction of the warp
        // covers one contiguous 64- or 128-byte run, so the round trip through shared memory that gave each lane four
        // consecutive samples (r01) bought nothing and cost 16 of the block's 144 shared-memory wavefronts.
        float2 o[4];
#pragma unroll
        for (int h = 0; h < 4; h++) o[h] = p_mul(v[4 + h], make_float2(1.0f / 256.0f, -1.0f / 256.0f));

        if (to_dnr) {
#pragma unroll
            for (int h = 0; h < 4; h++) a.out_f32_L[cb * RDSP_BLK + lane + 32 * h] = o[h].x;
        } else {
            if (a.out_mono) {
#pragma unroll
                for (int h = 0; h < 4; h++) a.out_mono[cb * RDSP_BLK + lane + 32 * h] = (int16_t)f32_to_q15(o[h].x);
            } else {
                uint32_t *dst = reinterpret_cast<uint32_t *>(a.out_stereo + cb * 2 * RDSP_BLK);
#pragma unroll
                for (int h = 0; h < 4; h++) dst[lane + 32 * h] = mk16(f32_to_q15(o[h].x), f32_to_q15(o[h].y));
            }
            if (RING) {
                // one-block calls: the audio-spectrum kernel would launch a CTA per channel only to copy this row into its ring
                // on three ticks out of four; the row is appended here instead (k_spec1024.cu, `appended`)
                int16_t *rrow = a.ring + ((size_t)ch * 8 + (size_t)((a.tick_in->tick + (unsigned long long)t) & 7ull)) * RDSP_BLK;
#pragma unroll
                for (int h = 0; h < 4; h++) rrow[lane + 32 * h] = (int16_t)f32_to_q15(o[h].x);
            }
            if (a.dbg) {
                float2 *dp = reinterpret_cast<float2 *>(a.dbg + cb * 2 * RDSP_BLK);
#pragma unroll
                for (int h = 0; h < 4; h++) dp[lane + 32 * h] = o[h];
            }
        }
    }

    {
        uint32_t *lrow = reinterpret_cast<uint32_t *>(a.last + (size_t)ch * 2 * RDSP_BLK);
#pragma unroll
        for (int j = 0; j < 4; j++) lrow[lane + 32 * j] = pw[j];
    }
    if (spectral && lane == 0) a.nfloor[ch] = nfloor;
}

}  // namespace

void launch_fftfilt(const FftFiltArgs &a, cudaStream_t st)
{
    RDSP_CARVEOUT_ONCE(k_fftfilt<false>); RDSP_CARVEOUT_ONCE(k_fftfilt<true>);
    if (a.n <= 0) return;
    if (a.ring) rdsp_launch(k_fftfilt<true>, (a.n + WARPS - 1) / WARPS, WARPS * 32, 0, st, a.pdl != 0, a);
    else rdsp_launch(k_fftfilt<false>, (a.n + WARPS - 1) / WARPS, WARPS * 32, 0, st, a.pdl != 0, a);
}
