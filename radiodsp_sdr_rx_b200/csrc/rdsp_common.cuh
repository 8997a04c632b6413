// rdsp_common.cuh — shared device helpers and the HBM data layout of the batched receive chain.
//
// Layout conventions (all per handle, channel-major, see DESIGN.md "Data layout in HBM"):
//   iq_in      int16 [n_blocks][C][128][2]   (I,Q interleaved = I2S frame order)
//   audio_out  int16 [n_blocks][C][128][2]   (L,R interleaved)
//   mono mids  int16 [n_blocks][C][128]
//   f32 mids   float [n_blocks][C][128]
//   per-channel state: one contiguous row per channel per stage (coalesced 128-bit accesses).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define RDSP_BLK        128
#define RDSP_NTAPS      129
#define RDSP_TAPS_PAD   132          // tap rows padded to a multiple of 4 words (LDS.128)
#define RDSP_LMS_NTAPS  96
#define RDSP_N_DEMOD    5
#define RDSP_N_FILTER   5
#define RDSP_DEMOD_AM_  4            // AM envelope; RDSP_DEMOD_SAM_ shares its tap rows
#define RDSP_DEMOD_SAM_ 5

// SAM carrier loop (shim-defined, the same constants in oracle/rdsp_oracle.c): natural frequency 100 Hz, damping 0.707
// at 44.1 kHz; pull-in limited to +-1 kHz; carrier-level tracker 100 ms
#define RDSP_SAM_K1   0.020146f
#define RDSP_SAM_K2   2.02995e-4f
#define RDSP_SAM_WMAX 0.142476f
#define RDSP_SAM_ADC  2.2673e-4f
#define RDSP_SAM_PI   3.14159265358979f

// per-channel parameters as the kernels see them (written by rdsp_gpu_set_mode)
struct __align__(16) RdspChanParams {
    int32_t mult_i;        // K0 AudioMixer4-style gain, 65536 = unity
    int32_t mult_q;
    float   out_gain;      // SDR.setOutputGain
    float   mu_notch;      // K3 NLMS step size
    float   mu_dnr;        // K6 NLMS step size
    float   agc_alpha_d;   // K4 decay coefficient of the selected AGC mode
    float   nr_spec_level; // K8 iNRLevel
    int32_t mask_id;       // K5 row of the mask table
    uint32_t nb_mult_q8;   // noise blanker threshold as a Q8 factor on the running IQ magnitude (0: blanker off)
    uint8_t demod;         // RDSP_DEMOD_*
    uint8_t filter;        // RDSP_FILTER_*
    uint8_t agc_mode;      // RDSP_AGC_*
    uint8_t notch_on;
    uint8_t nr_kind;       // RDSP_NR_* (0 when level == 0)
    uint8_t als_peak;      // K3 emits the NLMS estimate instead of the error
    uint8_t pad[6];
};
static_assert(sizeof(RdspChanParams) == 48, "RdspChanParams layout");

// Tick bookkeeping that is uniform over channels, kept ON THE DEVICE so that a process call needs no per-call kernel
// argument that changes from call to call (a captured CUDA graph replays as is).  Two copies: a call reads copy
// `parity` and its kernels write the advanced values into copy `parity ^ 1`, which nothing reads during that call.
struct RdspTick {
    unsigned long long tick;   // index of the first block of the call (AudioAnalyzeFFT1024 frame cadence); advanced by k_spec1024
    int have_prev;             // 0 until the IQ spectrum has seen a block (analyze_fft256iq.cpp:73-77); advanced by k_spec256
    int count;                 // averaging counter of the IQ spectrum (analyze_fft256iq.cpp:99-113)
};

// ---- saturating / packed-q15 arithmetic (ARMv7E-M DSP instruction semantics) --------------
__device__ __forceinline__ int32_t sat16(int32_t v) { return max(-32768, min(32767, v)); }
__device__ __forceinline__ int32_t lo16(uint32_t a) { return (int32_t)(int16_t)(a & 0xFFFFu); }
__device__ __forceinline__ int32_t hi16(uint32_t a) { return ((int32_t)a) >> 16; }
__device__ __forceinline__ uint32_t mk16(int32_t lo, int32_t hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }

__device__ __forceinline__ uint32_t SHADD16(uint32_t a, uint32_t b) { return mk16((lo16(a) + lo16(b)) >> 1, (hi16(a) + hi16(b)) >> 1); }
__device__ __forceinline__ uint32_t SHSUB16(uint32_t a, uint32_t b) { return mk16((lo16(a) - lo16(b)) >> 1, (hi16(a) - hi16(b)) >> 1); }
__device__ __forceinline__ uint32_t QADD16(uint32_t a, uint32_t b) { return mk16(sat16(lo16(a) + lo16(b)), sat16(hi16(a) + hi16(b))); }
__device__ __forceinline__ uint32_t QSUB16(uint32_t a, uint32_t b) { return mk16(sat16(lo16(a) - lo16(b)), sat16(hi16(a) - hi16(b))); }
__device__ __forceinline__ uint32_t QASX(uint32_t a, uint32_t b) { return mk16(sat16(lo16(a) - hi16(b)), sat16(hi16(a) + lo16(b))); }
__device__ __forceinline__ uint32_t QSAX(uint32_t a, uint32_t b) { return mk16(sat16(lo16(a) + hi16(b)), sat16(hi16(a) - lo16(b))); }
__device__ __forceinline__ uint32_t SHASX(uint32_t a, uint32_t b) { return mk16((lo16(a) - hi16(b)) >> 1, (hi16(a) + lo16(b)) >> 1); }
__device__ __forceinline__ uint32_t SHSAX(uint32_t a, uint32_t b) { return mk16((lo16(a) + hi16(b)) >> 1, (hi16(a) - lo16(b)) >> 1); }
// (cos + j sin) twiddle word times sample word, both products keep their top 16 bits
__device__ __forceinline__ uint32_t CMULPACK(uint32_t c, uint32_t r)
{
    uint32_t re = (uint32_t)(lo16(c) * lo16(r)) + (uint32_t)(hi16(c) * hi16(r));   // SMUAD
    uint32_t im = (uint32_t)(lo16(c) * hi16(r)) - (uint32_t)(hi16(c) * lo16(r));   // SMUSDX
    return (im & 0xFFFF0000u) | (re >> 16);
}

// sqrt_uint32_approx of the Teensy Audio library: table seed + two Newton steps, 0 -> 0
__constant__ uint16_t c_sqrt_guess[33] = {
    55109, 38968, 27555, 19484, 13778, 9742, 6889, 4871, 3445, 2436, 1723, 1218, 862, 609, 431, 305,
    216, 153, 108, 77, 54, 39, 27, 20, 14, 10, 7, 5, 4, 3, 2, 1, 0 };
__device__ __forceinline__ uint32_t sqrt_u32_approx(uint32_t in)
{
    if (in == 0u) return 0u;
    uint32_t n = c_sqrt_guess[__clz((int)in)];
    n = ((in / n) + n) >> 1;
    n = ((in / n) + n) >> 1;
    return n;
}

// The same function with its two divisions done by a float estimate and an exact fix-up (the quotients stay below
// 2^18, where the estimate is off by at most one).  The int <-> float
// conversions use the 2^23 magic number (FADD / LOP, full rate) instead of I2F / F2I (16 lanes/clk/SM), and the seed
// table is read through `guess` (a shared-memory copy of c_sqrt_guess: lanes index it divergently, which constant
// memory would serialise).  Bit-identical to sqrt_u32_approx for every input (tools/front_tc_check, test_gpu_parity).
__device__ __forceinline__ float rdsp_small_u2f(uint32_t v) { return __uint_as_float(0x4B000000u | v) - 8388608.0f; }   // v < 2^23
__device__ __forceinline__ uint32_t rdsp_udiv_small_q(uint32_t n, float n_f, uint32_t d)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(rdsp_small_u2f(d)));
    uint32_t q = __float_as_uint(n_f * r + 8388608.0f) & 0x7FFFFFu;
    // n_f, the reciprocal and the product carry 2^-24 + 2^-23 + 2^-24 of relative error: with q < 2^18 the rounded
    // estimate is within 0.57 of n / d, i.e. off by at most one in either direction
    const int32_t rem = (int32_t)(n - q * d);
    q += rem >= (int32_t)d ? 1u : (rem < 0 ? 0xFFFFFFFFu : 0u);
    return q;
}
__device__ __forceinline__ uint32_t sqrt_u32_approx_fast(uint32_t in, const uint16_t *guess)
{
    const float in_f = fmaf(rdsp_small_u2f(in >> 16), 65536.0f, rdsp_small_u2f(in & 0xFFFFu));
    uint32_t n = guess[__clz((int)in)];                                 // branch free: in == 0 falls out at the end
    n = (rdsp_udiv_small_q(in, in_f, n) + n) >> 1;
    n = (rdsp_udiv_small_q(in, in_f, n) + n) >> 1;
    return in == 0u ? 0u : n;
}

// arm_float_to_q15: truncate toward zero, saturate.  cvt.rzi.s16.f32 is exactly that in one instruction (F2I.S16.TRUNC):
// a float-to-integer cvt clamps to the range of its destination type and turns NaN into 0 (the compare / select /
// min / max form of r01 was eight instructions per sample).
__device__ __forceinline__ int32_t f32_to_q15(float v)
{
    short r;
    asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(r) : "f"(v * 32768.0f));
    return (int32_t)r;
}

// streaming (read-once / write-once) 128-bit global accesses
__device__ __forceinline__ int4 ld_stream16(const void *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ int2 ld_stream8(const void *p)
{
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream16(void *p, int4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream8(void *p, int2 v)
{
    asm volatile("st.global.L1::no_allocate.v2.s32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
