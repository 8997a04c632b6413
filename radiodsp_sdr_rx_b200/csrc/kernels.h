// kernels.h — argument blocks and launchers of the sm_100a kernels (internal to librdsp_gpu.so).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "rdsp_common.cuh"

// Every kernel of the chain asks for the same L1 / shared-memory split (all shared): kernels of different stages run
// next to each other on an SM, and an SM does not co-host kernels that want different carve-outs.
#include <cstdlib>
template <typename K>
inline void rdsp_uniform_carveout(K kernel)
{
    static const int pct = [] { const char *e = getenv("RDSP_CARVEOUT"); return e ? atoi(e) : 100; }();
    if (pct >= 0) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}
// function attributes belong to a (kernel, device) pair: once per device, and safe when handles on different GPUs
// launch from different host threads
#include <atomic>
constexpr int RDSP_MAX_DEVICES = 64;
inline int rdsp_current_device() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < RDSP_MAX_DEVICES) ? d : 0; }
#define RDSP_CARVEOUT_ONCE(kernel) do { static std::atomic<bool> done_[RDSP_MAX_DEVICES]; const int d_ = rdsp_current_device(); \
        if (!done_[d_].load(std::memory_order_acquire)) { rdsp_uniform_carveout(kernel); done_[d_].store(true, std::memory_order_release); } } while (0)

// Programmatic dependent launch for the kernels of one dependent chain (notch -> AGC -> FFT filter -> DNR -> audio spectrum on
// one stream).  A kernel launched with `pdl` may start while its predecessor still runs: it loads ITS OWN per-channel state
// (written by its previous launch, one call ago), then executes griddepcontrol.wait, which returns when the predecessor has
// completed and its stores are visible, and only then touches the predecessor's output.  Every chain kernel releases its
// successor at its first instruction.  What it buys is the launch latency and the state prologue of each stage — which is most
// of a call that covers a single 128-sample block (the sketch's calling pattern); with many blocks per call the early CTAs
// only take slots from the running stage (measured in r01: 0.67 -> 0.79 ms at 8 blocks), so the host enables it for short calls.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_release_successor() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_predecessor() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename A>
inline void rdsp_launch(void (*kernel)(A), int grid, int block, size_t smem, cudaStream_t st, bool pdl, const A &args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1u : 0u;
    cudaLaunchKernelEx(&cfg, kernel, args);
}
#endif

// K0+K1+K2 -----------------------------------------------------------------------------------
struct FrontArgs {
    const int16_t *iq;          // [T][C][128][2]
    int16_t *out_mono;          // [T][C][128]       or nullptr
    int16_t *out_stereo;        // [T][C][128][2]    or nullptr (front end is the last stage)
    float *dbg;                 // [T][C][128][2]    or nullptr
    int16_t *hist;              // [C][3][128] delay lines: I', Q', demodulated
    int16_t *hist_out;          // k_front_tc only: where the new delay lines go (nullptr = in place; a second buffer lets
                                // the blocks of a call be cut into concurrent time segments)
    float *sam_state;           // [C][4] SAM carrier loop: phase, frequency, carrier level, pad (k_front_tc only)
    int32_t *nb_ref;            // [C] noise blanker: running IQ magnitude (k_front_tc only)
    const RdspChanParams *par;  // [C]
    const int32_t *taps;        // [15][132]: hilbert_i[5], hilbert_q[5], bandpass[5]
    int C, T;
};
void launch_front(const FrontArgs &a, cudaStream_t st);

// K0+K1+K2 on tcgen05 (k_front_tc.cu): channels grouped into tiles of 128 that share their tap rows
struct FrontTcTables {
    const int *tile_ch;         // [n_tiles][128] channel of every MMA row, -1 = padding
    const int4 *tile_rows;      // [n_tiles] tap row of the I', Q' and band-pass FIR; w = detector (0 sideband sum, 1 AM envelope, 2 SAM)
                                // | 256 if the tile runs from the paired images of its rows
    const uint8_t *toep;        // [15][2 images][10240] banded Toeplitz byte images of the tap rows (front_tc_build_toeplitz)
    CUtensorMap toep_map;       // TMA descriptor of `toep` as a 2-D byte tensor [15 * 40 rows][256]: one box = one tap row's
                                // two planes (front_tc_make_tensor_map); the kernel loads its three images with it
    int n_tiles;
    int n_am_tiles;             // how many of them are AM-envelope tiles (their chunks take ~1.5 x as long: more time segments)
    int seg_bounds[2][9];       // filled by launch_front_tc: block range of every time segment, per kind (0 sideband / SAM, 1 AM)
    int n_seg[2];
    int sam_tiles;              // a SAM tile exists: launch the instantiation that carries the SAM detector
    int any_sam;                // a SAM tile or a noise-blanked channel exists: their state is sequential over the whole call, one segment
};
void launch_front_tc(const FrontArgs &a, const FrontTcTables &tb, cudaStream_t st);
size_t front_tc_toeplitz_bytes();
int front_tc_make_tensor_map(const uint8_t *d_toep, CUtensorMap *map);   // 0 = ok (cuTensorMapEncodeTiled through the runtime's driver entry point)
void front_tc_build_toeplitz(const int16_t *taps, int stride, uint8_t *out);
int front_tc_build_tiles(const RdspChanParams *par, int C, const int16_t *taps, int stride,
                         std::vector<int> &tile_ch, std::vector<int4> &tile_rows);

// K3 / K6: normalised LMS ----------------------------------------------------------------------
struct NlmsArgs {
    const int *list;            // channels to run (nullptr: 0..n_list-1)
    int n_list;
    int C, T;
    const int16_t *in_q15;      // [T][C][128]  (notch) or nullptr
    const float *in_f32;        // [T][C][128]  (DNR)   or nullptr
    float *out_f32;             // [T][C][128]  error signal (notch)
    int16_t *out_stereo;        // [T][C][128][2] 1.1*y, L = R (DNR) ...
    int16_t *out_mono;          // ... or [T][C][128] (RDSP_AUDIO_MONO); exactly one is set in mode 1
    float *dbg;                 // [T][C][128][2] or nullptr
    float *coeff;               // [C][96] CMSIS order (index 0 multiplies the oldest sample)
    float *prev;                // [C][128] previous input block (= de-correlation ring + FIR history)
    float *energy;              // [C]
    uint8_t *first;             // [C] 1 until the instance ran once (RDSP_noise_reduction.h:69 statics)
    const RdspChanParams *par;
    int mode;                   // 0 = notch (output error), 1 = DNR (output estimate)
    int16_t *ring;              // one-block calls (mode 1): [C][8][128] ring of K10; the kernel appends its own L row (slot = tick mod 8) ...
    const RdspTick *tick_in;    // ... tick index of this call's block; both nullptr otherwise
    int contended;              // 1: other kernels run beside this one (spectrum branches): the FFMA2 form pays (fewer issue slots)
    int direct;                 // 1: the sample-by-sample cross-check kernel (k_nlms_direct.cu; RDSP_NLMS_IMPL=direct at create)
    int pdl;                    // launched as a programmatic dependent of the kernel before it on the stream
};
void launch_nlms(const NlmsArgs &a, cudaStream_t st);

// K4: AGC + output gain -------------------------------------------------------------------------
struct AgcArgs {
    const int *list;            // channels to run (nullptr: ch0 .. ch0 + n_list - 1)
    int n_list;
    int ch0;
    const int16_t *in_q15;      // [T][C][128] q15 rows (channels that bypass the notch) ...
    const float *in_f32;        // ... or [T][C][128] f32 rows (channels whose notch ran); exactly one is set
    int16_t *out_mono;          // [T][C][128]    or nullptr
    int16_t *out_stereo;        // [T][C][128][2] or nullptr
    float *dbg;                 // [T][C][128][2] or nullptr
    float *env;                 // [C]
    const RdspChanParams *par;
    int C, T;
    int agc_stage;              // 0: only quantise (notch without AGC stage)
    float target, max_gain, alpha_a;
    int pdl;                    // launched as a programmatic dependent of the kernel before it on the stream
};
void launch_agc(const AgcArgs &a, cudaStream_t st);

// K5 (+K8) + K7 ---------------------------------------------------------------------------------
struct FftFiltArgs {
    const int16_t *in_mono;     // [T][C][128]     (L = R) or nullptr
    const int16_t *in_stereo;   // [T][C][128][2]  or nullptr
    int16_t *out_stereo;        // [T][C][128][2] ...
    int16_t *out_mono;          // ... or [T][C][128], L only (RDSP_AUDIO_MONO); exactly one is set
    float *out_f32_L;           // [T][C][128] for channels whose DNR follows
    float *dbg;                 // [T][C][128][2] or nullptr
    int16_t *last;              // [C][128][2] previous input block (q15, exact)
    float *nfloor;              // [C] spectral-NR noise floor
    const float2 *masks;        // [n_masks][256]
    const float2 *tw256;        // [256] (cos, sin)(2*pi*k/256)
    const float *sin512;        // [513] sinTable_f32 of arm_sin_f32 / arm_cos_f32 (K8 rebuilds bins through it)
    const RdspChanParams *par;
    int C, T;
    const int *list;            // channels of this launch (n entries), or nullptr: [ch0, ch0 + n)
    int ch0, n;
    int nr_stage;               // RDSP_STAGE_NR present
    int pdl;                    // launched as a programmatic dependent of the kernel before it on the stream
    int16_t *ring;              // one-block calls: [C][8][128] ring of K10; channels whose audio this kernel emits append their L row ...
    const RdspTick *tick_in;    // ... at slot = tick mod 8; both nullptr otherwise
};
void launch_fftfilt(const FftFiltArgs &a, cudaStream_t st);

// a11 -------------------------------------------------------------------------------------------
struct BiquadArgs {
    const int16_t *iq;          // [T][C][128][2] raw IQ
    int16_t *out;               // [T][C][128][2] high-passed IQ
    int32_t *state;             // [C][2][4]: bprev, aprev, sum, pad  for I and Q
    int C, T;
    int ch0, n;                 // channel range of this launch: [ch0, ch0 + n)
    int32_t b0, b1, b2, a1, a2; // Q30, feedback already negated
};
void launch_biquad(const BiquadArgs &a, cudaStream_t st);

// K9 --------------------------------------------------------------------------------------------
struct Spec256Args {
    const int16_t *iq;          // [T][C][128][2] high-passed IQ (output of k_biquad)
    int16_t *prev;              // [C][128][2] previous post-biquad block
    uint32_t *sum;              // [C][256]
    uint16_t *output;           // [C][256]
    int C, T;
    int ch0, n;                 // channel range of this launch: [ch0, ch0 + n)
    const RdspTick *tick_in;    // have_prev / count at the first tick of this call ...
    RdspTick *tick_out;         // ... and where their values after the call go (the other copy)
    int naverage;
    unsigned long long div_magic;   // ceil(2^div_shift / naverage), div_shift = 32 + ceil(log2 naverage):
    int div_shift;                  // (magsq * div_magic) >> div_shift == magsq / naverage for magsq <= 2^31
    const int2 *tw;             // [3072] twiddleCoef_4096_q15 as (cos, sin) int pairs
    const int16_t *win;         // [256] Hann
    int2 tw3[4][3];             // tw[256 d0 k], k = 1..3: the twiddles of stage 3 are the same for every lane, so they travel as kernel
                                // parameters (constant-bank operands of the multiplies) instead of 12 shared-memory reads per frame
};
void launch_spec256(const Spec256Args &a, cudaStream_t st);

// K10 -------------------------------------------------------------------------------------------
struct Spec1024Args {
    const int16_t *audio;       // [T][C][128][2] (L used), or [T][C][128] when audio_mono
    int audio_mono;
    int16_t *ring;              // [C][8][128] last blocks of L, slot = tick mod 8
    uint16_t *output;           // [C][512]
    int C, T;
    const int *list;            // channels of this launch (n entries), or nullptr: [ch0, ch0 + n)
    int ch0, n;
    const RdspTick *tick_in;    // tick index of block 0 of this call ...
    RdspTick *tick_out;         // ... and where tick + T goes (the other copy)
    const int2 *tw;             // [3072]
    const int16_t *win;         // [1024] Hann
    int pdl;                    // launched as a programmatic dependent of the kernel before it on the stream
    int appended;               // one-block calls: the kernels that emitted the audio appended the row already (FftFiltArgs / NlmsArgs ring)
};
void launch_spec1024(const Spec1024Args &a, cudaStream_t st);

// K11 -------------------------------------------------------------------------------------------
struct PanArgs {
    const uint16_t *spec;       // [C][256]
    uint16_t *view;             // [C][256] SpectrumView (state: SpectrumViewOld)
    float *smeter;              // [C]
    uint16_t *waterfall;        // [C][50][128] ring, or nullptr
    const int *wf_head;         // [ch_count] ring slot that receives the new line of channel ch_first + i
    int ch_first, ch_count;
};
void launch_panadapter(const PanArgs &a, cudaStream_t st);

struct WaterfallArgs {
    const uint16_t *ring;       // [C][50][128]
    const int *wf_head;         // [ch_count] slot of the newest line
    uint16_t *rows;             // [ch_count][50][128], row 0 newest
    uint8_t *colour;            // [ch_count][50][128] colour class 0..6, or nullptr
    int ch_first, ch_count;
};
void launch_waterfall_read(const WaterfallArgs &a, cudaStream_t st);

