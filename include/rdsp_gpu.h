/*
 * rdsp_gpu.h — C ABI of the batched RadioDSP_SDR_RX receive chain on B200 (sm_100a).
 *
 * This is the drop-in boundary for the reference's per-block audio graph.  One
 * rdsp_gpu_t handle owns N independent receiver channels on ONE GPU; one call
 * of rdsp_gpu_process_block() is exactly one AudioStream update() tick (128
 * samples, reference: analyze_fft256iq.cpp:65 / analyze_fft256iq.h:98) of the
 * whole graph wired at RadioDSP_SDR_RX.ino:71-89, for every channel at once.
 *
 * Reference interface each entry point replaces (paths relative to
 * /root/reference/src/RadioDSP_SDR_RX/):
 *
 *   rdsp_gpu_create            object construction + setup() defaults     RadioDSP_SDR_RX.ino:52-67,117-183
 *   rdsp_gpu_set_mode          SDR.setDemodMode / setAudioFilter /        RDSP_controls.h:149-191,196-232,
 *                              setAGCmode / enableALSfilter / nr_level /  237-297,330-423,569-612;
 *                              reInitializeFilter(lo,hi) / gains          RDSP_convolutional.h:209-224
 *   rdsp_gpu_process_block(s)  AudioStream::update() tick of every node + RadioDSP_SDR_RX.ino:71-89,198;
 *                              doConvolutionalProcessing() in loop()      RDSP_convolutional.h:228-353
 *   rdsp_gpu_read_spectrum     FFT.available() + FFT.output[256]          analyze_fft256iq.h:61-67,99
 *   rdsp_gpu_read_audio_spectrum  AudioFFT.available() + output[512]      RadioDSP_SDR_RX.ino:58,87,222; RDSP_display.h:219
 *   rdsp_gpu_read_panadapter   Update_Panadapter / Update_smeter maths    RDSP_display.h:260-280,366-374
 *   rdsp_gpu_read_waterfall    WaterfallData history + colour classes     RDSP_display.h:30,282-319
 *   rdsp_gpu_set_taps          (coefficient tables are data; AudioSDR filter presets)
 *   rdsp_gpu_set_mask          init_filter_mask() result as data          RDSP_convolutional.h:87-110
 *   rdsp_gpu_destroy           (none: the sketch never tears down)
 *
 * Plain C, plain pointers and sizes.  No CPU fallback exists: every entry point
 * that computes fails with RDSP_ERR_CUDA when no sm_100-class device is usable.
 *
 * Threading: a handle is single-threaded (calls serialised by the caller);
 * different handles are independent (one per GPU, or several per GPU).
 */
#ifndef RDSP_GPU_H_INCLUDED
#define RDSP_GPU_H_INCLUDED

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDSP_BLOCK_SAMPLES   128          /* AUDIO_BLOCK_SAMPLES, RDSP_convolutional.h:34 */
#define RDSP_SAMPLE_RATE_HZ  44100.0      /* AUDIO_SAMPLE_RATE_EXACT, RDSP_convolutional.h:35 */
#define RDSP_FFT_LEN         256          /* FFT_L, RDSP_convolutional.h:36 */
#define RDSP_FIR_TAPS        129          /* m_NumTaps = FFT_L/2+1, RDSP_convolutional.h:72 */
#define RDSP_LMS_TAPS        96           /* calc_taps, RDSP_noise_reduction.h:39 */
#define RDSP_SPEC256_BINS    256
#define RDSP_SPEC1024_BINS   512

typedef struct rdsp_gpu rdsp_gpu_t;

/* status codes: 0 = OK, < 0 = error (never throws across the ABI) */
enum {
    RDSP_OK            =  0,
    RDSP_ERR_INVALID   = -1,   /* NULL pointer / malformed argument */
    RDSP_ERR_RANGE     = -2,   /* channel range or parameter value out of range */
    RDSP_ERR_CUDA      = -3,   /* CUDA runtime error or no usable device */
    RDSP_ERR_NOMEM     = -4,
    RDSP_ERR_STATE     = -5    /* call not valid for this handle's stage mask */
};

/* demodulation modes, RDSP_controls.h:330-423 */
enum {
    RDSP_DEMOD_LSB    = 0,     /* LSBmode     */
    RDSP_DEMOD_USB    = 1,     /* USBmode     */
    RDSP_DEMOD_CW_LSB = 2,     /* CW_LSBmode  */
    RDSP_DEMOD_CW_USB = 3,     /* CW_USBmode  */
    RDSP_DEMOD_AM     = 4,     /* AMmode      */
    RDSP_DEMOD_COUNT  = 5,     /* rows of the Hilbert tap tables (rdsp_gpu_set_taps index) */
    RDSP_DEMOD_SAM    = 5,     /* SAMmode, RDSP_controls.h:384-391: synchronous AM — the AM tap rows, carrier PLL
                                  + coherent detection instead of the envelope (DESIGN.md "SAM") */
    RDSP_DEMOD_MODES  = 6
};

/* audio filter presets, RDSP_controls.h:149-191 */
enum {
    RDSP_FILTER_CW    = 0,     /* audioCW    "500 Hz"  */
    RDSP_FILTER_2100  = 1,     /* audio2100  "2.1 kHz" */
    RDSP_FILTER_2700  = 2,     /* audio2700  "2.7 kHz" */
    RDSP_FILTER_3100  = 3,     /* audio3100  "3.1 kHz" */
    RDSP_FILTER_AM    = 4,     /* audioAM    "3.9 kHz" */
    RDSP_FILTER_COUNT = 5
};

/* AGC modes, RDSP_controls.h:196-232 */
enum { RDSP_AGC_OFF = 0, RDSP_AGC_FAST = 1, RDSP_AGC_MEDIUM = 2, RDSP_AGC_SLOW = 3, RDSP_AGC_COUNT = 4 };

/* noise-reduction kinds */
enum {
    RDSP_NR_OFF      = 0,
    RDSP_NR_LMS      = 1,      /* in-tree NLMS "DNR", RDSP_noise_reduction.h:35-80; nr_level in {20,30,40,50} */
    RDSP_NR_SPECTRAL = 2       /* backup spectral subtraction, backup/RDSP_convolutional_spec.h:181-252; nr_level in {1,2,3} */
};

/* stage mask: which nodes of the graph this handle runs (bit-or) */
enum {
    RDSP_STAGE_FRONTEND = 1u << 0,  /* K0+K1+K2: gain/IQ balance, Hilbert pair, sideband/envelope, band-pass bank (q15) */
    RDSP_STAGE_NOTCH    = 1u << 1,  /* K3: ALS LMS auto-notch (f32)            — needs FRONTEND */
    RDSP_STAGE_AGC      = 1u << 2,  /* K4: AGC + output gain (f32 -> q15)      — needs FRONTEND */
    RDSP_STAGE_FFTFILT  = 1u << 3,  /* K5: FFT-256 overlap-save band-pass, RDSP_convolutional.h:228-318 */
    RDSP_STAGE_NR       = 1u << 4,  /* K6/K8: NLMS DNR or spectral subtraction — needs FFTFILT */
    RDSP_STAGE_SPEC256  = 1u << 5,  /* a11+K9: HP biquads + 256-pt IQ q15 spectrum, analyze_fft256iq.cpp:65-118 */
    RDSP_STAGE_SPEC1024 = 1u << 6,  /* K10: 1024-pt q15 audio spectrum of output L */
    RDSP_STAGE_ALL      = 0x7Fu
};

/* where iq_in / audio_out of process_block(s) live */
enum { RDSP_IO_DEVICE = 0, RDSP_IO_HOST = 1 };

/* layout of audio_out */
enum {
    RDSP_AUDIO_STEREO = 0,     /* [n_blocks][n_channels][128][2]  L,R interleaved (the two AudioPlayQueue ports) */
    RDSP_AUDIO_MONO   = 1      /* [n_blocks][n_channels][128]     L only: the sketch plays L == R whenever the DNR runs
                                  (RDSP_convolutional.h:333-336) and whenever the chain ends in the SDR block; with the
                                  FFT filter last and no DNR, R (the quadrature output of the analytic filter) is dropped */
};

/* CUDA graphs on the block path */
enum {
    RDSP_GRAPH_AUTO = 0,       /* a call shape (n_blocks, buffers, tables) seen twice is captured and replayed with one launch */
    RDSP_GRAPH_OFF  = 1        /* always enqueue kernel by kernel */
};

/* tap-table kinds for rdsp_gpu_set_taps / get_taps */
enum {
    RDSP_TAPS_HILBERT_I = 0,   /* filter applied to I, index = demod mode */
    RDSP_TAPS_HILBERT_Q = 1,   /* filter applied to Q, index = demod mode */
    RDSP_TAPS_BANDPASS  = 2    /* audio band-pass, index = audio filter preset */
};

typedef struct {
    uint32_t struct_size;        /* = sizeof(rdsp_gpu_config_t) */
    uint32_t n_channels;         /* receivers on this handle (>= 1) */
    int32_t  device;             /* CUDA device ordinal */
    uint32_t stage_mask;         /* RDSP_STAGE_* */
    uint32_t max_blocks_per_call;/* upper bound for n_blocks of process_blocks (scratch sizing), >= 1 */
    uint32_t io_location;        /* RDSP_IO_DEVICE or RDSP_IO_HOST */
    uint32_t async;              /* 1: process_* returns after enqueue; call rdsp_gpu_synchronize */
    uint32_t debug_f32;          /* 1: keep the f32 pre-quantisation output for rdsp_gpu_read_debug_f32 */
    uint32_t spec256_naverage;   /* FFT.averageTogether(30), RadioDSP_SDR_RX.ino:145 */
    /* AGC constants (shim-defined, see DESIGN.md "AGC") */
    float    agc_target;         /* envelope target level (full scale = 1.0) */
    float    agc_max_gain;       /* linear */
    float    agc_attack_ms;
    float    agc_decay_ms[4];    /* per RDSP_AGC_* mode; [RDSP_AGC_OFF] unused */
    uint32_t pipeline_chunks;    /* channel groups: behind the front end the channels of a call walk the graph as this
                                    many independent groups on streams of their own (0 = default = 1 group; max 8;
                                    reduced silently while a group would hold < 256 channels).  Every launch still
                                    covers all n_blocks of the call.  Measured: 1 is fastest */
    uint32_t audio_layout;       /* RDSP_AUDIO_STEREO (default) or RDSP_AUDIO_MONO */
    uint32_t graph_mode;         /* RDSP_GRAPH_AUTO (default) or RDSP_GRAPH_OFF */
} rdsp_gpu_config_t;

typedef struct {
    int32_t demod;               /* RDSP_DEMOD_*   default LSB    RadioDSP_SDR_RX.ino:139 */
    int32_t audio_filter;        /* RDSP_FILTER_*  default 2700   RadioDSP_SDR_RX.ino:138 */
    int32_t agc_mode;            /* RDSP_AGC_*     default MEDIUM RadioDSP_SDR_RX.ino:121 */
    int32_t notch_on;            /* 0/1            default 0      RadioDSP_SDR_RX.ino:125 */
    int32_t notch_level;         /* strength -> mu as RDSP_noise_reduction.h:48-56, default 20 */
    int32_t nr_kind;             /* RDSP_NR_*      default OFF */
    int32_t nr_level;            /* LMS: 20/30/40/50 (RDSP_controls.h:265-294); SPECTRAL: 1..3 */
    float   pbt_lo_hz;           /* default 300   RadioDSP_SDR_RX.ino:183; range RDSP_general_includes.h:76-82 */
    float   pbt_hi_hz;           /* default 4000 */
    float   in_gain;             /* default 1.0   RadioDSP_SDR_RX.ino:133 */
    float   out_gain;            /* default 0.5   RadioDSP_SDR_RX.ino:134 */
    float   iq_balance;          /* default 1.020 RadioDSP_SDR_RX.ino:135 */
    int32_t als_peak;            /* 0: SDR.setALSfilterNotch() (RDSP_controls.h:258, the notch emits the NLMS error),
                                    1: ALS "peak" (backup sketch: the notch stage emits the NLMS estimate); default 0 */
    int32_t nb_on;               /* SDR.enableNoiseBlanker / disableNoiseBlanker, RadioDSP_SDR_RX.ino:129-131; default 0 */
    float   nb_threshold_db;     /* SDR.setNoiseBlankerThresholdDb(20.0), RadioDSP_SDR_RX.ino:130; 0 .. 40 dB over the
                                    running IQ magnitude (DESIGN.md "Noise blanker"); default 20 */
} rdsp_chan_params_t;

/* fill *cfg / *p with the reference's setup() defaults */
void rdsp_gpu_default_config(rdsp_gpu_config_t *cfg);
void rdsp_gpu_default_params(rdsp_chan_params_t *p);

/* The band-pass designer of the default tap bank for an arbitrary audio band (Hz): 129 q15 taps for
 * rdsp_gpu_set_taps(RDSP_TAPS_BANDPASS, ...).  audioWSPR (RDSP_controls.h:392-402: 200 Hz around 1500 Hz) is
 * rdsp_gpu_design_bandpass(1400, 1600, taps) loaded into a preset row.  No handle, no device. */
int  rdsp_gpu_design_bandpass(float lo_hz, float hi_hz, int16_t *taps, uint32_t n_taps);

int  rdsp_gpu_create(const rdsp_gpu_config_t *cfg, rdsp_gpu_t **out);
void rdsp_gpu_destroy(rdsp_gpu_t *h);

/* Apply *p to channels [ch_first, ch_first+ch_count).  Takes effect at the next
 * block boundary (mirrors RDSP_convolutional.h:327).  Invalid values are
 * rejected here, never on the block path. */
int  rdsp_gpu_set_mode(rdsp_gpu_t *h, uint32_t ch_first, uint32_t ch_count, const rdsp_chan_params_t *p);
int  rdsp_gpu_get_mode(rdsp_gpu_t *h, uint32_t ch, rdsp_chan_params_t *p);

/* One update() tick for every channel.
 *   iq_in     [n_channels][128][2] int16  (I,Q interleaved = I2S frame order)
 *   audio_out [n_channels][128][2] int16  (L,R interleaved); [n_channels][128] (L) with cfg.audio_layout = RDSP_AUDIO_MONO
 * Both 16-byte aligned, on the device or on the host per cfg.io_location. */
int  rdsp_gpu_process_block(rdsp_gpu_t *h, const int16_t *iq_in, int16_t *audio_out);

/* n_blocks sequential ticks: iq_in [n_blocks][n_channels][128][2], same for audio_out. */
int  rdsp_gpu_process_blocks(rdsp_gpu_t *h, uint32_t n_blocks, const int16_t *iq_in, int16_t *audio_out);

/* Blocks the host until every call issued so far is complete, including the device-to-host copies of
 * RDSP_IO_HOST handles. */
int  rdsp_gpu_synchronize(rdsp_gpu_t *h);

/* RDSP_IO_HOST + async: copies run on the handle's own copy streams so that they overlap the kernels of the
 * neighbouring calls (a ring of device staging buffers: two, or four for handles of short calls, max_blocks_per_call <= 4,
 * where copy-in, kernels and copy-out of a call take about as long as each other).  A call returns after enqueue: its
 * iq_in is read and its audio_out written later, in issue order — keep both untouched until rdsp_gpu_synchronize, or
 * until an event recorded after rdsp_gpu_stream_join has completed.  rdsp_gpu_stream_join makes the handle's stream wait (on the device, without blocking
 * the host) for all outstanding device-to-host copies, so that an event recorded on the stream afterwards
 * covers them.  A no-op for RDSP_IO_DEVICE handles. */
int  rdsp_gpu_stream_join(rdsp_gpu_t *h);

/* Use the caller's CUDA stream (a cudaStream_t passed as void*); NULL = the handle's own. */
int  rdsp_gpu_set_stream(rdsp_gpu_t *h, void *cuda_stream);

/* 256-bin IQ spectrum: out [ch_count][256] uint16 (host), ready [ch_count] (host):
 * ready=1 if a new spectrum was completed since the last read (FFT.available()). */
int  rdsp_gpu_read_spectrum(rdsp_gpu_t *h, uint32_t ch_first, uint32_t ch_count, uint16_t *out, uint8_t *ready);
/* 512-bin audio spectrum of output L (AudioAnalyzeFFT1024). */
int  rdsp_gpu_read_audio_spectrum(rdsp_gpu_t *h, uint32_t ch_first, uint32_t ch_count, uint16_t *out, uint8_t *ready);
/* Panadapter trace (RDSP_display.h:260-280) u16[256] and S-meter level (RDSP_display.h:366-374,
 * value passed to displayPeak before its IIR) computed on the device from the last spectrum. */
int  rdsp_gpu_read_panadapter(rdsp_gpu_t *h, uint32_t ch_first, uint32_t ch_count, uint16_t *trace, float *smeter);
/* Waterfall history (RDSP_display.h:30,282-319): every rdsp_gpu_read_panadapter call pushes the new trace line
 * (SpectrumView[2x], x <= 127) on top of a 50-row history kept on the device.  rows [ch_count][50][128] uint16 (row 0
 * newest), colour (optional) [ch_count][50][128] = the sketch's colour class of each cell: 6 red >= 75, 5 magenta >= 50,
 * 4 orange >= 40, 3 yellow >= 25, 2 blue >= 15, 1 navy >= 5, 0 black. */
int  rdsp_gpu_read_waterfall(rdsp_gpu_t *h, uint32_t ch_first, uint32_t ch_count, uint16_t *rows, uint8_t *colour);

/* Coefficient tables are data.  taps: n_taps (= RDSP_FIR_TAPS) q15 values. */
int  rdsp_gpu_set_taps(rdsp_gpu_t *h, int kind, int index, const int16_t *taps, uint32_t n_taps);
int  rdsp_gpu_get_taps(rdsp_gpu_t *h, int kind, int index, int16_t *taps, uint32_t n_taps);
/* Explicit FFT-domain mask (512 floats, interleaved re/im) for a channel range. */
int  rdsp_gpu_set_mask(rdsp_gpu_t *h, uint32_t ch_first, uint32_t ch_count, const float *mask512);
int  rdsp_gpu_get_mask(rdsp_gpu_t *h, uint32_t ch, float *mask512);

/* f32 pre-quantisation output of the last process call (cfg.debug_f32 = 1):
 * out [n_blocks][ch_count][128][2] floats on the host. */
int  rdsp_gpu_read_debug_f32(rdsp_gpu_t *h, uint32_t n_blocks, uint32_t ch_first, uint32_t ch_count, float *out);

/* Instrumentation */
uint64_t    rdsp_gpu_kernel_launches(const rdsp_gpu_t *h);     /* kernels launched so far by this handle (graph replays count their kernels) */
uint64_t    rdsp_gpu_graph_replays(const rdsp_gpu_t *h);       /* process calls that ran as ONE cudaGraphLaunch */
int         rdsp_gpu_profile(rdsp_gpu_t *h, int enable);        /* 1: bracket every kernel with CUDA events */
/* Accumulated per-kernel device time since profiling was enabled. Returns the number of
 * kernel kinds; for i < n: names[i] (static string), ms[i], launches[i]. */
int         rdsp_gpu_profile_read(rdsp_gpu_t *h, int max, const char **names, double *ms, uint64_t *launches);
const char *rdsp_gpu_last_error(const rdsp_gpu_t *h);          /* h may be NULL: last create() error */
const char *rdsp_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RDSP_GPU_H_INCLUDED */
