/*
 * rdsp_oracle.h — TEST INFRASTRUCTURE.  CPU oracle ("port") of the receive chain.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may call this.  The product (radiodsp_sdr_rx_b200/) never does.
 *
 * One rdsp_oracle_chan_t = all state of ONE receiver channel, so many channels
 * can live in one process (the reference keeps its state in file-scope globals,
 * RDSP_convolutional.h:42-80, RDSP_noise_reduction.h:18-32,69).
 *
 * Parity status per stage:
 *   K5 FFT-256 overlap-save filter, K6 NLMS DNR, K7 f32->q15, K9 IQ spectrum:
 *     restated from in-tree code and PINNED against the reference's own sources
 *     compiled unmodified for x86 (oracle/_ref, tests/test_oracle_vs_ref.py)
 *     on top of the CMSIS/Teensy primitive shim (primitives themselves unpinned).
 *   K0-K4 (AudioSDR), K8 (backup sketch, not runnable as shipped), K10, a11:
 *     PARITY UNPINNED — restated from documentation / SURVEY.md Appendix A, G.
 */
#ifndef RDSP_ORACLE_H
#define RDSP_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/rdsp_gpu.h"      /* parameter structs and enums of the boundary */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rdsp_oracle_chan rdsp_oracle_chan_t;

void rdsp_oracle_default_params(rdsp_chan_params_t *p);
void rdsp_oracle_default_config(rdsp_gpu_config_t *cfg);

/* cfg: stage_mask, spec256_naverage and the agc_* fields are used */
rdsp_oracle_chan_t *rdsp_oracle_chan_create(const rdsp_gpu_config_t *cfg);
void rdsp_oracle_chan_destroy(rdsp_oracle_chan_t *c);
int  rdsp_oracle_chan_set_mode(rdsp_oracle_chan_t *c, const rdsp_chan_params_t *p);

/* n_blocks ticks: iq [n_blocks][128][2], audio [n_blocks][128][2]; stride_* in int16 units
 * between consecutive blocks (so one channel of a [blocks][channels][128][2] array can be
 * addressed in place).  f32 (optional, same stride in floats) = pre-quantisation output. */
void rdsp_oracle_chan_process(rdsp_oracle_chan_t *c, uint32_t n_blocks,
                              const int16_t *iq, size_t stride_in,
                              int16_t *audio, size_t stride_out,
                              float *f32, size_t stride_f32);

void rdsp_oracle_chan_dnr_f32(rdsp_oracle_chan_t *c, uint32_t n_blocks, const float *x, float *y);   /* K6 alone (+ the 1.1 gain) on f32 blocks */
void rdsp_oracle_chan_spec256_raw(rdsp_oracle_chan_t *c, const int16_t *i_blk, const int16_t *q_blk);  /* K9 without biquads */
int  rdsp_oracle_chan_read_spectrum(rdsp_oracle_chan_t *c, uint16_t *out256);        /* returns available() */
int  rdsp_oracle_chan_read_audio_spectrum(rdsp_oracle_chan_t *c, uint16_t *out512);
void rdsp_oracle_chan_read_panadapter(rdsp_oracle_chan_t *c, uint16_t *trace256, float *smeter);
void rdsp_oracle_chan_read_waterfall(rdsp_oracle_chan_t *c, uint16_t *rows50x128, uint8_t *colour50x128);
void rdsp_oracle_chan_get_mask(rdsp_oracle_chan_t *c, float *mask512);
void rdsp_oracle_chan_set_mask(rdsp_oracle_chan_t *c, const float *mask512);

/* design helpers (shared by every channel) */
void rdsp_oracle_calc_cplx_fir(double *cI, double *cQ, int n, double lo, double hi, double fs);  /* RDSP_convolutional.h:127-185 */
void rdsp_oracle_design_mask(double lo, double hi, float *mask512);                              /* + :87-110 */
void rdsp_oracle_get_taps(int kind, int index, int16_t *taps129);
void rdsp_oracle_set_taps(int kind, int index, const int16_t *taps129);
float rdsp_oracle_lms_mu(int strength);                                                          /* RDSP_noise_reduction.h:48-56 */

/* Multi-channel convenience for benchmarks: channels [0,n) of arrays laid out
 * [n_blocks][n_channels_total][128][2]; processes channels ch_first.. ch_first+ch_count. */
void rdsp_oracle_bank_process(rdsp_oracle_chan_t **chans, uint32_t ch_first, uint32_t ch_count,
                              uint32_t n_channels_total, uint32_t n_blocks,
                              const int16_t *iq, int16_t *audio);

#ifdef __cplusplus
}
#endif
#endif
