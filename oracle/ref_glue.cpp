/*
 * ref_glue.cpp — TEST INFRASTRUCTURE.  Builds oracle/_ref/librdsp_ref.so: the reference's own
 * in-tree DSP sources, compiled UNMODIFIED from /root/reference (never copied into this repo),
 * on top of the shim headers in oracle/shim/.  Exposes them through a narrow C API so the
 * tests can pin the restatement in rdsp_oracle.c against the real code, and bench.py can time
 * the real code on the host cores.
 *
 * All reference state is file-scope globals / function statics (RDSP_convolutional.h:42-80,
 * RDSP_noise_reduction.h:18-32,69) => ONE channel per loaded copy of this library.
 */
#define RDSP_GENERAL_INCLUDES_H_INCLUDED      /* guard of RDSP_general_includes.h:12-13: skip the hardware include list */
#include "Arduino.h"
#include "AudioStream.h"
#include "Audio.h"
#include "arm_math.h"
#include "arm_const_structs.h"

/* the four queue objects the sketch defines at RadioDSP_SDR_RX.ino:64-67 */
AudioRecordQueue Q_in_L;
AudioRecordQueue Q_in_R;
AudioPlayQueue   Q_out_L;
AudioPlayQueue   Q_out_R;

#include "RDSP_noise_reduction.h"             /* verbatim, from -I/root/reference/src/RadioDSP_SDR_RX */
#include "RDSP_convolutional.h"               /* verbatim */
#include "analyze_fft256iq.h"                 /* verbatim (its .cpp is a separate translation unit) */

static AudioAnalyzeFFT256IQ *g_fft;

extern "C" {

/* setup() order, RadioDSP_SDR_RX.ino:144-145,172,180,183 */
void ref_setup(int naverage)
{
    g_fft = new AudioAnalyzeFFT256IQ();
    g_fft->windowFunction(AudioWindowHanning256);
    g_fft->averageTogether((uint8_t)naverage);
    Init_LMS_NR(15);
    doConvolutionalInitialize();
    reInitializeFilter(300, 4000);
}

void ref_reinit_filter(double lo, double hi) { reInitializeFilter(lo, hi); }

/* one update() tick delivers one block to each record queue (RadioDSP_SDR_RX.ino:85-86) */
void ref_conv_push(const int16_t *L, const int16_t *R) { Q_in_L.shim_push(L); Q_in_R.shim_push(R); }

/* one loop() pass, RadioDSP_SDR_RX.ino:198; returns 1 and fills outL/outR if a block was played */
int ref_conv_loop(int nr_level, int16_t *outL, int16_t *outR)
{
    doConvolutionalProcessing((float)nr_level, true, 300.0, 4000.0);
    if (Q_out_L.shim_available() && Q_out_R.shim_available()) {
        Q_out_L.shim_pop(outL);
        Q_out_R.shim_pop(outR);
        return 1;
    }
    return 0;
}

/* f32 buffers right after the last loop pass (pre-quantisation signal) */
void ref_conv_float_out(float *L, float *R)
{
    memcpy(L, float_buffer_L, sizeof(float) * BUFFER_SIZE);
    memcpy(R, float_buffer_R, sizeof(float) * BUFFER_SIZE);
}

void ref_get_mask(float *m) { memcpy(m, FIR_filter_mask, sizeof(FIR_filter_mask)); }
void ref_get_fir(double *cI, double *cQ)
{
    memcpy(cI, FIR_Coef_I, sizeof(FIR_Coef_I));
    memcpy(cQ, FIR_Coef_Q, sizeof(FIR_Coef_Q));
}
float ref_lms_mu(void) { return LMS_Norm_instance.mu; }
void ref_lms_coeffs(float *c) { memcpy(c, LMS_NormCoeff_f32, sizeof(float) * 96); }

/* one update() tick of AudioAnalyzeFFT256IQ with blocks on both ports; returns available() */
int ref_fft256_update(const int16_t *I, const int16_t *Q)
{
    g_fft->shim_feed(0, I);
    g_fft->shim_feed(1, Q);
    g_fft->update();
    return g_fft->available() ? 1 : 0;
}
void ref_fft256_output(uint16_t *out) { memcpy(out, g_fft->output, sizeof(g_fft->output)); }

}  /* extern "C" */
