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

/* Timing loop for bench.py's "reference"-kind CPU baseline: n_blocks ticks of the stages whose sources are in the reference
 * tree — doConvolutionalProcessing (K5 + K6 + K7) in loop() and, with_fft, AudioAnalyzeFFT256IQ::update (K9) — on
 * iq [n_blocks][128][2], audio [n_blocks][128][2].  Same call order as ref_conv_push / ref_conv_loop above (the gate at
 * RDSP_convolutional.h:231 needs one block of look-ahead).  Returns the number of blocks played. */
int ref_run_blocks(int n_blocks, const int16_t *iq, int16_t *audio, int nr_level, int with_fft)
{
    int16_t L[2][128], R[2][128], oL[128], oR[128];
    int played = 0;
    auto split = [&](int b, int s) { for (int i = 0; i < 128; i++) { L[s][i] = iq[(size_t)b * 256 + 2 * i]; R[s][i] = iq[(size_t)b * 256 + 2 * i + 1]; } };
    split(0, 0);
    Q_in_L.shim_push(L[0]); Q_in_R.shim_push(R[0]);
    for (int b = 0; b < n_blocks; b++) {
        const int cur = b & 1, nxt = cur ^ 1;
        if (with_fft) { g_fft->shim_feed(0, L[cur]); g_fft->shim_feed(1, R[cur]); g_fft->update(); (void)g_fft->available(); }
        if (b + 1 < n_blocks) split(b + 1, nxt); else { memset(L[nxt], 0, sizeof(L[nxt])); memset(R[nxt], 0, sizeof(R[nxt])); }
        Q_in_L.shim_push(L[nxt]); Q_in_R.shim_push(R[nxt]);
        doConvolutionalProcessing((float)nr_level, true, 300.0, 4000.0);
        if (Q_out_L.shim_available() && Q_out_R.shim_available()) {
            Q_out_L.shim_pop(oL); Q_out_R.shim_pop(oR);
            for (int i = 0; i < 128; i++) { audio[(size_t)b * 256 + 2 * i] = oL[i]; audio[(size_t)b * 256 + 2 * i + 1] = oR[i]; }
            played++;
        }
    }
    return played;
}

}  /* extern "C" */
