// rdsp_sketch_api.hpp — host-side C++ mirror of the reference sketch's control vocabulary over the C ABI.
//
// The reference drives its DSP through `AudioSDR SDR;` setters, the global `nr_level`, `reInitializeFilter()`
// and `AudioAnalyzeFFT256IQ FFT;` read-outs (RadioDSP_SDR_RX.ino:117-148,183,212-226; RDSP_controls.h:149-297,
// 330-423,569-612).  These thin classes keep the same names, argument meaning and (absence of) error
// behaviour for ONE channel of a bank, so control code written against the sketch ports line by line:
//
//     rdsp::Bank bank(cfg);                       // N receivers on one GPU
//     rdsp::AudioSDR SDR(bank, ch);               // was: AudioSDR SDR;
//     SDR.enableAGC(); SDR.setAGCmode(AGCmedium); // RadioDSP_SDR_RX.ino:120-121
//     TuningOffset = SDR.setDemodMode(LSBmode);   // :139
//     rdsp::reInitializeFilter(bank, ch, 300, 4000);   // :183
//     bank.update(iq, audio);                     // one AudioStream tick for every channel
//
// Header-only; link with -lrdsp_gpu.  Errors of the C ABI surface as rdsp::Error exceptions from Bank only;
// the per-channel setters swallow nothing — they forward the ABI status through Bank::check.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/rdsp_gpu.h"
#include "rdsp_controls.hpp"

namespace rdsp {

// the reference's enumerators
enum DemodMode { LSBmode = RDSP_DEMOD_LSB, USBmode = RDSP_DEMOD_USB, CW_LSBmode = RDSP_DEMOD_CW_LSB,
                 CW_USBmode = RDSP_DEMOD_CW_USB, AMmode = RDSP_DEMOD_AM, SAMmode = RDSP_DEMOD_SAM };
enum AudioFilter { audioCW = RDSP_FILTER_CW, audio2100 = RDSP_FILTER_2100, audio2700 = RDSP_FILTER_2700,
                   audio3100 = RDSP_FILTER_3100, audioAM = RDSP_FILTER_AM };
enum AGCMode { AGCoff = RDSP_AGC_OFF, AGCfast = RDSP_AGC_FAST, AGCmedium = RDSP_AGC_MEDIUM, AGCslow = RDSP_AGC_SLOW };

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

class Bank {
public:
    explicit Bank(const rdsp_gpu_config_t &cfg) : cfg_(cfg)
    {
        int rc = rdsp_gpu_create(&cfg_, &h_);
        if (rc != RDSP_OK) throw Error(rc, rdsp_gpu_last_error(nullptr));
    }
    ~Bank() { rdsp_gpu_destroy(h_); }
    Bank(const Bank &) = delete;
    Bank &operator=(const Bank &) = delete;

    static rdsp_gpu_config_t defaults(uint32_t n_channels, uint32_t stage_mask = RDSP_STAGE_ALL)
    {
        rdsp_gpu_config_t c;
        rdsp_gpu_default_config(&c);
        c.n_channels = n_channels;
        c.stage_mask = stage_mask;
        return c;
    }
    // one AudioStream update() tick for every channel: iq [C][128][2], audio [C][128][2]
    void update(const int16_t *iq, int16_t *audio) { check(rdsp_gpu_process_block(h_, iq, audio)); }
    void update(uint32_t n_blocks, const int16_t *iq, int16_t *audio) { check(rdsp_gpu_process_blocks(h_, n_blocks, iq, audio)); }
    void synchronize() { check(rdsp_gpu_synchronize(h_)); }

    rdsp_chan_params_t get(uint32_t ch) { rdsp_chan_params_t p; check(rdsp_gpu_get_mode(h_, ch, &p)); return p; }
    void set(uint32_t ch, const rdsp_chan_params_t &p) { check(rdsp_gpu_set_mode(h_, ch, 1, &p)); }
    void set(uint32_t ch_first, uint32_t ch_count, const rdsp_chan_params_t &p) { check(rdsp_gpu_set_mode(h_, ch_first, ch_count, &p)); }

    rdsp_gpu_t *handle() { return h_; }
    uint32_t channels() const { return cfg_.n_channels; }
    void check(int rc) { if (rc < 0) throw Error(rc, rdsp_gpu_last_error(h_)); }

private:
    rdsp_gpu_config_t cfg_;
    rdsp_gpu_t *h_ = nullptr;
};

// `AudioSDR SDR;` for one channel
class AudioSDR {
public:
    AudioSDR(Bank &b, uint32_t ch) : b_(b), ch_(ch) {}
    uint32_t setDemodMode(DemodMode m) { auto p = b_.get(ch_); p.demod = m; b_.set(ch_, p); return 0; /* zero-IF: tuning offset 0 Hz */ }
    void setAudioFilter(AudioFilter f) { auto p = b_.get(ch_); p.audio_filter = f; b_.set(ch_, p); }
    void enableAudioFilter() {}
    void enableAGC() { agc_on_ = true; apply_agc(); }
    void disableAGC() { agc_on_ = false; apply_agc(); }
    void setAGCmode(AGCMode m) { agc_mode_ = m; apply_agc(); }
    void enableALSfilter() { auto p = b_.get(ch_); p.notch_on = 1; b_.set(ch_, p); }
    void disableALSfilter() { auto p = b_.get(ch_); p.notch_on = 0; b_.set(ch_, p); }
    void setALSfilterNotch() { auto p = b_.get(ch_); p.als_peak = 0; b_.set(ch_, p); }
    void setALSfilterPeak() { auto p = b_.get(ch_); p.als_peak = 1; b_.set(ch_, p); }
    void setALSfilterAdaptive() {}
    void enableNoiseBlanker() { auto p = b_.get(ch_); p.nb_on = 1; b_.set(ch_, p); }
    void disableNoiseBlanker() { auto p = b_.get(ch_); p.nb_on = 0; b_.set(ch_, p); }
    void setNoiseBlankerThresholdDb(float db) { auto p = b_.get(ch_); p.nb_threshold_db = db; b_.set(ch_, p); }
    void setInputGain(float g) { auto p = b_.get(ch_); p.in_gain = g; b_.set(ch_, p); }
    void setOutputGain(float g) { auto p = b_.get(ch_); p.out_gain = g; b_.set(ch_, p); }
    void setIQgainBalance(float v) { auto p = b_.get(ch_); p.iq_balance = v; b_.set(ch_, p); }
    void setMute(bool) {}

private:
    void apply_agc() { auto p = b_.get(ch_); p.agc_mode = agc_on_ ? agc_mode_ : RDSP_AGC_OFF; b_.set(ch_, p); }
    Bank &b_;
    uint32_t ch_;
    bool agc_on_ = true;
    int agc_mode_ = RDSP_AGC_MEDIUM;
};

// reInitializeFilter(lo, hi), RDSP_convolutional.h:209-224
inline void reInitializeFilter(Bank &b, uint32_t ch, double dFLoCut, double dFHiCut)
{
    auto p = b.get(ch);
    p.pbt_lo_hz = (float)dFLoCut;
    p.pbt_hi_hz = (float)dFHiCut;
    b.set(ch, p);
}

// the global `nr_level` (RDSP_general_includes.h:111, RDSP_controls.h:265-294): 0, 20, 30, 40, 50
inline void set_nr_level(Bank &b, uint32_t ch, int nr_level)
{
    auto p = b.get(ch);
    p.nr_kind = nr_level > 0 ? RDSP_NR_LMS : RDSP_NR_OFF;
    p.nr_level = nr_level;
    b.set(ch, p);
}

// `AudioAnalyzeFFT256IQ FFT;` read-out side for one channel (analyze_fft256iq.h:61-99)
class AudioAnalyzeFFT256IQ {
public:
    AudioAnalyzeFFT256IQ(Bank &b, uint32_t ch) : b_(b), ch_(ch) {}
    bool available()
    {
        uint8_t ready = 0;
        b_.check(rdsp_gpu_read_spectrum(b_.handle(), ch_, 1, output, &ready));
        return ready != 0;
    }
    float read(unsigned int binNumber) { return binNumber > 255 ? 0.0f : (float)output[binNumber] * (1.0f / 16384.0f); }
    uint16_t output[256] = {0};

private:
    Bank &b_;
    uint32_t ch_;
};

// ---- rdsp_controls.hpp: SketchControls<ChannelRadio> drives one channel of a bank ------------------------------
inline ChannelRadio::ChannelRadio(Bank &b, uint32_t ch) : bank_(b), ch_(ch) {}
inline uint32_t ChannelRadio::setDemodMode(int m) { auto p = bank_.get(ch_); p.demod = m; bank_.set(ch_, p); return 0; }
inline void ChannelRadio::setAudioFilter(int f) { auto p = bank_.get(ch_); p.audio_filter = f; bank_.set(ch_, p); }
inline void ChannelRadio::setAGCmode(int m) { agc_mode_ = m; if (agc_on_) { auto p = bank_.get(ch_); p.agc_mode = m; bank_.set(ch_, p); } }
inline void ChannelRadio::enableAGC() { agc_on_ = true; auto p = bank_.get(ch_); p.agc_mode = agc_mode_; bank_.set(ch_, p); }
inline void ChannelRadio::enableALSfilter() { auto p = bank_.get(ch_); p.notch_on = 1; bank_.set(ch_, p); }
inline void ChannelRadio::disableALSfilter() { auto p = bank_.get(ch_); p.notch_on = 0; bank_.set(ch_, p); }
inline void ChannelRadio::reInitializeFilter(double lo, double hi) { rdsp::reInitializeFilter(bank_, ch_, lo, hi); }
inline void ChannelRadio::set_nr_level(int level) { rdsp::set_nr_level(bank_, ch_, level); }

}  // namespace rdsp
