// rdsp_controls.hpp — the sketch's control plane (mode / filter / AGC / NR cycling and PBT stepping) as a host-side
// C++ state machine that drives one receiver channel through the AudioSDR-style mirror of rdsp_sketch_api.hpp.
//
// Restates the tables of RDSP_controls.h: filterMode() :149-191, setAgc() :196-232, setNRMode() :237-297,
// tuningMode() :330-423, checkPBT_Increase/Decrease() :569-612 with the limits of RDSP_general_includes.h:76-82 and the
// start indices of :103-111.  Everything that touches hardware there (buttons, TFT labels, delays) is left out; the
// label strings are kept because a UI on top of the batched receiver needs them.  Quirks are preserved:
//   * the index is advanced AFTER the action, so a press applies the entry the index pointed at (C15: the "CW" entry
//     selects audio2100 but parks fndx at 2);
//   * CW picks the sideband from the VFO frequency: above 10 MHz CW_USB, else CW_LSB (C16);
//   * "NOTCH" re-enables the AGC, the DNR entries leave it alone and only set nr_level.
//
// `Radio` is anything with the AudioSDR setters plus reInitializeFilter(lo, hi) and set_nr_level(level): the adapter
// below binds it to (rdsp::Bank, channel); the unit test binds it to a recorder.
#pragma once
#include <cstdint>
#include <string>

namespace rdsp {

template <class Radio>
class SketchControls {
public:
    explicit SketchControls(Radio &r) : radio(r) {}

    // RDSP_general_includes.h:68-82,103-111
    uint32_t vfoFreq = 7050000, TuningOffset = 0;
    double dFLoCut = 300.0, dFHiCut = 4000.0;
    static constexpr double MIN_LOW = 0.0, MAX_LOW = 700.0, MIN_HI = 800.0, MAX_HI = 4000.0;
    int mndx = 3, fndx = 2, nrndx = 0, andx = 2, nr_level = 0;
    std::string newMode, newFilter, newAgc, newNR;

    // RDSP_controls.h:149-191
    void filterMode()
    {
        static const struct { int filter; const char *label; } tab[5] = {
            {RDSP_FILTER_CW, "500 Hz"}, {RDSP_FILTER_2100, "2.1 kHz"}, {RDSP_FILTER_2700, "2.7 kHz"},
            {RDSP_FILTER_3100, "3.1 kHz"}, {RDSP_FILTER_AM, "3.9 kHz"}};
        if (fndx >= 0 && fndx <= 4) { radio.setAudioFilter(tab[fndx].filter); newFilter = tab[fndx].label; }
        fndx = (fndx == 4) ? 0 : fndx + 1;
    }

    // RDSP_controls.h:196-232
    void setAgc()
    {
        static const struct { int mode; const char *label; } tab[4] = {
            {RDSP_AGC_OFF, "AGC O"}, {RDSP_AGC_FAST, "AGC F"}, {RDSP_AGC_MEDIUM, "AGC M"}, {RDSP_AGC_SLOW, "AGC S"}};
        if (andx >= 0 && andx <= 3) { radio.setAGCmode(tab[andx].mode); newAgc = tab[andx].label; }
        andx = (andx == 3) ? 0 : andx + 1;
    }

    // RDSP_controls.h:237-297 — here the index moves FIRST
    void setNRMode()
    {
        nrndx = (nrndx == 5) ? 0 : nrndx + 1;
        switch (nrndx) {
        case 0: radio.disableALSfilter(); radio.enableAGC(); newNR = ""; nr_level = 0; break;
        case 1: radio.enableAGC(); radio.enableALSfilter(); radio.setALSfilterNotch(); radio.setALSfilterAdaptive(); newNR = "NOTCH"; break;
        case 2: radio.disableALSfilter(); newNR = "DNR 1"; nr_level = 20; break;
        case 3: radio.disableALSfilter(); newNR = "DNR 2"; nr_level = 30; break;
        case 4: radio.disableALSfilter(); newNR = "DNR 3"; nr_level = 40; break;
        case 5: radio.disableALSfilter(); newNR = "DNR 4"; nr_level = 50; break;
        }
        radio.set_nr_level(nr_level);          // the sketch's DSP reads the global at the next block, RDSP_convolutional.h:326-330
    }

    // RDSP_controls.h:330-423
    void tuningMode()
    {
        const int cw = vfoFreq > 10000000u ? RDSP_DEMOD_CW_USB : RDSP_DEMOD_CW_LSB;
        switch (mndx) {
        case 0: newMode = "CW N"; radio.setAudioFilter(RDSP_FILTER_CW);   TuningOffset = radio.setDemodMode(cw); newFilter = "500 Hz"; break;
        case 1: newMode = "CW";   radio.setAudioFilter(RDSP_FILTER_2100); TuningOffset = radio.setDemodMode(cw); newFilter = "2.1 kHz"; fndx = 2; break;
        case 2: newMode = "USB";  radio.setAudioFilter(RDSP_FILTER_2700); TuningOffset = radio.setDemodMode(RDSP_DEMOD_USB); newFilter = "2.7 kHz"; fndx = 2; break;
        case 3: newMode = "LSB";  radio.setAudioFilter(RDSP_FILTER_2700); TuningOffset = radio.setDemodMode(RDSP_DEMOD_LSB); newFilter = "2.7 kHz"; fndx = 2; break;
        case 4: newMode = "AM";   radio.setAudioFilter(RDSP_FILTER_AM);   TuningOffset = radio.setDemodMode(RDSP_DEMOD_AM); newFilter = "3.9 kHz"; fndx = 4; break;
        case 5: newMode = "SAM";  radio.setAudioFilter(RDSP_FILTER_AM);   TuningOffset = radio.setDemodMode(RDSP_DEMOD_SAM); newFilter = "3.9 kHz"; fndx = 4; break;
        case 6: newMode = "RTTY"; radio.setAudioFilter(RDSP_FILTER_2100); TuningOffset = radio.setDemodMode(RDSP_DEMOD_USB); newFilter = "2.1 kHz"; fndx = 1; break;
        }
        mndx = (mndx == 6) ? 0 : mndx + 1;
    }

    // RDSP_controls.h:569-612; d3 / d6 = the two buttons held (low-cut / high-cut)
    bool checkPBT_Increase(bool d3, bool d6)
    {
        if (d3) dFLoCut = (dFLoCut + 50) <= MAX_LOW ? dFLoCut + 50 : dFLoCut;
        else if (d6) dFHiCut = (dFHiCut + 50) <= MAX_HI ? dFHiCut + 50 : dFHiCut;
        else return false;
        radio.reInitializeFilter(dFLoCut, dFHiCut);
        return true;
    }
    bool checkPBT_Decrease(bool d3, bool d6)
    {
        if (d3) { dFLoCut = (dFLoCut - 50) > MIN_LOW ? dFLoCut - 50 : dFLoCut; if (dFLoCut < 0.0) dFLoCut = 0.0; }
        else if (d6) dFHiCut = (dFHiCut - 50) > MIN_HI ? dFHiCut - 50 : dFHiCut;
        else return false;
        radio.reInitializeFilter(dFLoCut, dFHiCut);
        return true;
    }

private:
    Radio &radio;
};

#ifdef RDSP_GPU_H_INCLUDED
// binds the controls of one receiver to (bank, channel)
class ChannelRadio {
public:
    ChannelRadio(class Bank &b, uint32_t ch);
    uint32_t setDemodMode(int m);
    void setAudioFilter(int f);
    void setAGCmode(int m);
    void enableAGC();
    void enableALSfilter();
    void disableALSfilter();
    void setALSfilterNotch() {}
    void setALSfilterAdaptive() {}
    void reInitializeFilter(double lo, double hi);
    void set_nr_level(int level);
private:
    class Bank &bank_;
    uint32_t ch_;
    bool agc_on_ = true;
    int agc_mode_ = RDSP_AGC_MEDIUM;
};
#endif

}  // namespace rdsp
