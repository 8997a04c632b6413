// test_controls.cpp — the control-plane port (host/rdsp_controls.hpp) against the tables of RDSP_controls.h, on a
// recording radio (no GPU).  Expected sequences are written out from the reference: file:line in the comments.
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/rdsp_gpu.h"
#include "../../radiodsp_sdr_rx_b200/host/rdsp_controls.hpp"

struct Recorder {
    std::vector<std::string> log;
    uint32_t setDemodMode(int m) { log.push_back("demod" + std::to_string(m)); return 0; }
    void setAudioFilter(int f) { log.push_back("filter" + std::to_string(f)); }
    void setAGCmode(int m) { log.push_back("agc" + std::to_string(m)); }
    void enableAGC() { log.push_back("agcOn"); }
    void enableALSfilter() { log.push_back("alsOn"); }
    void disableALSfilter() { log.push_back("alsOff"); }
    void setALSfilterNotch() { log.push_back("alsNotch"); }
    void setALSfilterAdaptive() { log.push_back("alsAdaptive"); }
    void reInitializeFilter(double lo, double hi) { char b[64]; snprintf(b, sizeof b, "pbt%.0f-%.0f", lo, hi); log.push_back(b); }
    void set_nr_level(int l) { log.push_back("nr" + std::to_string(l)); }
};

static int fails = 0;
#define EXPECT(c) do { if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); fails++; } } while (0)

int main()
{
    Recorder r;
    rdsp::SketchControls<Recorder> ui(r);
    // setup() calls tuningMode() once with mndx = 3 (RadioDSP_SDR_RX.ino:109, RDSP_general_includes.h:104): LSB + audio2700
    ui.tuningMode();
    EXPECT(ui.newMode == "LSB" && ui.newFilter == "2.7 kHz" && ui.mndx == 4 && ui.fndx == 2);
    EXPECT((r.log == std::vector<std::string>{"filter2", "demod0"}));
    // next presses: AM, SAM (-> AM here), RTTY, CW N, CW, USB   (RDSP_controls.h:376-413, 332-365)
    const char *modes[] = {"AM", "SAM", "RTTY", "CW N", "CW", "USB", "LSB"};
    const int fnd[] = {4, 4, 1, 1, 2, 2, 2};          // fndx after the press: CW N leaves it alone (still 1 from RTTY), CW parks it at 2 (C15)
    for (int i = 0; i < 7; i++) { ui.tuningMode(); EXPECT(ui.newMode == modes[i]); EXPECT(ui.fndx == fnd[i]); }
    EXPECT(ui.mndx == 4);
    // CW sideband follows the band (RDSP_controls.h:336-340)
    r.log.clear(); ui.mndx = 0; ui.vfoFreq = 14000000; ui.tuningMode();
    EXPECT((r.log == std::vector<std::string>{"filter0", "demod3"}));
    r.log.clear(); ui.mndx = 0; ui.vfoFreq = 7000000; ui.tuningMode();
    EXPECT((r.log == std::vector<std::string>{"filter0", "demod2"}));
    // filterMode(): applies entry fndx, then advances (RDSP_controls.h:149-191)
    r.log.clear(); ui.fndx = 2;
    const char *fl[] = {"2.7 kHz", "3.1 kHz", "3.9 kHz", "500 Hz", "2.1 kHz"};
    for (int i = 0; i < 5; i++) { ui.filterMode(); EXPECT(ui.newFilter == fl[i]); }
    EXPECT(ui.fndx == 2 && (r.log == std::vector<std::string>{"filter2", "filter3", "filter4", "filter0", "filter1"}));
    // setAgc(): starts at andx = 2 (medium) (RDSP_controls.h:196-232)
    r.log.clear();
    const char *al[] = {"AGC M", "AGC S", "AGC O", "AGC F"};
    for (int i = 0; i < 4; i++) { ui.setAgc(); EXPECT(ui.newAgc == al[i]); }
    EXPECT((r.log == std::vector<std::string>{"agc2", "agc3", "agc0", "agc1"}));
    // setNRMode(): index first, then NOTCH, DNR 1..4, off (RDSP_controls.h:237-297)
    r.log.clear();
    const char *nl[] = {"NOTCH", "DNR 1", "DNR 2", "DNR 3", "DNR 4", ""};
    const int lv[] = {0, 20, 30, 40, 50, 0};
    for (int i = 0; i < 6; i++) { ui.setNRMode(); EXPECT(ui.newNR == nl[i]); EXPECT(ui.nr_level == lv[i]); }
    EXPECT(r.log.front() == "agcOn" && r.log[1] == "alsOn" && r.log[4] == "nr0" && r.log[5] == "alsOff" && r.log[6] == "nr20");
    EXPECT(r.log.back() == "nr0" && r.log[r.log.size() - 3] == "alsOff" && r.log[r.log.size() - 2] == "agcOn");
    // PBT stepping and its limits (RDSP_controls.h:569-612, RDSP_general_includes.h:76-82)
    r.log.clear();
    for (int i = 0; i < 10; i++) ui.checkPBT_Increase(true, false);
    EXPECT(ui.dFLoCut == 700.0);                                   // 300 -> 700, clamps at MAX_LOW
    for (int i = 0; i < 20; i++) ui.checkPBT_Decrease(true, false);
    EXPECT(ui.dFLoCut == 50.0);                                    // (lo - 50) > 0 fails at 50: the sketch never reaches 0
    ui.checkPBT_Increase(false, true);
    EXPECT(ui.dFHiCut == 4000.0);                                  // already at MAX_HI
    for (int i = 0; i < 100; i++) ui.checkPBT_Decrease(false, true);
    EXPECT(ui.dFHiCut == 850.0);                                   // (hi - 50) > 800 fails at 850
    EXPECT(!ui.checkPBT_Increase(false, false));
    EXPECT(r.log.front() == "pbt350-4000" && r.log.back() == "pbt50-850");
    printf(fails ? "FAILED (%d)\n" : "ALL OK\n", fails);
    return fails ? 1 : 0;
}
