// sketch_port.cpp — the sketch's setup() / loop() control flow (RadioDSP_SDR_RX.ino:102-233) driving a bank of
// receivers through the C ABI.  Build: see INTEGRATION.md.  Needs a B200 to run; without one it reports the
// library's error and exits 2 (there is no CPU fallback).
#include <cstdio>
#include <cmath>
#include <vector>
#include "../radiodsp_sdr_rx_b200/host/rdsp_sketch_api.hpp"

int main(int argc, char **argv)
{
    const uint32_t N = argc > 1 ? (uint32_t)atoi(argv[1]) : 64;
    try {
        rdsp_gpu_config_t cfg = rdsp::Bank::defaults(N);
        cfg.io_location = RDSP_IO_HOST;
        rdsp::Bank bank(cfg);
        // setup(), RadioDSP_SDR_RX.ino:117-148,183 — for every receiver
        for (uint32_t ch = 0; ch < N; ch++) {
            rdsp::AudioSDR SDR(bank, ch);
            SDR.enableAGC();
            SDR.setAGCmode(rdsp::AGCmedium);
            SDR.disableALSfilter();
            SDR.setInputGain(1.0f);
            SDR.setOutputGain(0.5f);
            SDR.setIQgainBalance(1.020f);
            SDR.setAudioFilter(rdsp::audio2700);
            (void)SDR.setDemodMode(ch % 2 ? rdsp::USBmode : rdsp::LSBmode);
            rdsp::reInitializeFilter(bank, ch, 300, 4000);
            rdsp::set_nr_level(bank, ch, ch % 4 == 0 ? 30 : 0);     // "DNR 2" on every fourth receiver
        }
        // loop(): one tick per 128 samples
        std::vector<int16_t> iq((size_t)N * 256), audio((size_t)N * 256);
        rdsp::AudioAnalyzeFFT256IQ FFT(bank, 0);
        double ph = 0.0;
        for (int tick = 0; tick < 64; tick++) {
            for (int n = 0; n < 128; n++, ph += 2.0 * M_PI * 1000.0 / 44100.0)
                for (uint32_t ch = 0; ch < N; ch++) {
                    const double s = ch % 2 ? 1.0 : -1.0;                  // tone on the channel's own sideband
                    iq[((size_t)ch * 128 + n) * 2] = (int16_t)lrint(8000 * cos(ph));
                    iq[((size_t)ch * 128 + n) * 2 + 1] = (int16_t)lrint(s * 8000 * sin(ph));
                }
            bank.update(iq.data(), audio.data());
            if (FFT.available()) printf("tick %d: spectrum ready, peak bin value %.4f\n", tick, FFT.read(127 - 6));
        }
        double e = 0;
        for (int n = 0; n < 128; n++) e += (double)audio[n * 2] * audio[n * 2];
        printf("receiver 0 audio rms = %.1f LSB over the last block (%u receivers)\n", sqrt(e / 128), N);
    } catch (const rdsp::Error &e) {
        fprintf(stderr, "rdsp error %d: %s\n", e.code, e.what());
        return 2;
    }
    return 0;
}
