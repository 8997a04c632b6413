"""Host-side mirror of the reference sketch's control vocabulary for one channel of a ReceiverBank.

Same names and argument meaning as the calls the sketch makes on its `AudioSDR SDR;` object and on the
convolution block (RadioDSP_SDR_RX.ino:117-139,183; RDSP_controls.h:149-297,330-423,569-612;
RDSP_convolutional.h:209), so that control code written against the reference reads the same here.
Every setter forwards to rdsp_gpu_set_mode for this channel; it takes effect at the next block.
"""
from __future__ import annotations

from . import native as N

# the reference's enumerators
LSBmode, USBmode, CW_LSBmode, CW_USBmode, AMmode = N.DEMOD_LSB, N.DEMOD_USB, N.DEMOD_CW_LSB, N.DEMOD_CW_USB, N.DEMOD_AM
audioCW, audio2100, audio2700, audio3100, audioAM = N.FILTER_CW, N.FILTER_2100, N.FILTER_2700, N.FILTER_3100, N.FILTER_AM
AGCoff, AGCfast, AGCmedium, AGCslow = N.AGC_OFF, N.AGC_FAST, N.AGC_MEDIUM, N.AGC_SLOW

# PBT limits, RDSP_general_includes.h:76-82
MIN_LOW, MAX_LOW, MIN_HI, MAX_HI = 0.0, 700.0, 800.0, 4000.0


class SDRChannel:
    def __init__(self, bank: "N.ReceiverBank", ch: int):
        self.bank, self.ch = bank, ch
        self._agc_enabled = True
        self._agc_mode = N.AGC_MEDIUM

    def _update(self, **kw):
        p = self.bank.get_mode(self.ch)
        for k, v in kw.items():
            setattr(p, k, v)
        self.bank.set_mode(self.ch, 1, p)

    # --- AudioSDR API used by the sketch ---
    def setDemodMode(self, mode: int) -> int:
        """Returns the tuning offset (IF) in Hz: this zero-IF build returns 0 (the sketch subtracts it
        from the VFO frequency, RDSP_controls.h:445-448)."""
        self._update(demod=mode)
        return 0

    def setAudioFilter(self, f: int):
        self._update(audio_filter=f)

    def enableAGC(self):
        self._agc_enabled = True
        self._update(agc_mode=self._agc_mode)

    def disableAGC(self):
        self._agc_enabled = False
        self._update(agc_mode=N.AGC_OFF)

    def setAGCmode(self, mode: int):
        self._agc_mode = mode
        if self._agc_enabled:
            self._update(agc_mode=mode)

    def enableALSfilter(self):
        self._update(notch_on=1)

    def disableALSfilter(self):
        self._update(notch_on=0)

    def setALSfilterNotch(self):
        self._update(als_peak=0)

    def setALSfilterPeak(self):
        self._update(als_peak=1)

    def enableNoiseBlanker(self):            # RadioDSP_SDR_RX.ino:129
        self._update(nb_on=1)

    def disableNoiseBlanker(self):           # :131
        self._update(nb_on=0)

    def setNoiseBlankerThresholdDb(self, db: float):   # :130
        self._update(nb_threshold_db=db)

    def setALSfilterAdaptive(self):
        pass

    def setInputGain(self, g: float):
        self._update(in_gain=g)

    def setOutputGain(self, g: float):
        self._update(out_gain=g)

    def setIQgainBalance(self, b: float):
        self._update(iq_balance=b)

    # --- convolution block / DNR ---
    def reInitializeFilter(self, lo: float, hi: float):
        """RDSP_convolutional.h:209-224"""
        self._update(pbt_lo_hz=lo, pbt_hi_hz=hi)

    def set_nr_level(self, nr_level: int):
        """global `nr_level` of the sketch: 0, 20, 30, 40, 50 (RDSP_controls.h:265-294)"""
        self._update(nr_kind=N.NR_LMS if nr_level > 0 else N.NR_OFF, nr_level=nr_level)

    def set_spectral_nr(self, level: int):
        """iNRLevel of the backup sketch's spectral subtraction, 0..3"""
        self._update(nr_kind=N.NR_SPECTRAL if level > 0 else N.NR_OFF, nr_level=level)
