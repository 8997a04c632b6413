"""Behavioural known-answer tests of the shim-defined stages of the oracle (AudioSDR / Teensy pieces that are
absent from the reference tree): they pin the sign conventions and the integer semantics."""
import ctypes as C

import numpy as np

from radiodsp_sdr_rx_b200 import synth

FS = 44100.0


def _tone_iq(f_hz, nb, amp=8000):
    n = np.arange(nb * 128)
    z = amp * np.exp(2j * np.pi * f_hz * n / FS)
    return np.stack([np.rint(z.real), np.rint(z.imag)], -1).astype(np.int16).reshape(nb, 128, 2)


def _rms(x):
    return float(np.sqrt(np.mean((x.astype(np.float64)) ** 2)))


def test_sideband_convention(po):
    """a tone below the carrier survives LSB and vanishes in USB, and vice versa (A.4)"""
    cfg = po.default_config(stage_mask=po.STAGE_FRONTEND)
    lower, upper = _tone_iq(-1000.0, 12), _tone_iq(+1000.0, 12)
    r = {}
    for name, dm in (("lsb", po.DEMOD_LSB), ("usb", po.DEMOD_USB)):
        for sname, sig in (("lower", lower), ("upper", upper)):
            out = po.OracleChan(cfg, po.default_params(demod=dm, iq_balance=1.0)).process(sig)
            r[name, sname] = _rms(out[4:, :, 0])
    assert r["lsb", "lower"] > 5000 and r["usb", "upper"] > 5000
    assert r["lsb", "upper"] < 0.002 * r["lsb", "lower"] and r["usb", "lower"] < 0.002 * r["usb", "upper"]
    # unity pass-band gain of the pair + band-pass: 8000 peak -> 5657 rms
    assert abs(r["lsb", "lower"] - 8000 / np.sqrt(2)) < 150      # -0.15 dB of filter skirt at 1 kHz


def test_am_envelope_and_cw_filter(po):
    cfg = po.default_config(stage_mask=po.STAGE_FRONTEND)
    n = np.arange(40 * 128)
    env = 6000 * (1 + 0.5 * np.cos(2 * np.pi * 1000 * n / FS))
    iq = np.stack([np.rint(env * np.cos(0.3)), np.rint(env * np.sin(0.3))], -1).astype(np.int16).reshape(40, 128, 2)
    out = po.OracleChan(cfg, po.default_params(demod=po.DEMOD_AM, audio_filter=po.FILTER_AM, iq_balance=1.0)).process(iq)
    a = out[6:, :, 0].reshape(-1).astype(float)
    assert synth.tone_snr_db(a - a.mean(), [1000.0]) > 25
    # CW filter passes 700 Hz, rejects 2500 Hz
    cw = po.default_params(demod=po.DEMOD_CW_USB, audio_filter=po.FILTER_CW, iq_balance=1.0)
    p700 = _rms(po.OracleChan(cfg, cw).process(_tone_iq(700.0, 12))[4:, :, 0])
    p2500 = _rms(po.OracleChan(cfg, cw).process(_tone_iq(2500.0, 12))[4:, :, 0])
    assert p700 > 5000 and p2500 < 1e-3 * p700


def test_sam_locks_to_an_offset_carrier_where_the_envelope_detector_cannot_help(po):
    """SAMmode (RDSP_controls.h:384-391): the carrier loop pulls in a carrier that is 40 Hz off tune and detects
    coherently; the 1 kHz modulation comes out clean, and the loop frequency settles at the offset."""
    cfg = po.default_config(stage_mask=po.STAGE_FRONTEND)
    nb = 400                                                 # the carrier-level tracker (100 ms) needs ~1 s to settle
    n = np.arange(nb * 128)
    env = 6000 * (1 + 0.5 * np.cos(2 * np.pi * 1000 * n / FS))
    car = np.exp(1j * (2 * np.pi * 40.0 * n / FS + 0.7))
    iq = np.stack([np.rint(env * car.real), np.rint(env * car.imag)], -1).astype(np.int16).reshape(nb, 128, 2)
    ch = po.OracleChan(cfg, po.default_params(demod=po.DEMOD_SAM, audio_filter=po.FILTER_AM, iq_balance=1.0))
    out = ch.process(iq)
    a = out[340:, :, 0].reshape(-1).astype(float)
    assert synth.tone_snr_db(a - a.mean(), [1000.0]) > 40
    assert abs(a.std() - 0.5 * 6000 / np.sqrt(2)) < 60        # the modulation through unity-gain filters
    # a carrier-only input detects to (almost) nothing once the level tracker has settled
    iq0 = np.stack([np.rint(6000 * car.real), np.rint(6000 * car.imag)], -1).astype(np.int16).reshape(nb, 128, 2)
    quiet = po.OracleChan(cfg, po.default_params(demod=po.DEMOD_SAM, audio_filter=po.FILTER_AM, iq_balance=1.0)).process(iq0)
    assert _rms(quiet[340:, :, 0]) < 30


def test_als_peak_emits_the_estimate(po):
    """ALS "peak": the notch stage emits what the NLMS predicts (the steady heterodyne) instead of the residue"""
    cfg = po.default_config(stage_mask=po.STAGE_FRONTEND | po.STAGE_NOTCH)
    sig = _tone_iq(-1500.0, 60, amp=6000)
    off = po.OracleChan(cfg, po.default_params(notch_on=0, iq_balance=1.0)).process(sig)
    notch = po.OracleChan(cfg, po.default_params(notch_on=1, iq_balance=1.0)).process(sig)
    peak = po.OracleChan(cfg, po.default_params(notch_on=1, als_peak=1, iq_balance=1.0)).process(sig)
    assert _rms(notch[40:, :, 0]) < 0.05 * _rms(off[40:, :, 0])
    assert abs(_rms(peak[40:, :, 0]) / _rms(off[40:, :, 0]) - 1.0) < 0.05


def test_noise_blanker_zeroes_impulses_and_leaves_the_signal(po):
    """SDR.enableNoiseBlanker + setNoiseBlankerThresholdDb (RadioDSP_SDR_RX.ino:129-130): frames far above the running
    IQ magnitude are zeroed before the filters; a clean signal passes untouched"""
    cfg = po.default_config(stage_mask=po.STAGE_FRONTEND)
    clean = _tone_iq(-1000.0, 40, amp=2000)
    dirty = clean.copy()
    dirty.reshape(-1, 2)[777::1500] = 30000                     # ignition-noise style impulses, ~24 dB above the tone
    off = po.default_params(iq_balance=1.0)
    on = po.default_params(iq_balance=1.0, nb_on=1, nb_threshold_db=12.0)
    ref = po.OracleChan(cfg, off).process(clean)
    assert np.array_equal(po.OracleChan(cfg, on).process(clean), ref)          # nothing to blank
    hurt = po.OracleChan(cfg, off).process(dirty)
    healed = po.OracleChan(cfg, on).process(dirty)
    e_hurt = _rms((hurt - ref.astype(np.int32))[4:, :, 0]); e_healed = _rms((healed - ref.astype(np.int32))[4:, :, 0])
    assert e_healed < 0.12 * e_hurt


def test_q15_fir_wraps_and_saturates(po):
    """arm_fir_fast_q15 convention: 32-bit wrap-around accumulator, then SSAT(acc >> 15)"""
    L = po.lib()
    taps = np.full(129, 32767, np.int16)
    hist = np.full(128, 32767, np.int16)
    x = np.full(16, 32767, np.int16)
    y = np.zeros(16, np.int16)
    L.oracle_fir_q15(taps.ctypes.data, 129, hist.ctypes.data, x.ctypes.data, y.ctypes.data, 16)
    acc = (129 * 32767 * 32767) & 0xFFFFFFFF
    acc = acc - (1 << 32) if acc & 0x80000000 else acc
    want = max(-32768, min(32767, acc >> 15))
    assert (y == want).all()
    taps[:] = 0
    taps[0] = 16384                    # 0.5
    L.oracle_fir_q15(taps.ctypes.data, 129, hist.ctypes.data, x.ctypes.data, y.ctypes.data, 16)
    assert (y == (16384 * 32767) >> 15).all()


def test_float_to_q15_truncates_and_saturates(po):
    src = np.array([0.99999, -0.99999, 1.5, -1.5, 0.5 / 32768, -0.5 / 32768, 1.9 / 32768, -1.9 / 32768, np.nan], np.float32)
    dst = np.zeros(src.size, np.int16)
    po.lib().arm_float_to_q15(src.ctypes.data, dst.ctypes.data, src.size)
    assert list(dst) == [32767, -32767, 32767, -32768, 0, 0, 1, -1, 0]


def test_biquad_removes_dc(po):
    bq = (C.c_int32 * 8)()
    po.lib().oracle_biquad_set_highpass(bq, 500.0, 0.5)
    assert bq[0] == bq[2] and bq[1] == -2 * bq[0] or abs(bq[1] + 2 * bq[0]) <= 2
    blk = np.full(128, 10000, np.int16)
    last = None
    for _ in range(20):
        b = blk.copy()
        po.lib().oracle_biquad_update(bq, b.ctypes.data)
        last = b
    assert np.abs(last).max() <= 2


def test_nlms_zero_in_zero_out_and_notch(po):
    cfg = po.default_config(stage_mask=po.STAGE_FFTFILT | po.STAGE_NR)
    ch = po.OracleChan(cfg, po.default_params(nr_kind=po.NR_LMS, nr_level=30))
    assert not ch.process(np.zeros((5, 128, 2), np.int16)).any()
    # auto-notch: a steady heterodyne is predictable => removed from the error output
    cfg = po.default_config(stage_mask=po.STAGE_FRONTEND | po.STAGE_NOTCH)
    sig = _tone_iq(-1500.0, 60, amp=6000)
    on = po.OracleChan(cfg, po.default_params(notch_on=1, iq_balance=1.0)).process(sig)
    off = po.OracleChan(cfg, po.default_params(notch_on=0, iq_balance=1.0)).process(sig)
    assert _rms(on[40:, :, 0]) < 0.05 * _rms(off[40:, :, 0])


def test_agc_levels_and_off(po):
    cfg = po.default_config(stage_mask=po.STAGE_FRONTEND | po.STAGE_AGC)
    weak, strong = _tone_iq(-1000.0, 400, amp=300), _tone_iq(-1000.0, 400, amp=12000)
    pk = {}
    for name, sig in (("weak", weak), ("strong", strong)):
        out = po.OracleChan(cfg, po.default_params(agc_mode=po.AGC_FAST, iq_balance=1.0)).process(sig)
        pk[name] = np.abs(out[300:, :, 0]).max() / 32768.0
    # target 0.25 * output gain 0.5
    assert abs(pk["weak"] - 0.125) < 0.02 and abs(pk["strong"] - 0.125) < 0.02
    off = po.OracleChan(cfg, po.default_params(agc_mode=po.AGC_OFF, iq_balance=1.0)).process(strong)
    assert abs(np.abs(off[10:, :, 0]).max() - 6000) < 150     # two filter skirts at 1 kHz


def test_fft1024_tone_bin(po):
    cfg = po.default_config(stage_mask=po.STAGE_FFTFILT | po.STAGE_SPEC1024)
    n = np.arange(16 * 128)
    f = 43 * FS / 1024                                   # exactly bin 43
    z = 9000 * np.exp(2j * np.pi * f * n / FS)           # analytic: passes the one-sided PBT filter
    iq = np.stack([np.rint(z.real), np.rint(z.imag)], -1).astype(np.int16).reshape(16, 128, 2)
    ch = po.OracleChan(cfg)
    ready_at = []
    for k in range(16):
        ch.process(iq[k:k + 1])
        spec, ready = ch.read_audio_spectrum()
        if ready:
            ready_at.append(k)
            last = spec
    assert ready_at == [7, 11, 15]                       # 8 blocks, then every 4 (50 % overlap)
    assert int(np.argmax(last)) == 43


def test_spectral_nr_reduces_noise_floor(po):
    cfg = po.default_config(stage_mask=po.STAGE_FFTFILT | po.STAGE_NR)
    rng = np.random.default_rng(5)
    noise = rng.normal(0, 300, (60, 128, 2))
    sig = _tone_iq(1000.0, 60, amp=8000).astype(float)
    iq = np.clip(np.rint(sig + noise), -32768, 32767).astype(np.int16)
    on = po.OracleChan(cfg, po.default_params(nr_kind=po.NR_SPECTRAL, nr_level=2)).process(iq)
    assert synth.tone_snr_db(on[20:, :, 0].reshape(-1), [1000.0]) > 20
    quiet = np.clip(np.rint(noise), -32768, 32767).astype(np.int16)
    a = po.OracleChan(cfg, po.default_params(nr_kind=po.NR_SPECTRAL, nr_level=3)).process(quiet)
    assert _rms(a[20:]) < 0.6 * _rms(quiet[20:])


def test_panadapter_postprocessing(po):
    cfg = po.default_config(stage_mask=po.STAGE_SPEC256, spec256_naverage=2)
    ch = po.OracleChan(cfg)
    ch.process(_tone_iq(8 * FS / 256, 5, amp=9000))
    trace, sm = ch.read_panadapter()
    assert trace.dtype == np.uint16 and int(np.argmax(trace)) in (118, 119, 120)
    t2, _ = ch.read_panadapter()                          # time smoothing converges upwards
    assert t2[119] >= trace[119]


def test_waterfall_history(po):
    """RDSP_display.h:282-319: each refresh pushes SpectrumView[2x] on top, older lines move down one row"""
    cfg = po.default_config(stage_mask=po.STAGE_SPEC256, spec256_naverage=1)
    ch = po.OracleChan(cfg)
    lines = []
    for k in range(4):
        ch.process(_tone_iq((8 + 6 * k) * FS / 256, 3, amp=9000))
        trace, _ = ch.read_panadapter()
        lines.append(trace[0::2].copy())
    rows, col = ch.read_waterfall()
    for r in range(4):
        assert np.array_equal(rows[r], lines[3 - r])
    assert not rows[4:].any()
    assert col.max() <= 6 and ((rows >= 75) == (col == 6)).all() and ((rows < 5) == (col == 0)).all()


def test_dnr_hook_equals_the_chain(po):
    """rdsp_oracle_chan_dnr_f32 (K6 alone, used by the GPU tests to hand both sides identical inputs) is the DNR branch of
    the chain: K5-only output pushed through it == the K5+K6 chain, bit for bit, through a level change"""
    from radiodsp_sdr_rx_b200 import synth
    iq = synth.synth_iq([3], 24, [0])[:, 0]
    k5 = po.OracleChan(po.default_config(stage_mask=po.STAGE_FFTFILT))
    _, f5 = k5.process(iq, True)
    full = po.OracleChan(po.default_config(stage_mask=po.STAGE_FFTFILT | po.STAGE_NR))
    hook = po.OracleChan(po.default_config(stage_mask=po.STAGE_FFTFILT | po.STAGE_NR))
    for b0, lvl in ((0, 30), (12, 50)):
        p = po.default_params(nr_kind=po.NR_LMS, nr_level=lvl)
        full.set_mode(p); hook.set_mode(p)
        _, f6 = full.process(iq[b0:b0 + 12], True)
        y = hook.dnr_f32(f5[b0:b0 + 12, :, 0])
        assert np.array_equal(y, f6[..., 0]) and np.array_equal(f6[..., 0], f6[..., 1])
