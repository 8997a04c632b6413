"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): integer / q15 stages and all indexing bit-exact; float32 stages within 1e-4
relative RMS on the pre-quantisation signal (and q15 outputs within 1 LSB where an f32 value straddles a
truncation boundary); demodulated-audio SNR equal within 0.1 dB.
"""
import os

import numpy as np
import pytest

import bench
from conftest import GOLDEN
from parity_util import REL_RMS_TOL, SNR_TOL_DB, check_chain, rel_rms
from radiodsp_sdr_rx_b200 import synth

pytestmark = pytest.mark.gpu


def make_bank(rd, n_channels, stage_mask, max_blocks=64, **kw):
    cfg = rd.default_config(n_channels=n_channels, stage_mask=stage_mask, max_blocks_per_call=max_blocks,
                            io_location=rd.IO_HOST, debug_f32=1, **kw)
    return rd.ReceiverBank(cfg)


def to_rd_params(rd, p):
    """pyoracle.Params -> radiodsp Params (identical layout)"""
    return rd.Params.from_buffer_copy(p)


def run_both(rd, po, stage_mask, params, iq, blocks_per_call=None, **cfgkw):
    """params: list of pyoracle Params per channel.  Returns (gpu_out, gpu_f32, ora_out, ora_f32, bank, chans)."""
    nb, nc = iq.shape[:2]
    step = blocks_per_call or nb
    bank = make_bank(rd, nc, stage_mask, max_blocks=step, **cfgkw)
    for c, p in enumerate(params):
        bank.set_mode(c, 1, to_rd_params(rd, p))
    outs, f32s = [], []
    for b0 in range(0, nb, step):
        chunk = np.ascontiguousarray(iq[b0:b0 + step])
        outs.append(bank.process_host(chunk))
        f32s.append(bank.read_debug_f32(chunk.shape[0]))
    g_out, g_f32 = np.concatenate(outs), np.concatenate(f32s)
    ocfg = po.default_config(stage_mask=stage_mask, **{k: v for k, v in cfgkw.items() if k == "spec256_naverage"})
    o_out, o_f32, chans = po.process_bank(ocfg, list(params), iq, want_f32=True)
    return g_out, g_f32, o_out, o_f32, bank, chans


# ------------------------------------------------------------------------------------------ K0-K2

@pytest.mark.parametrize("blocks_per_call", [1, 5, 20])
def test_frontend_bit_exact_all_modes(rd, po, blocks_per_call):
    combos = [(d, f) for d in range(5) for f in range(5)]
    nc = len(combos) + 2
    demod = [d for d, _ in combos] + [0, 1]
    iq = synth.synth_iq(np.arange(nc), 20, demod, interferer=[c % 3 == 0 for c in range(nc)])
    iq[:, -1] = np.clip(iq[:, -1].astype(np.int32) * 9, -32768, 32767)          # drive the saturating paths
    params = [po.default_params(demod=d, audio_filter=f) for d, f in combos]
    params += [po.default_params(demod=0, in_gain=2.5, iq_balance=0.9), po.default_params(demod=1, iq_balance=1.0)]
    g_out, _, o_out, _, _, _ = run_both(rd, po, rd.STAGE_FRONTEND, params, iq, blocks_per_call)
    assert np.array_equal(g_out, o_out)
    assert np.array_equal(g_out[..., 0], g_out[..., 1])


def test_taps_are_data(rd, po):
    """rdsp_gpu_set_taps: an arbitrary q15 table gives the oracle's result for the same table"""
    rng = np.random.default_rng(3)
    t_i = rng.integers(-3000, 3000, 129).astype(np.int16)
    t_q = rng.integers(-3000, 3000, 129).astype(np.int16)
    t_b = rng.integers(-2000, 2000, 129).astype(np.int16)
    iq = synth.synth_iq([5, 6], 6, [1, 1])
    bank = make_bank(rd, 2, rd.STAGE_FRONTEND)
    assert np.array_equal(bank.get_taps(rd.TAPS_BANDPASS, rd.FILTER_2700), po.get_taps(po.TAPS_BANDPASS, po.FILTER_2700))
    bank.set_mode(0, 2, rd.default_params(demod=rd.DEMOD_USB))
    bank.set_taps(rd.TAPS_HILBERT_I, rd.DEMOD_USB, t_i)
    bank.set_taps(rd.TAPS_HILBERT_Q, rd.DEMOD_USB, t_q)
    bank.set_taps(rd.TAPS_BANDPASS, rd.FILTER_2700, t_b)
    g = bank.process_host(iq)
    saved = [po.get_taps(k, i) for k, i in ((0, 1), (1, 1), (2, 2))]
    try:
        po.lib().rdsp_oracle_set_taps(0, 1, t_i.ctypes.data); po.lib().rdsp_oracle_set_taps(1, 1, t_q.ctypes.data)
        po.lib().rdsp_oracle_set_taps(2, 2, t_b.ctypes.data)
        o, _ = po.process_bank(po.default_config(stage_mask=po.STAGE_FRONTEND), po.default_params(demod=po.DEMOD_USB), iq)
    finally:
        for (k, i), t in zip(((0, 1), (1, 1), (2, 2)), saved):
            po.lib().rdsp_oracle_set_taps(k, i, t.ctypes.data)
    assert np.array_equal(g, o)


@pytest.mark.parametrize("nc,nblocks", [(300, 5), (1200, 12)])
def test_frontend_tensor_core_wraps_like_the_cuda_core_kernel(rd, po, nc, nblocks, monkeypatch):
    """k_front_tc (tcgen05, byte-split q15 products) against k_front (IMAD) and the oracle with FULL-RANGE taps and
    inputs, where the 32-bit fast-FIR accumulator wraps: bit-exact outputs over several calls (state carry),
    ragged tiles (channel classes that do not fill 128 rows) and calls long enough to be cut into time segments."""
    rng = np.random.default_rng(11)
    taps = {(k, i): rng.integers(-32768, 32768, 129).astype(np.int16) for k in range(3) for i in range(5)}
    taps[(0, 1)] = taps[(0, 0)].copy(); taps[(1, 1)] = taps[(1, 0)].copy()       # LSB / USB share their Hilbert rows
    for i in range(5):
        taps[(2, i)] = (taps[(2, i)] // (1 << i)).astype(np.int16)                # a spread of band-pass magnitudes
    demod = rng.integers(0, 5, nc)
    filt = rng.integers(0, 5, nc)
    gains = rng.choice([1.0, 0.37, 2.5, 1.02], nc)
    iq = rng.integers(-32768, 32768, (nblocks, nc, 128, 2)).astype(np.int16)
    iq[:, ::7] = np.where(rng.random((nblocks, (nc + 6) // 7, 128, 2)) < 0.5, 32767, -32768).astype(np.int16)

    def run(impl):
        if impl:
            monkeypatch.setenv("RDSP_FRONT_IMPL", impl)
        else:
            monkeypatch.delenv("RDSP_FRONT_IMPL", raising=False)
        bank = make_bank(rd, nc, rd.STAGE_FRONTEND, max_blocks=nblocks)
        for (k, i), t in taps.items():
            bank.set_taps(k, i, t)
        for c in range(nc):
            bank.set_mode(c, 1, rd.default_params(demod=int(demod[c]), audio_filter=int(filt[c]), in_gain=float(gains[c])))
        outs = [bank.process_host(iq)]
        outs.append(bank.process_host(iq[: max(1, nblocks // 2)]))              # second call: delay lines carried over
        outs.append(bank.process_host(iq[:1]))
        return np.concatenate(outs)

    g_tc = run(None)
    g_cc = run("cuda-core")
    assert np.array_equal(g_tc, g_cc)
    saved = {(k, i): po.get_taps(k, i) for (k, i) in taps}
    try:
        for (k, i), t in taps.items():
            po.lib().rdsp_oracle_set_taps(k, i, t.ctypes.data)
        sub = np.arange(0, nc, max(1, nc // 40))                                  # the oracle on a sample of the channels
        params = [po.default_params(demod=int(demod[c]), audio_filter=int(filt[c]), in_gain=float(gains[c])) for c in sub]
        seq = np.concatenate([iq[:, sub], iq[: max(1, nblocks // 2), sub], iq[:1, sub]])
        o, _ = po.process_bank(po.default_config(stage_mask=po.STAGE_FRONTEND), params, np.ascontiguousarray(seq))
    finally:
        for (k, i), t in saved.items():
            po.lib().rdsp_oracle_set_taps(k, i, t.ctypes.data)
    assert np.array_equal(g_tc[:, sub], o)


def test_sam_matches_the_oracle(rd, po):
    """SAMmode: carrier PLL + coherent detection inside k_front_tc's epilogue (sequential over the samples of a row),
    f32 loop against the oracle's libm loop: demodulated q15 within 2 LSB, relative RMS <= 1e-4 of full scale signal"""
    nc, nb = 140, 40                                        # 140 SAM channels: one full tile and a ragged one
    n = np.arange(nb * 128)
    rng = np.random.default_rng(5)
    iq = np.zeros((nb, nc, 128, 2), np.int16)
    for c in range(nc):
        off, ph, mod = rng.uniform(-300, 300), rng.uniform(0, 6.28), rng.uniform(300, 2500)
        env = 5000 * (1 + 0.6 * np.cos(2 * np.pi * mod * n / 44100.0))
        z = env * np.exp(1j * (2 * np.pi * off * n / 44100.0 + ph)) + rng.normal(0, 200, n.size) + 1j * rng.normal(0, 200, n.size)
        iq[:, c, :, 0] = np.rint(z.real).reshape(nb, 128); iq[:, c, :, 1] = np.rint(z.imag).reshape(nb, 128)
    params = [po.default_params(demod=po.DEMOD_SAM, audio_filter=po.FILTER_AM) for _ in range(nc)]
    params[3] = po.default_params(demod=po.DEMOD_AM, audio_filter=po.FILTER_AM)          # a neighbour in another class
    g_out, _, o_out, _, _, _ = run_both(rd, po, rd.STAGE_FRONTEND, params, iq, blocks_per_call=8)
    assert np.array_equal(g_out[:, 3], o_out[:, 3])                                         # the AM channel stays bit-exact
    d = g_out.astype(np.int32) - o_out
    assert np.abs(d).max() <= 2, np.abs(d).max()
    assert rel_rms(g_out, o_out) <= REL_RMS_TOL
    assert abs(synth.snr_db(o_out, g_out)) > 60.0


def test_als_peak_matches_the_oracle(rd, po):
    nc = 12
    iq = synth.synth_iq(np.arange(nc), 30, 3, interferer=True)
    params = [po.default_params(demod=po.DEMOD_CW_USB, audio_filter=po.FILTER_CW, notch_on=1, als_peak=c % 2) for c in range(nc)]
    g_out, g_f32, o_out, o_f32, _, _ = run_both(rd, po, rd.STAGE_FRONTEND | rd.STAGE_NOTCH | rd.STAGE_AGC, params, iq, blocks_per_call=6)
    assert rel_rms(g_f32, o_f32) <= REL_RMS_TOL
    assert np.abs(g_out.astype(np.int32) - o_out).max() <= 1
    assert not np.array_equal(g_out[:, 0], g_out[:, 1])      # notch and peak differ


def test_wspr_band_through_set_taps(rd, po):
    """audioWSPR (RDSP_controls.h:392-402): rdsp_gpu_design_bandpass(1400, 1600) loaded into a preset row"""
    from radiodsp_sdr_rx_b200 import native
    t = native.design_bandpass(1400.0, 1600.0)
    iq = synth.synth_iq([1, 2, 3], 12, [1, 1, 1])
    bank = make_bank(rd, 3, rd.STAGE_FRONTEND)
    bank.set_mode(0, 3, rd.default_params(demod=rd.DEMOD_USB, audio_filter=rd.FILTER_CW))
    bank.set_taps(rd.TAPS_BANDPASS, rd.FILTER_CW, t)
    g = bank.process_host(iq)
    saved = po.get_taps(2, 0)
    try:
        po.lib().rdsp_oracle_set_taps(2, 0, t.ctypes.data)
        o, _ = po.process_bank(po.default_config(stage_mask=po.STAGE_FRONTEND),
                               po.default_params(demod=po.DEMOD_USB, audio_filter=po.FILTER_CW), iq)
    finally:
        po.lib().rdsp_oracle_set_taps(2, 0, saved.ctypes.data)
    assert np.array_equal(g, o)


def test_noise_blanker_bit_exact(rd, po):
    """the integer noise blanker inside k_front_tc's loader (chunk sums by shared-memory atomics, the running magnitude
    closed per 32-sample chunk) against the oracle: bit-exact, blanked and unblanked channels side by side"""
    nc, nb = 150, 24
    demod = [c % 5 for c in range(nc)]
    iq = synth.synth_iq(np.arange(nc), nb, demod)
    rng = np.random.default_rng(9)
    hits = rng.random(iq.shape[:3]) < 0.004
    iq[hits] = np.where(rng.random((int(hits.sum()), 2)) < 0.5, 32767, -32768).astype(np.int16)
    params = [po.default_params(demod=demod[c], audio_filter=c % 5, nb_on=int(c % 3 != 0), nb_threshold_db=float(6 + 3 * (c % 7)))
              for c in range(nc)]
    g_out, _, o_out, _, _, _ = run_both(rd, po, rd.STAGE_FRONTEND, params, iq, blocks_per_call=7)
    assert np.array_equal(g_out, o_out)


# ------------------------------------------------------------------------------------------ K3, K4

def test_notch_and_agc_f32_parity(rd, po):
    nc = 12
    demod = [0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 0, 1]
    iq = synth.synth_iq(np.arange(100, 100 + nc), 60, demod, interferer=True)
    params = [po.default_params(demod=d, audio_filter=(0 if d in (2, 3) else 2), notch_on=int(c % 2 == 0),
                                notch_level=(20, 30, 15)[c % 3], agc_mode=c % 4) for c, d in enumerate(demod)]
    sm = rd.STAGE_FRONTEND | rd.STAGE_NOTCH | rd.STAGE_AGC
    g_out, g_f32, o_out, o_f32, _, _ = run_both(rd, po, sm, params, iq, blocks_per_call=7)
    for c in range(nc):
        assert rel_rms(g_f32[:, c], o_f32[:, c]) <= REL_RMS_TOL, c
    d = np.abs(g_out.astype(np.int32) - o_out)
    assert d.max() <= 1 and (d != 0).mean() < 5e-3


def test_notch_without_agc_stage(rd, po):
    iq = synth.synth_iq([1, 2, 3], 30, [0, 1, 2], interferer=True)
    params = [po.default_params(demod=d, notch_on=1) for d in (0, 1, 2)]
    g_out, g_f32, o_out, o_f32, _, _ = run_both(rd, po, rd.STAGE_FRONTEND | rd.STAGE_NOTCH, params, iq, 4)
    assert rel_rms(g_f32, o_f32) <= REL_RMS_TOL
    assert np.abs(g_out.astype(np.int32) - o_out).max() <= 1


# ------------------------------------------------------------------------------------------ K5-K8

@pytest.mark.parametrize("name", ["conv_nr0", "conv_nr30", "conv_levels", "conv_kat"])
def test_conv_against_reference_golden(rd, po, name):
    """golden vectors = outputs of the reference's own sources compiled unmodified (tools/make_golden.py)"""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    iq = np.stack([g["in_L"], g["in_R"]], axis=-1)[:, None]                     # [nb,1,128,2]
    nb = iq.shape[0]
    levels = g["nr_level"] if "nr_level" in g else np.zeros(nb, int)
    lo, hi = (g["pbt"] if "pbt" in g else (300.0, 4000.0))
    bank = make_bank(rd, 1, rd.STAGE_FFTFILT | rd.STAGE_NR)
    outs, f32s, cur = [], [], None
    for k in range(nb):
        if levels[k] != cur:
            cur = int(levels[k])
            bank.set_mode(0, 1, rd.default_params(pbt_lo_hz=float(lo), pbt_hi_hz=float(hi),
                                                  nr_kind=rd.NR_LMS if cur > 0 else rd.NR_OFF, nr_level=cur))
        outs.append(bank.process_host(iq[k:k + 1]))
        f32s.append(bank.read_debug_f32(1))
    out, f32 = np.concatenate(outs)[:, 0], np.concatenate(f32s)[:, 0]
    if "f32_L" in g:
        assert rel_rms(f32[..., 0], g["f32_L"]) <= REL_RMS_TOL
        nr_on = np.asarray(levels) > 0
        if (~nr_on).any():
            assert rel_rms(f32[~nr_on][..., 1], g["f32_R"][~nr_on]) <= REL_RMS_TOL
    want = np.stack([g["out_L"], g["out_R"]], axis=-1)
    d = np.abs(out.astype(np.int32) - want)
    assert d.max() <= 1 and (d != 0).mean() < 5e-3
    if "mask" in g:
        assert np.abs(bank.get_mask(0) - g["mask"]).max() < 2e-6


def test_conv_and_nr_kinds_vs_oracle(rd, po):
    nc = 10
    iq = synth.synth_iq(np.arange(300, 300 + nc), 50, [c % 5 for c in range(nc)])
    params = []
    for c in range(nc):
        kind = (po.NR_OFF, po.NR_LMS, po.NR_SPECTRAL, po.NR_LMS, po.NR_SPECTRAL)[c % 5]
        level = (0, 20, 1, 50, 3)[c % 5]
        params.append(po.default_params(nr_kind=kind, nr_level=level, pbt_lo_hz=float(50 * (c % 8)), pbt_hi_hz=float(4000 - 250 * c)))
    g_out, g_f32, o_out, o_f32, _, _ = run_both(rd, po, rd.STAGE_FFTFILT | rd.STAGE_NR, params, iq, blocks_per_call=6)
    # K8 rebuilds every bin through atan2 + arm_cos_f32 / arm_sin_f32 like the sketch; the kernel interpolates the same
    # 512-entry table (k_fftfilt.cu, arm_trig_f32), so the spectral-subtraction channels meet the same bar as the rest
    for c in range(nc):
        assert rel_rms(g_f32[:, c, :, 0], o_f32[:, c, :, 0]) <= REL_RMS_TOL, c
        assert rel_rms(g_f32[:, c, :, 1], o_f32[:, c, :, 1]) <= REL_RMS_TOL or params[c].nr_kind == po.NR_LMS, c
        assert np.abs(g_out[:, c].astype(np.int32) - o_out[:, c]).max() <= 1, c


def test_explicit_mask_is_data(rd, po):
    rng = np.random.default_rng(11)
    mask = rng.normal(0, 0.5, 512).astype(np.float32)
    iq = synth.synth_iq([1], 8, [0])
    bank = make_bank(rd, 1, rd.STAGE_FFTFILT)
    bank.set_mask(0, 1, mask)
    g = bank.process_host(iq)
    gf = bank.read_debug_f32(8)
    ch = po.OracleChan(po.default_config(stage_mask=po.STAGE_FFTFILT))
    ch.set_mask(mask)
    o, of = ch.process(iq[:, 0], True)
    assert rel_rms(gf[:, 0], of) <= REL_RMS_TOL and np.abs(g[:, 0].astype(np.int32) - o).max() <= 1


# ------------------------------------------------------------------------------------------ K9, K10, K11

@pytest.mark.parametrize("nav,blocks_per_call", [(30, 1), (30, 16), (4, 3), (1, 7)])
def test_spec256_bit_exact(rd, po, nav, blocks_per_call):
    nc, nb = 27, 66
    iq = synth.synth_iq(np.arange(nc), nb, [c % 5 for c in range(nc)])
    iq[:, 3] = np.where((np.arange(nb * 128) // 5) % 2 == 0, 32767, -32768).astype(np.int16).reshape(nb, 128, 1)   # saturation
    iq[:, 4] = 0
    # random +-A frames from moderate levels up to hard clipping (the saturating butterflies at work)
    rng = np.random.default_rng(21)
    for c, amp in zip(range(19, 27), (9000, 10900, 11000, 11100, 11600, 13000, 20000, 32767)):
        iq[:, c] = (rng.integers(0, 2, (nb, 128, 2)) * 2 - 1) * amp
    bank = make_bank(rd, nc, rd.STAGE_SPEC256, spec256_naverage=nav)
    chans = [po.OracleChan(po.default_config(stage_mask=po.STAGE_SPEC256, spec256_naverage=nav)) for _ in range(nc)]
    n_ready = 0
    for b0 in range(0, nb, blocks_per_call):
        chunk = np.ascontiguousarray(iq[b0:b0 + blocks_per_call])
        bank.process_blocks(chunk.shape[0], chunk, None)
        spec, ready = bank.read_spectrum()
        o_ready = []
        for c, ch in enumerate(chans):
            ch.process(chunk[:, c])
            o_spec, r = ch.read_spectrum()
            o_ready.append(r)
            assert np.array_equal(spec[c], o_spec), (b0, c)
        assert list(ready.astype(bool)) == o_ready
        n_ready += int(ready[0])
        trace, sm = bank.read_panadapter()
        for c, ch in enumerate(chans):
            o_trace, o_sm = ch.read_panadapter()
            assert np.array_equal(trace[c], o_trace) and sm[c] == np.float32(o_sm), (b0, c)
    assert n_ready > 0
    # waterfall history kept on the device: rows and colour classes equal the oracle's (row 0 = newest line)
    rows, col = bank.read_waterfall()
    for c, ch in enumerate(chans):
        o_rows, o_col = ch.read_waterfall()
        assert np.array_equal(rows[c], o_rows) and np.array_equal(col[c], o_col), c
    assert rows.any()


@pytest.mark.parametrize("blocks_per_call", [1, 3, 8, 13])
def test_spec1024_bit_exact(rd, po, blocks_per_call):
    nc, nb = 12, 39
    iq = synth.synth_iq(np.arange(50, 50 + nc), nb, [c % 5 for c in range(nc)])
    params = [po.default_params(demod=c % 5) for c in range(nc)]
    # loud audio on both sides of the bound below which the kernel takes its saturation-free butterflies
    # (fft_q15.cuh: NO_SAT_BOUND_1024_REAL = 11000 on the windowed samples), up to hard clipping
    for c, g in zip(range(7, 12), (1.6, 2.2, 3.0, 5.0, 12.0)):
        params[c] = po.default_params(demod=c % 2, in_gain=g)
    sm = rd.STAGE_FRONTEND | rd.STAGE_SPEC1024           # bit-exact audio => the spectrum must be bit-exact too
    bank = make_bank(rd, nc, sm)
    chans = []
    for c, p in enumerate(params):
        bank.set_mode(c, 1, to_rd_params(rd, p))
        chans.append(po.OracleChan(po.default_config(stage_mask=sm), p))
    seen = 0
    for b0 in range(0, nb, blocks_per_call):
        chunk = np.ascontiguousarray(iq[b0:b0 + blocks_per_call])
        g = bank.process_host(chunk)
        spec, ready = bank.read_audio_spectrum()
        for c, ch in enumerate(chans):
            o = ch.process(chunk[:, c])
            assert np.array_equal(g[:, c], o)
            o_spec, r = ch.read_audio_spectrum()
            assert bool(ready[c]) == r
            assert np.array_equal(spec[c], o_spec), (b0, c)
        seen += int(ready[0])
    assert seen > 0


# ------------------------------------------------------------------------------------------ whole chain

def _all_mode_params(po, nc):
    """config 5 of BASELINE.json as bench.py defines it (bench.channel_params): mode by c mod 4, notch on CW channels,
    DNR level by c mod 5, AGC by c mod 4"""
    params = [po.default_params(**bench.channel_params("cfg5", c)) for c in range(nc)]
    return params, [p.demod for p in params]


def test_full_chain_all_modes(rd, po):
    """The whole graph on 40 all-mode channels over 96 blocks: every f32 stage group within 1e-4 on identical inputs,
    integer stages and spectra bit-exact, demodulated-audio SNR equal within 0.1 dB on every channel — NR on or off,
    AGC on or off (tests/parity_util.py)."""
    nc, nb = 40, 96
    iq = bench.make_inputs("cfg5", 0, nc, nb)
    report = {}
    g_out, bank, chans = check_chain(rd, po, lambda c: bench.channel_params("cfg5", c), rd.STAGE_ALL, nc, iq, np.arange(nc), 8,
                                     snr_from=40, report=report)
    print("full chain:", report)
    aspec, aready = bank.read_audio_spectrum()
    assert aready.all() and aspec.any()


@pytest.mark.parametrize("stage", ["notch", "dnr"])
def test_ten_seconds_no_drift(rd, po, stage):
    """BASELINE config 1 length (10 s = 3446 blocks): the NLMS recurrences do not drift away from the oracle
    although the GPU sums the 96 taps in a different order."""
    nb, nc, win = 3446, 6, 200
    demod = [0, 1, 2, 3, 4, 0]
    iq = synth.synth_iq(np.arange(700, 700 + nc), nb, demod, interferer=True)
    if stage == "notch":
        sm = rd.STAGE_FRONTEND | rd.STAGE_NOTCH
        params = [po.default_params(demod=d, notch_on=1, notch_level=(15, 20, 30)[c % 3]) for c, d in enumerate(demod)]
    else:
        sm = rd.STAGE_FFTFILT | rd.STAGE_NR
        params = [po.default_params(nr_kind=po.NR_LMS, nr_level=(20, 30, 40, 50)[c % 4]) for c in range(nc)]
    _, g_f32, _, o_f32, _, _ = run_both(rd, po, sm, params, iq, blocks_per_call=53)
    errs = np.array([[rel_rms(g_f32[b:b + win, c, :, 0], o_f32[b:b + win, c, :, 0]) for b in range(0, nb - win, win)]
                     for c in range(nc)])
    assert errs.max() <= REL_RMS_TOL, errs.max()
    # the slow modes of the adaptive filter let the rounding-noise difference settle over seconds (30 s runs level off
    # at ~2e-5 for mu = 0.063 and stay < 1e-6 for mu <= 0.02, tools/diag_drift.py); it must not run away
    assert (errs[:, -3:].mean(axis=1) <= 0.5 * REL_RMS_TOL).all(), errs


def test_nlms_bits_do_not_depend_on_the_launch_list(rd, po):
    """what a channel computes depends neither on the form the launcher picks nor on which other channels run the notch / the
    DNR: toggling the neighbours' settings (=> other lists, other warps, other CTAs) leaves a channel's notch error
    signal, DNR estimate and audio bit-identical"""
    nc, nb = 41, 12
    iq = bench.make_inputs("cfg5", 0, nc, nb)
    keep = [2, 6, 10, 14, 17, 23, 38]                                       # CW channels with notch (+ DNR on most), and others
    outs = []
    for variant in range(3):
        bank = make_bank(rd, nc, rd.STAGE_ALL, max_blocks=4)
        for c in range(nc):
            p = dict(bench.channel_params("cfg5", c))
            if c not in keep and variant == 1:
                p.update(notch_on=1 - p["notch_on"], nr_kind=1, nr_level=(30, 50)[c % 2])
            if c not in keep and variant == 2:
                p.update(notch_on=int(c % 3 == 0), nr_kind=0, nr_level=0)
            bank.set_mode(c, 1, rd.default_params(**p))
        o = np.concatenate([bank.process_host(iq[b:b + 4]) for b in range(0, nb, 4)])
        outs.append((o[:, keep], bank.read_debug_f32(4)[:, keep], bank.read_audio_spectrum()[0][keep]))
    for v in (1, 2):
        for a, b in zip(outs[0], outs[v]):
            assert np.array_equal(a, b)


def test_nlms_forms_are_bit_identical(rd, po, monkeypatch):
    """k_nlms has three forms — 8 lanes per channel with scalar FMAs, 8 lanes on the packed f32x2 FMA (FFMA2), and 4 lanes
    that hold two tap segments each and add their partial sums first (= the xor-4 stage of the 8-lane butterfly).  Every
    form performs the same roundings in the same order: notch error signal, DNR estimate, audio and audio spectrum are
    the same bits, so the launcher is free to pick by list size (cfg3's 16 384 notch channels run the 4-lane form)."""
    nc, nb = 40, 12
    iq = bench.make_inputs("cfg5", 0, nc, nb)
    outs = []
    for lanes, packed in (("8", "0"), ("8", "1"), ("4", "0")):
        monkeypatch.setenv("RDSP_NLMS_LANES", lanes)
        monkeypatch.setenv("RDSP_NLMS_PACKED", packed)
        bank = make_bank(rd, nc, rd.STAGE_ALL, max_blocks=4)
        for c in range(nc):
            bank.set_mode(c, 1, rd.default_params(**bench.channel_params("cfg5", c)))
        o = [bank.process_host(iq[b:b + 4]) for b in range(0, nb, 4)]
        outs.append((np.concatenate(o), bank.read_debug_f32(4), bank.read_audio_spectrum()[0]))
    for k in (1, 2):
        for a, b in zip(outs[0], outs[k]):
            assert np.array_equal(a, b), k


def test_blocks_per_call_invariance(rd, po):
    """process_blocks(T) == T x process_block, bit for bit (state makes a clean round trip through HBM)"""
    nc, nb = 21, 24
    params, demod = _all_mode_params(po, nc)
    iq = synth.synth_iq(np.arange(nc), nb, demod)
    outs = []
    for step in (1, 4, 24):
        bank = make_bank(rd, nc, rd.STAGE_ALL, max_blocks=nb)
        for c, p in enumerate(params):
            bank.set_mode(c, 1, to_rd_params(rd, p))
        o = np.concatenate([bank.process_host(np.ascontiguousarray(iq[b:b + step])) for b in range(0, nb, step)])
        outs.append((o, bank.read_spectrum()[0], bank.read_audio_spectrum()[0]))
    for o, s, a in outs[1:]:
        assert np.array_equal(o, outs[0][0]) and np.array_equal(s, outs[0][1]) and np.array_equal(a, outs[0][2])


@pytest.mark.parametrize("layout", ["stereo", "mono"])
def test_one_block_calls_append_the_audio_row_themselves(rd, po, layout):
    """With one block per call (the sketch's own calling pattern) the kernel that emits a channel's audio — the DNR, else the
    FFT filter — appends the row to the ring of the audio spectrum itself, and k_spec1024 only runs on the ticks that
    complete a frame (rdsp_gpu.cu, fused_append).  Audio, both spectra and the `ready` flags are what 8-block calls give,
    in both audio layouts, through a parameter change in the middle (DNR off / on moves channels between the two emitters)."""
    nc, nb = 45, 32
    iq = bench.make_inputs("cfg5", 0, nc, nb)
    res = []
    for T in (8, 1):
        cfg = rd.default_config(n_channels=nc, stage_mask=rd.STAGE_ALL, max_blocks_per_call=8, io_location=rd.IO_HOST,
                                audio_layout=rd.AUDIO_MONO if layout == "mono" else rd.AUDIO_STEREO)
        bank = rd.ReceiverBank(cfg)
        for c in range(nc):
            bank.set_mode(c, 1, rd.default_params(**bench.channel_params("cfg5", c)))
        outs, specs = [], []
        for b in range(0, nb, T):
            if b == 16:
                for c in range(0, nc, 3):
                    p = dict(bench.channel_params("cfg5", c))
                    p.update(nr_kind=0 if p["nr_level"] else 1, nr_level=0 if p["nr_level"] else 30)
                    bank.set_mode(c, 1, rd.default_params(**p))
            outs.append(bank.process_host(iq[b:b + T]))
            if (b + T) % 8 == 0:
                specs.append(bank.read_audio_spectrum())
        res.append((np.concatenate(outs), specs, bank.read_spectrum()[0]))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][2], res[1][2])
    for (s8, r8), (s1, r1) in zip(res[0][1], res[1][1]):
        assert np.array_equal(s8, s1) and np.array_equal(r8, r1)
    assert res[1][1][-1][0].any()


def test_one_block_calls_mixed_with_longer_ones(rd, po):
    """one-block calls (rows appended by the emitting kernels, k_spec1024 gated by the host's mirror of the frame cadence) and
    longer calls (k_spec1024 appends) in any order keep ONE ring and ONE tick counter: audio and spectra equal those of 8-block calls"""
    nc, nb = 30, 40
    iq = bench.make_inputs("cfg5", 0, nc, nb)
    res = []
    for pattern in ((8,) * 5, (1, 1, 3, 1, 8, 1, 1, 1, 1, 2, 1, 5, 1, 1, 8, 1, 1, 1, 1)):
        assert sum(pattern) == nb
        bank = make_bank(rd, nc, rd.STAGE_ALL, max_blocks=8)
        for c in range(nc):
            bank.set_mode(c, 1, rd.default_params(**bench.channel_params("cfg5", c)))
        outs, b = [], 0
        for T in pattern:
            outs.append(bank.process_host(iq[b:b + T]))
            b += T
        res.append((np.concatenate(outs), bank.read_spectrum()[0], bank.read_audio_spectrum()[0]))
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)
    assert res[1][2].any()


def test_channel_range_sharding_matches_single_handle(rd, po):
    """N handles over contiguous channel ranges reproduce one handle byte for byte"""
    nc, nb = 37, 16
    params, demod = _all_mode_params(po, nc)
    iq = synth.synth_iq(np.arange(nc), nb, demod)
    whole = make_bank(rd, nc, rd.STAGE_ALL, max_blocks=nb)
    for c, p in enumerate(params):
        whole.set_mode(c, 1, to_rd_params(rd, p))
    ref = whole.process_host(iq)
    bounds = [0, 9, 10, 37]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        part = make_bank(rd, hi - lo, rd.STAGE_ALL, max_blocks=nb)
        for c in range(lo, hi):
            part.set_mode(c - lo, 1, to_rd_params(rd, params[c]))
        got = part.process_host(np.ascontiguousarray(iq[:, lo:hi]))
        assert np.array_equal(got, ref[:, lo:hi])
        assert np.array_equal(part.read_spectrum()[0], whole.read_spectrum(lo, hi - lo)[0])


def test_mode_changes_between_blocks(rd, po):
    """set_mode takes effect at the next block boundary; NR level changes re-init the NLMS like the sketch (C7)"""
    nc, nb = 6, 40
    iq = synth.synth_iq(np.arange(nc), nb, [0, 1, 2, 3, 4, 0], interferer=True)
    sm = rd.STAGE_ALL
    schedule = {
        0: dict(),
        8: dict(demod=po.DEMOD_USB, nr_kind=po.NR_LMS, nr_level=20, notch_on=1),
        16: dict(demod=po.DEMOD_AM, audio_filter=po.FILTER_AM, nr_kind=po.NR_LMS, nr_level=50, agc_mode=po.AGC_FAST, pbt_hi_hz=2500.0),
        24: dict(nr_kind=po.NR_OFF, nr_level=0, notch_on=0, out_gain=0.8),
        32: dict(demod=po.DEMOD_CW_LSB, audio_filter=po.FILTER_CW, nr_kind=po.NR_LMS, nr_level=50, notch_on=1, notch_level=30),
    }
    cum, cur = {}, {}
    for b0 in sorted(schedule):
        cur = dict(cur, **schedule[b0])
        cum[b0] = cur
    # every f32 stage group within 1e-4 on identical inputs THROUGH the re-initialisations (the schedule applies to the
    # stage-wise banks too), integer stages bit-exact, tone SNR equal within 0.1 dB over the last setting
    check_chain(rd, po, lambda c, b0: cum[b0], sm, nc, iq, np.arange(nc), 8, snr_from=32,
                tones=[400.0, 700.0, 1000.0, 1100.0, 1500.0, 1700.0, 2300.0])


# ------------------------------------------------------------------------------------------ edges, ABI on device

@pytest.mark.parametrize("nc", [1, 5, 17, 33])
def test_ragged_channel_counts(rd, po, nc):
    iq = bench.make_inputs("cfg5", 0, nc, 12)
    check_chain(rd, po, lambda c: bench.channel_params("cfg5", c), rd.STAGE_ALL, nc, iq, np.arange(nc), 5)


def test_extreme_inputs(rd, po):
    """digital silence, full-scale DC of both signs, full-scale Nyquist; gains that saturate every q15 stage"""
    nb = 10
    iq = np.zeros((nb, 4, 128, 2), np.int16)
    iq[:, 1] = 32767
    iq[:, 2] = -32768
    iq[:, 3] = np.where(np.arange(nb * 128).reshape(nb, 128, 1) % 2 == 0, 32767, -32768)
    params = [po.default_params(demod=c % 5, nr_kind=po.NR_LMS, nr_level=30, notch_on=1, in_gain=4.0) for c in range(4)]
    g_out, _, o_out, _, bank, chans = run_both(rd, po, rd.STAGE_ALL, params, iq, blocks_per_call=3)
    assert not g_out[:, 0].any()                                           # silence in, silence out
    # Blocks 3-4 are where the saturated front end collapses to EXACT digital zero: energy + eps -> eps, so the
    # reference recurrence multiplies every rounding difference by 1/eps = 8.4e6 and any two f32 evaluation orders
    # (the oracle's sequential sums, this kernel's grouped look-ahead) decorrelate until the window has drained.
    # Outside that window the usual closeness holds; inside it the output must stay bounded and drain to zero too.
    keep = np.ones(nb, bool)
    keep[3:5] = False
    assert np.abs(g_out[3:5].astype(np.int32)).max() <= 2 * np.abs(o_out.astype(np.int32)).max() + 64
    assert not g_out[5:].any() and not o_out[5:].any()
    spec = bank.read_audio_spectrum()[0]
    assert np.isfinite(spec.astype(float)).all()
    # the stages in front of the collapse are unaffected: bit-exact / 1 LSB ...
    g2, _, o2, _, _, _ = run_both(rd, po, rd.STAGE_FRONTEND | rd.STAGE_NOTCH | rd.STAGE_AGC, params, iq, blocks_per_call=3)
    assert np.abs(g2.astype(np.int32) - o2).max() <= 1
    # ... and so are K5, and K6 on identical inputs (the oracle's NLMS on the signal the GPU's K5 produced), outside the
    # collapse window
    g5, g5f, o5, _, _, _ = run_both(rd, po, rd.STAGE_FFTFILT, params, g2, blocks_per_call=3)
    assert np.abs(g5.astype(np.int32) - o5).max() <= 1
    g6, _, _, _, _, _ = run_both(rd, po, rd.STAGE_FFTFILT | rd.STAGE_NR, params, g2, blocks_per_call=3)
    for c in range(4):
        ch = po.OracleChan(po.default_config(stage_mask=po.STAGE_FFTFILT | po.STAGE_NR), params[c])
        want = np.zeros((nb, 128), np.int16)
        y = ch.dnr_f32(g5f[:, c, :, 0])
        po.lib().arm_float_to_q15(y.ctypes.data, want.ctypes.data, y.size)
        assert np.abs(g6[:, c, :, 0].astype(np.int32) - want)[keep].max() <= 1, c


def test_level_collapse_with_a_noise_floor(rd, po):
    """keyed carrier over a noise floor (what CW does to the DNR all day).  At a 40 dB on/off ratio the conditioning of
    the reference recurrence (gain mu / (energy + eps), energy kept as a running difference) costs digits: the ORACLE
    ITSELF moves by a few 1e-4 when its input moves by half an ulp.  The bar here is therefore the oracle's own
    sensitivity: on identical inputs (the oracle's NLMS on the signal the GPU's K5 produced) the kernel stays within
    1e-4 or within 4 x what a half-ulp input perturbation does to the oracle, whichever is larger.  At 60 dB the
    reference itself bursts above full scale and parity stops being defined; that case only has to stay finite."""
    nb, nc = 64, 4
    rng = np.random.default_rng(9)
    n = np.arange(nb * 128)
    key = ((n // 2646) % 2 == 0).astype(float)                              # 60 ms elements
    for floor, checked in ((120.0, True), (12.0, False)):
        iq = np.zeros((nb, nc, 128, 2), np.int16)
        for c in range(nc):
            amp = 12000.0 * key + floor
            z = amp * np.exp(2j * np.pi * (600.0 + 150 * c) * n / 44100.0) + rng.normal(0, floor / 2, n.size) + 1j * rng.normal(0, floor / 2, n.size)
            iq[:, c, :, 0] = np.rint(z.real).reshape(nb, 128)
            iq[:, c, :, 1] = np.rint(z.imag).reshape(nb, 128)
        params = [po.default_params(nr_kind=po.NR_LMS, nr_level=(20, 30, 40, 50)[c]) for c in range(nc)]
        _, g_f5, _, _, _, _ = run_both(rd, po, rd.STAGE_FFTFILT, params, iq, blocks_per_call=8)
        _, g_f6, _, _, _, _ = run_both(rd, po, rd.STAGE_FFTFILT | rd.STAGE_NR, params, iq, blocks_per_call=8)
        assert np.isfinite(g_f6[-8:]).all()                                 # whatever happens in a burst, the channel recovers
        if not checked:
            continue
        assert np.isfinite(g_f6).all()
        for c in range(nc):
            x = g_f5[:, c, :, 0]
            mk = lambda: po.OracleChan(po.default_config(stage_mask=po.STAGE_FFTFILT | po.STAGE_NR), params[c])
            want = mk().dnr_f32(x)
            moved = mk().dnr_f32((x.astype(np.float64) * (1.0 + 2.0 ** -24 * rng.choice([-1.0, 1.0], x.shape))).astype(np.float32))
            own = rel_rms(moved, want)                                      # the oracle against itself, input moved by half an ulp
            err = rel_rms(g_f6[:, c, :, 0], want)
            assert err <= max(REL_RMS_TOL, 4.0 * own), (c, err, own)


def test_argument_errors_on_device(rd):
    bank = make_bank(rd, 4, rd.STAGE_ALL, max_blocks=2)
    iq = np.zeros((3, 4, 128, 2), np.int16)
    out = np.zeros_like(iq)
    with pytest.raises(rd.RdspError) as e:
        bank.process_blocks(3, iq, out)                      # exceeds max_blocks_per_call
    assert e.value.code == -2
    bank.process_blocks(0, iq, out)                          # empty call is a no-op
    with pytest.raises(rd.RdspError):
        bank.process_blocks(1, iq.ctypes.data + 2, out)      # misaligned
    with pytest.raises(rd.RdspError):
        bank.set_mode(3, 2, rd.default_params())             # range overflow
    for bad in (dict(demod=9), dict(audio_filter=-1), dict(agc_mode=4), dict(nr_kind=3), dict(pbt_lo_hz=500.0, pbt_hi_hz=400.0)):
        with pytest.raises(rd.RdspError):
            bank.set_mode(0, 1, rd.default_params(**bad))
    with pytest.raises(rd.RdspError):
        make_bank(rd, 2, rd.STAGE_FRONTEND).read_spectrum()  # stage not present
    assert bank.kernel_launches == 0


def test_device_pointers_and_caller_stream(rd, po):
    """IO_DEVICE: torch tensors in HBM, work enqueued on the caller's stream (async), matches the host-IO path"""
    import torch
    nc, nb = 16, 8
    params, demod = _all_mode_params(po, nc)
    iq = synth.synth_iq(np.arange(nc), nb, demod)
    host_bank = make_bank(rd, nc, rd.STAGE_ALL, max_blocks=nb)
    cfg = rd.default_config(n_channels=nc, stage_mask=rd.STAGE_ALL, max_blocks_per_call=nb, io_location=rd.IO_DEVICE)
    cfg.async_ = 1
    dev_bank = rd.ReceiverBank(cfg)
    for c, p in enumerate(params):
        host_bank.set_mode(c, 1, to_rd_params(rd, p))
        dev_bank.set_mode(c, 1, to_rd_params(rd, p))
    want = host_bank.process_host(iq)
    s = torch.cuda.Stream()
    dev_bank.set_stream(s.cuda_stream)
    d_in = torch.from_numpy(iq).cuda()
    d_out = torch.zeros_like(d_in)
    with torch.cuda.stream(s):
        dev_bank.process_blocks(nb, d_in, d_out)
    s.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), want)
    assert dev_bank.kernel_launches >= 6
    prof = dev_bank.profile_read()
    assert set(prof) >= {"k_front", "k_fftfilt", "k_biquad", "k_spec256"}


def test_full_size_slot_independence(rd):
    """BASELINE config sizes (8192 channels): the same input in every slot gives identical output in every slot,
    and the result does not depend on the neighbours (size-independent property, no oracle needed)."""
    import torch
    nc, nb = 8192, 8
    one = synth.synth_iq([7], nb, [0])                                    # [nb,1,128,2]
    cfg = rd.default_config(n_channels=nc, stage_mask=rd.STAGE_ALL, max_blocks_per_call=nb, io_location=rd.IO_DEVICE)
    bank = rd.ReceiverBank(cfg)
    bank.set_mode(0, nc, rd.default_params(nr_kind=rd.NR_LMS, nr_level=30, notch_on=1))
    d_in = torch.from_numpy(one).cuda().expand(nb, nc, 128, 2).contiguous()
    d_out = torch.zeros_like(d_in)
    bank.process_blocks(nb, d_in, d_out)
    out = d_out.cpu().numpy()
    assert (out == out[:, :1]).all()
    spec = bank.read_spectrum()[0]
    assert (spec == spec[:1]).all()
    small = make_bank(rd, 3, rd.STAGE_ALL, max_blocks=nb)
    small.set_mode(0, 3, rd.default_params(nr_kind=rd.NR_LMS, nr_level=30, notch_on=1))
    ref = small.process_host(np.ascontiguousarray(np.broadcast_to(one, (nb, 3, 128, 2))))
    assert np.array_equal(out[:, 4000], ref[:, 1])


def test_nlms_direct_cross_check_kernel(rd, po, monkeypatch):
    """RDSP_NLMS_IMPL=direct (read at create): the sample-by-sample k_nlms_direct — the textbook evaluation order — meets
    the same bar against the oracle as the look-ahead kernel, on the notch and on the DNR"""
    monkeypatch.setenv("RDSP_NLMS_IMPL", "direct")
    nc, nb = 20, 24
    iq = bench.make_inputs("cfg5", 0, nc, nb)
    check_chain(rd, po, lambda c: bench.channel_params("cfg5", c), rd.STAGE_ALL, nc, iq, np.arange(nc), 8)


def test_two_devices_in_one_process(rd, po):
    """one handle per GPU inside ONE process: per-device kernel attributes (the > 48 KB shared-memory
    opt-in of k_front_tc, the carve-outs) are set on each device, and both handles give the single-device result"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    nc, nb = 300, 8
    iq = bench.make_inputs("cfg5", 0, nc, nb)
    outs = []
    for dev in (0, 1, 0):
        cfg = rd.default_config(n_channels=nc, device=dev, stage_mask=rd.STAGE_ALL, max_blocks_per_call=nb, io_location=rd.IO_HOST)
        bank = rd.ReceiverBank(cfg)
        for c in range(nc):
            bank.set_mode(c, 1, rd.default_params(**bench.channel_params("cfg5", c)))
        outs.append(bank.process_host(iq))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


@pytest.mark.parametrize("T", [1, 8])
def test_graph_replay_is_bit_identical(rd, po, T):
    """A call shape seen twice is captured into a CUDA graph and replayed with one launch (rdsp_gpu.cu, run_call): same
    bits as the kernel-by-kernel path — audio, both spectra, state carried over 40+ ticks, a parameter change in the middle
    (tables change => the captured shapes are dropped and rebuilt)."""
    import torch
    nc, nb = 300, 48
    iq = bench.make_inputs("cfg5", 0, nc, nb)
    res = []
    for mode in (rd.GRAPH_AUTO, rd.GRAPH_OFF):
        cfg = rd.default_config(n_channels=nc, stage_mask=rd.STAGE_ALL, max_blocks_per_call=T, io_location=rd.IO_DEVICE, graph_mode=mode)
        bank = rd.ReceiverBank(cfg)
        for c in range(nc):
            bank.set_mode(c, 1, rd.default_params(**bench.channel_params("cfg5", c)))
        d_in = [torch.empty((T, nc, 128, 2), dtype=torch.int16, device="cuda") for _ in range(2)]   # two rotating input buffers
        d_out = torch.zeros((T, nc, 128, 2), dtype=torch.int16, device="cuda")
        outs = []
        for k, b0 in enumerate(range(0, nb, T)):
            if b0 == nb // 2:
                bank.set_mode(5, 7, rd.default_params(demod=rd.DEMOD_USB, nr_kind=rd.NR_LMS, nr_level=40, notch_on=1))
            d_in[k % 2].copy_(torch.from_numpy(iq[b0:b0 + T]))
            bank.process_blocks(T, d_in[k % 2], d_out)
            outs.append(d_out.cpu().numpy().copy())
        res.append((np.concatenate(outs), bank.read_spectrum()[0], bank.read_audio_spectrum()[0], bank.graph_replays, bank.kernel_launches))
    (a_out, a_s, a_as, a_rep, a_l), (b_out, b_s, b_as, b_rep, b_l) = res
    assert np.array_equal(a_out, b_out) and np.array_equal(a_s, b_s) and np.array_equal(a_as, b_as)
    assert b_rep == 0 and a_rep >= nb // T - 12, (a_rep, b_rep)      # 2 buffers x 2 phases, seen + captured, twice (tables changed once)
    assert a_l == b_l                                                 # a replay accounts for the kernels it runs


@pytest.mark.parametrize("stage", ["fe", "fe_agc", "ff", "all"])
def test_mono_layout_is_the_left_channel(rd, po, stage):
    """cfg.audio_layout = RDSP_AUDIO_MONO: audio_out is [n_blocks][C][128] = L, bit for bit what the stereo layout puts into
    its left channel — from every kernel that can end a chain (front end, AGC, FFT filter, DNR) — and the audio spectrum,
    which reads the output rows, is unchanged.  Half the device-to-host bytes where L == R anyway."""
    sm = {"fe": rd.STAGE_FRONTEND, "fe_agc": rd.STAGE_FRONTEND | rd.STAGE_NOTCH | rd.STAGE_AGC | rd.STAGE_SPEC1024,
          "ff": rd.STAGE_FFTFILT | rd.STAGE_NR, "all": rd.STAGE_ALL}[stage]
    nc, nb = 45, 16
    iq = bench.make_inputs("cfg5", 0, nc, nb)
    res = []
    for layout in (rd.AUDIO_STEREO, rd.AUDIO_MONO):
        cfg = rd.default_config(n_channels=nc, stage_mask=sm, max_blocks_per_call=8, io_location=rd.IO_HOST, audio_layout=layout)
        bank = rd.ReceiverBank(cfg)
        for c in range(nc):
            bank.set_mode(c, 1, rd.default_params(**bench.channel_params("cfg5", c)))
        out = np.concatenate([bank.process_host(iq[b:b + 8]) for b in range(0, nb, 8)])
        res.append((out, bank.read_audio_spectrum()[0] if sm & rd.STAGE_SPEC1024 else None))
    (st, st_spec), (mo, mo_spec) = res
    assert mo.shape == (nb, nc, 128) and np.array_equal(mo, st[..., 0])
    if st_spec is not None:
        assert np.array_equal(st_spec, mo_spec) and st_spec.any()


def test_installed_mask_survives_set_mode_until_the_pbt_changes(rd, po):
    """rdsp_gpu_set_mask installs a mask as data; the sketch-style setters re-send the whole parameter block (get_mode ->
    set_mode), which must not silently drop it — only a change of the PBT cut-offs redesigns the mask (reInitializeFilter,
    RDSP_convolutional.h:209-224).  Identical masks are stored once; re-sending unchanged parameters rebuilds nothing."""
    rng = np.random.default_rng(4)
    mask = rng.normal(0, 0.5, 512).astype(np.float32)
    bank = make_bank(rd, 3, rd.STAGE_FFTFILT)
    designed = bank.get_mask(0).copy()
    bank.set_mask(0, 2, mask)
    bank.set_mask(2, 1, mask)                                   # same table again: shared row
    p = bank.get_mode(0); p.nr_level = 30; p.nr_kind = rd.NR_LMS
    bank.set_mode(0, 1, p)                                      # unrelated setting: the installed mask stays
    assert np.array_equal(bank.get_mask(0), mask) and np.array_equal(bank.get_mask(1), mask) and np.array_equal(bank.get_mask(2), mask)
    iq = synth.synth_iq([1, 2, 3], 4, [0, 0, 0])
    g = bank.process_host(iq)
    ch = po.OracleChan(po.default_config(stage_mask=po.STAGE_FFTFILT)); ch.set_mask(mask)
    assert np.abs(g[:, 1].astype(np.int32) - ch.process(iq[:, 1])).max() <= 1
    n0 = bank.kernel_launches
    bank.set_mode(0, 1, p)                                      # re-sent unchanged: nothing to rebuild
    bank.process_host(iq)
    assert bank.kernel_launches - n0 == 1
    p.pbt_hi_hz = 2500.0
    bank.set_mode(0, 1, p)                                      # PBT moved: the mask is designed again
    assert not np.array_equal(bank.get_mask(0), mask) and not np.array_equal(bank.get_mask(0), designed)
    assert np.abs(bank.get_mask(0) - po.design_mask(300.0, 2500.0)).max() < 2e-6
    assert np.array_equal(bank.get_mask(1), mask)
