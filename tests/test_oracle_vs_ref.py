"""Pin the oracle: (1) the committed golden vectors produced by the reference's own sources compiled
unmodified (tools/make_golden.py), (2) the known-answer test of SURVEY.md section 4.3, (3) where the compiled
reference is present (oracle/_ref), a fresh bit-for-bit comparison on new seeded inputs."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from radiodsp_sdr_rx_b200 import synth


def _conv_cfg(po):
    return po.default_config(stage_mask=po.STAGE_FFTFILT | po.STAGE_NR)


def _run_levels(po, ch, iq, levels):
    """process block by block, changing nr_level like the sketch's UI does between ticks"""
    outs, f32s = [], []
    cur = None
    for k in range(iq.shape[0]):
        if levels[k] != cur:
            cur = levels[k]
            ch.set_mode(ch._par.copy(nr_kind=po.NR_LMS if cur > 0 else po.NR_OFF, nr_level=int(cur)))
        o, f = ch.process(iq[k:k + 1], True)
        outs.append(o)
        f32s.append(f)
    return np.concatenate(outs), np.concatenate(f32s)


@pytest.mark.parametrize("name", ["conv_nr0", "conv_nr30", "conv_levels"])
def test_port_matches_golden_conv(po, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    iq = np.stack([g["in_L"], g["in_R"]], axis=-1)
    nb = iq.shape[0]
    levels = g["nr_level"] if "nr_level" in g else np.zeros(nb, int)
    par = po.default_params(pbt_lo_hz=float(g["pbt"][0]), pbt_hi_hz=float(g["pbt"][1]))
    ch = po.OracleChan(_conv_cfg(po), par)
    ch._par = par
    out, f32 = _run_levels(po, ch, iq, levels)
    assert np.array_equal(out[..., 0], g["out_L"]) and np.array_equal(out[..., 1], g["out_R"])
    assert np.array_equal(f32[..., 0], g["f32_L"])
    if "mask" in g:
        assert np.array_equal(ch.get_mask(), g["mask"])


def test_known_answers_survey_4_3(po):
    g = np.load(os.path.join(GOLDEN, "conv_kat.npz"))
    ci, cq = po.calc_cplx_fir(300.0, 4000.0)
    assert np.array_equal(ci, g["fir_I"]) and np.array_equal(cq, g["fir_Q"])
    assert abs(ci[64] - 0.0839002268) < 1e-9 and cq[64] == 0.0 and abs(ci[0] + 1.99e-7) < 1e-8
    m = po.design_mask(300.0, 4000.0)
    assert np.array_equal(m, g["mask"])
    assert abs(m[24] - 1.000002) < 2e-6 and abs(m[25]) < 1e-6
    # mask response of the one-sided filter: 0 dB near 2.07 kHz, stop band on the negative side
    mag = np.abs(m[0::2] + 1j * m[1::2])
    assert abs(20 * np.log10(mag[12])) < 0.05 and 20 * np.log10(mag[256 - 6]) < -100
    assert abs(po.lib().rdsp_oracle_lms_mu(15) - 0.112202) < 1e-6
    for lvl, mu in ((20, 0.0631), (30, 0.01995), (40, 0.00631), (50, 0.001995)):
        assert abs(po.lib().rdsp_oracle_lms_mu(lvl) - mu) < 2e-5
    assert g["mu15"] == np.float32(po.lib().rdsp_oracle_lms_mu(15)) and g["mu30"] == np.float32(po.lib().rdsp_oracle_lms_mu(30))
    iq = np.stack([g["in_L"], g["in_R"]], axis=-1)
    ch = po.OracleChan(_conv_cfg(po))
    out = ch.process(iq)
    assert np.array_equal(out[..., 0], g["out_L"]) and np.array_equal(out[..., 1], g["out_R"])
    assert not out[0, :8].any()
    rms = np.sqrt(np.mean((out[-10:, :, 0] / 32768.0) ** 2))
    assert abs(rms - 0.1717) < 2e-4
    ch = po.OracleChan(_conv_cfg(po), po.default_params(nr_kind=po.NR_LMS, nr_level=30))
    out = ch.process(iq)
    assert np.array_equal(out[..., 0], g["out_L_nr30"]) and np.array_equal(out[..., 1], g["out_R_nr30"])
    assert abs(np.sqrt(np.mean((out[-10:, :, 0] / 32768.0) ** 2)) - 0.1889) < 2e-4


@pytest.mark.parametrize("key,nav", [("nav30", 30), ("nav4", 4), ("sat", 2)])
def test_port_matches_golden_spec256(po, key, nav):
    g = np.load(os.path.join(GOLDEN, "spec256.npz"))
    x = g["in_" + key]
    ch = po.OracleChan(po.default_config(stage_mask=po.STAGE_SPEC256, spec256_naverage=nav))
    res = ch.spec256_raw(x[..., 0], x[..., 1])
    assert [k for k, _ in res] == list(g["idx_" + key])
    assert np.array_equal(np.stack([o for _, o in res]), g["out_" + key])


def test_spec256_dc_lands_on_bin_127(po):
    """a DC IQ block peaks at output[127] (analyze_fft256iq.cpp:107)"""
    ch = po.OracleChan(po.default_config(stage_mask=po.STAGE_SPEC256, spec256_naverage=1))
    I = np.full((3, 128), 8000, np.int16)
    res = ch.spec256_raw(I, I)
    assert len(res) == 2 and int(np.argmax(res[-1][1])) == 127
    # positive frequencies sit at lower indices: +8 bins -> output[119]
    n = np.arange(3 * 128)
    z = 8000 * np.exp(2j * np.pi * 8 * n / 256)
    ch = po.OracleChan(po.default_config(stage_mask=po.STAGE_SPEC256, spec256_naverage=1))
    res = ch.spec256_raw(np.rint(z.real).astype(np.int16).reshape(3, 128), np.rint(z.imag).astype(np.int16).reshape(3, 128))
    assert int(np.argmax(res[-1][1])) == 119


def test_fresh_inputs_against_compiled_reference(po):
    if not po.ref_available():
        pytest.skip("oracle/_ref/librdsp_ref.so not built on this machine")
    iq = synth.synth_iq([900, 901], 30, [1, 4], seed=1234)
    levels = [0] * 4 + [40] * 9 + [20] * 7 + [0] * 3 + [20] * 7
    ref = po.RefChannel()
    ref.reinit_filter(0.0, 800.0)
    oL, oR, fL, fR = ref.conv(iq[:, 0, :, 0], iq[:, 0, :, 1], levels)
    par = po.default_params(pbt_lo_hz=0.0, pbt_hi_hz=800.0)
    ch = po.OracleChan(_conv_cfg(po), par)
    ch._par = par
    out, f32 = _run_levels(po, ch, iq[:, 0], levels)
    assert np.array_equal(out[..., 0], oL) and np.array_equal(out[..., 1], oR) and np.array_equal(f32[..., 0], fL)
    ref = po.RefChannel(naverage=7)
    want = ref.fft256(iq[:, 1, :, 0], iq[:, 1, :, 1])
    ch = po.OracleChan(po.default_config(stage_mask=po.STAGE_SPEC256, spec256_naverage=7))
    got = ch.spec256_raw(iq[:, 1, :, 0], iq[:, 1, :, 1])
    assert [k for k, _ in got] == [k for k, _ in want]
    assert all(np.array_equal(a[1], b[1]) for a, b in zip(got, want))
