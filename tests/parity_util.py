"""Shared machinery of the GPU parity tests: the tolerances of BASELINE.json's north_star and the stage-wise
comparison of a chain against the oracle.

The chain has a q15 boundary in its middle (the SDR output block, K4 -> K5).  An f32 value that straddles a
truncation boundary there flips one LSB of the next stage's input, so "within 1e-4" is asserted where north_star
puts it — on every float32 STAGE, each side given identical inputs: K3+K4 behind the bit-exact integer front
end, K5 (and K5+K8) on the q15 audio the GPU's own K4 produced (handed to the oracle as that stage's input), K6 on
the f32 signal the GPU's own K5 produced (handed to the oracle's NLMS).  The whole chain is then held to what north_star asks of it: bit-exact integer outputs and spectra, and equal
demodulated-audio SNR to 0.1 dB on every channel.
"""
import inspect

import numpy as np

from radiodsp_sdr_rx_b200 import synth

REL_RMS_TOL = 1e-4          # float32 stages, relative RMS of the pre-quantisation signal
SNR_TOL_DB = 0.1
S_FE, S_NOTCH, S_AGC, S_FF, S_NR, S_S256, S_S1024 = (1 << i for i in range(7))


def rel_rms(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    den = np.sqrt(np.mean(b * b))
    return float(np.sqrt(np.mean((a - b) ** 2)) / den) if den > 0 else float(np.sqrt(np.mean((a - b) ** 2)))


def tone_freqs(demod):
    """the tones synth.synth_iq puts into a channel of this mode (audio Hz)"""
    if demod in (0, 1):
        return [400.0, 700.0, 1100.0, 1700.0, 2300.0]
    return [700.0] if demod in (2, 3) else [1000.0]


def _params_at(params_of, c, b0):
    """params_of(c) -> dict, or params_of(c, b0) -> dict for parameters that change between calls (b0 = first block of the call)"""
    return params_of(c, b0) if len(inspect.signature(params_of).parameters) >= 2 else params_of(c)


def gpu_run(rd, params_of, stage, C, iq, T, debug=True, **cfgkw):
    """one bank of C channels, host I/O, T blocks per call; parameters are (re)sent before every call, like the sketch's
    loop() does — the library ignores settings that did not change.
    Returns (q15 out [nb,C,128,2], f32 [nb,C,128,2] or None, bank)."""
    cfg = rd.default_config(n_channels=C, stage_mask=stage, max_blocks_per_call=T, io_location=rd.IO_HOST,
                            debug_f32=int(debug), **cfgkw)
    bank = rd.ReceiverBank(cfg)
    outs, f32s = [], []
    for b0 in range(0, iq.shape[0], T):
        for c in range(C):
            bank.set_mode(c, 1, rd.default_params(**_params_at(params_of, c, b0)))
        chunk = np.ascontiguousarray(iq[b0:b0 + T])
        outs.append(bank.process_host(chunk))
        if debug:
            f32s.append(bank.read_debug_f32(chunk.shape[0]))
    return np.concatenate(outs), (np.concatenate(f32s) if debug else None), bank


def oracle_run(po, params_of, stage, chans, iq_sub, T=None):
    """the oracle on the channels `chans` (absolute ids) of iq_sub [nb, len(chans), 128, 2]; same parameter schedule"""
    nb = iq_sub.shape[0]
    T = T or nb
    out = np.zeros_like(iq_sub)
    f32 = np.zeros(iq_sub.shape, np.float32)
    objs = []
    cfg = po.default_config(stage_mask=stage)
    for i, c in enumerate(chans):
        ch = po.OracleChan(cfg, po.default_params(**_params_at(params_of, int(c), 0)))
        for b0 in range(0, nb, T):
            ch.set_mode(po.default_params(**_params_at(params_of, int(c), b0)))
            out[b0:b0 + T, i], f32[b0:b0 + T, i] = ch.process(np.ascontiguousarray(iq_sub[b0:b0 + T, i]), True)
        objs.append(ch)
    return out, f32, objs


def check_chain(rd, po, params_of, stage, C, iq, sub, T, snr_from=None, report=None, tones=None):
    """GPU (all C channels) against the oracle (channels `sub`): whole chain + every f32 stage group on identical
    inputs.  Returns (g_out, bank, oracle channel objects of `sub`)."""
    nb = iq.shape[0]
    sub = np.asarray(sub)
    g_out, _, bank = gpu_run(rd, params_of, stage, C, iq, T, debug=False)
    o_out, _, chans = oracle_run(po, params_of, stage, sub, iq[:, sub], T)
    if not stage & (S_NOTCH | S_AGC | S_FF | S_NR):
        assert np.array_equal(g_out[:, sub], o_out)                       # K0-K2 only: integer, bit-exact
    if stage & S_S256 and nb > 31:
        spec, ready = bank.read_spectrum()
        assert ready.all()
        for i, ch in enumerate(chans):
            assert np.array_equal(spec[sub[i]], ch.read_spectrum()[0]), int(sub[i])
    if stage & (S_FE | S_FF):
        b0 = nb // 2 if snr_from is None else snr_from
        for i, c in enumerate(sub):
            f = tones or tone_freqs(_params_at(params_of, int(c), b0).get("demod", 0))
            a = synth.tone_snr_db(g_out[b0:, c, :, 0], f)
            b = synth.tone_snr_db(o_out[b0:, i, :, 0], f)
            assert abs(a - b) <= SNR_TOL_DB, (int(c), _params_at(params_of, int(c), b0), a, b)
    if report is not None:
        d = np.abs(g_out[:, sub].astype(np.int32) - o_out)
        report["whole_chain_max_lsb"] = int(d.max())
        report["whole_chain_snr_db_min"] = float(min(synth.snr_db(o_out[:, i], g_out[:, c]) for i, c in enumerate(sub)))

    k4_q15 = None
    pre = stage & (S_FE | S_NOTCH | S_AGC)
    if pre & (S_NOTCH | S_AGC):
        fe_out, _, _ = gpu_run(rd, params_of, S_FE, C, iq, T, debug=False)
        fe_ora, _, _ = oracle_run(po, params_of, S_FE, sub, iq[:, sub], T)
        assert np.array_equal(fe_out[:, sub], fe_ora)                       # the input of K3 / K4 is bit-exact ...
        del fe_out
        k4_q15, g_f32, _ = gpu_run(rd, params_of, pre, C, iq, T)
        o_q15, o_f32, _ = oracle_run(po, params_of, pre, sub, iq[:, sub], T)
        errs = [rel_rms(g_f32[:, c], o_f32[:, i]) for i, c in enumerate(sub)]
        if report is not None:
            report["K3K4_rel_rms_max"] = max(errs)
        assert max(errs) <= REL_RMS_TOL, ("K3+K4", int(sub[int(np.argmax(errs))]), max(errs))   # ... so K3 + K4 see identical inputs
        assert np.abs(k4_q15[:, sub].astype(np.int32) - o_q15).max() <= 1
        del g_f32
    if stage & S_FF:
        src = k4_q15 if k4_q15 is not None else iq                          # the audio K4 produced (L = R), or the raw input
        # K5 alone (the NR stage switched off on both sides), identical q15 inputs
        g_q5, g_f5, _ = gpu_run(rd, params_of, S_FF, C, src, T)
        o_q5, o_f5, _ = oracle_run(po, params_of, S_FF, sub, src[:, sub], T)
        errs = [rel_rms(g_f5[:, c], o_f5[:, i]) for i, c in enumerate(sub)]
        if report is not None:
            report["K5_rel_rms_max"] = max(errs)
        assert max(errs) <= REL_RMS_TOL, ("K5", int(sub[int(np.argmax(errs))]), max(errs))
        assert np.abs(g_q5[:, sub].astype(np.int32) - o_q5).max() <= 1
    if stage & S_NR:
        g_q6, g_f6, _ = gpu_run(rd, params_of, S_FF | S_NR, C, src, T)
        o_q6, o_f6, _ = oracle_run(po, params_of, S_FF | S_NR, sub, src[:, sub], T)
        errs = []
        for i, c in enumerate(sub):
            want_f, want_q = o_f6[:, i, :, 0], o_q6[:, i, :, 0]
            if any(_params_at(params_of, int(c), b0).get("nr_kind", 0) == 1 for b0 in range(0, nb, T)):
                # K6 (NLMS DNR) alone, on IDENTICAL f32 inputs: the oracle's NLMS runs on the L signal the GPU's K5 produced.
                # (K6 multiplies the ~1e-7 by which two FFT algorithms differ — here the kernel's and the oracle's, on the
                # radio CMSIS's radix-8 — by ~1e3 in the two blocks after its same-block-reference first call, SURVEY.md C6:
                # comparing K5+K6 as one stage would measure that amplification, not the kernel.)
                ch = po.OracleChan(po.default_config(stage_mask=S_FF | S_NR), po.default_params(**_params_at(params_of, int(c), 0)))
                want_f = np.zeros((nb, 128), np.float32)
                for b0 in range(0, nb, T):
                    ch.set_mode(po.default_params(**_params_at(params_of, int(c), b0)))
                    want_f[b0:b0 + T] = ch.dnr_f32(g_f5[b0:b0 + T, c, :, 0])
                want_q = np.zeros((nb, 128), np.int16)
                po.lib().arm_float_to_q15(want_f.ctypes.data, want_q.ctypes.data, want_f.size)
            errs.append(rel_rms(g_f6[:, c, :, 0], want_f))                 # K8 / NR off: K5(+K8) on identical q15 inputs
            assert np.abs(g_q6[:, c, :, 0].astype(np.int32) - want_q).max() <= 1, int(c)
        if report is not None:
            report["K6K8_rel_rms_max"] = max(errs)
        assert max(errs) <= REL_RMS_TOL, ("K6/K8", int(sub[int(np.argmax(errs))]), max(errs))
    return g_out, bank, chans
