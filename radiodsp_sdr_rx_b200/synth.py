"""Deterministic synthetic "40 m band" IQ, shared by the parity tests, the CPU oracle runs and bench.py.

Counter-based: every sample is a pure function of (seed, channel, absolute sample index), so any block
range of any channel subset can be generated independently.  Zero-IF scene relative to
the LO, fs = 44.1 kHz, int16, peak about -12 dBFS:

  LSB / USB : 5-tone voice surrogate (400, 700, 1100, 1700, 2300 Hz; amplitudes 1, .8, .6, .5, .4)
              on the lower / upper side of the suppressed carrier at 0 Hz
  CW        : carrier 700 Hz off the LO on the mode's side, on/off keyed in 60 ms elements (20 WPM)
  AM        : carrier at 0 Hz, 1 kHz tone, m = 0.5
  + optional heterodyne interferer at audio 1500 Hz, -6 dB (notch configs)
  + white Gaussian noise, SNR 10 dB in 3 kHz
"""
from __future__ import annotations

import numpy as np

SEED = 0x5D5DB200
FS = 44100.0
BLK = 128

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _u01(key: np.ndarray) -> np.ndarray:
    """uniform in (0, 1] from a uint64 key array"""
    h = _splitmix64(key)
    return ((h >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740992.0)


def _key(seed: int, channel: np.ndarray, index: np.ndarray, stream: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        k = _splitmix64(np.uint64(seed) + np.uint64(stream) * np.uint64(0x632BE59BD9B4E019))
        k = _splitmix64(k ^ channel.astype(np.uint64))
        return k ^ (index.astype(np.uint64) * np.uint64(0xD1342543DE82EF95) & _M64)


def synth_iq(channels, n_blocks: int, demod, first_block: int = 0, interferer=False,
             snr_db: float = 10.0, seed: int = SEED) -> np.ndarray:
    """int16 [n_blocks, len(channels), 128, 2] for the given absolute channel ids.

    demod: int or per-channel sequence of RDSP_DEMOD_* (0 LSB, 1 USB, 2 CW_LSB, 3 CW_USB, 4 AM, 5 SAM)
    interferer: bool or per-channel sequence
    """
    channels = np.asarray(channels, dtype=np.int64).reshape(-1)
    nc = channels.size
    demod = np.broadcast_to(np.asarray(demod, dtype=np.int64), (nc,))
    interferer = np.broadcast_to(np.asarray(interferer, dtype=bool), (nc,))
    ns = n_blocks * BLK
    n = first_block * BLK + np.arange(ns, dtype=np.int64)             # absolute sample index
    t = n.astype(np.float64) / FS
    chc = channels[:, None]

    def phase(stream):                                               # per-channel constant phases
        return 2.0 * np.pi * _u01(_key(seed, channels, np.zeros(nc, np.int64), stream))

    sig = np.zeros((nc, ns), np.complex128)
    side = np.where((demod == 0) | (demod == 2), -1.0, 1.0)[:, None]  # lower side for LSB / CW_LSB
    ssb = (demod == 0) | (demod == 1)
    cw = (demod == 2) | (demod == 3)
    am = (demod == 4) | (demod == 5)                                  # AM and SAM share the AM scene
    if ssb.any():
        tones = ((400.0, 1.0), (700.0, 0.8), (1100.0, 0.6), (1700.0, 0.5), (2300.0, 0.4))
        acc = np.zeros((nc, ns), np.complex128)
        for i, (f, a) in enumerate(tones):
            acc += a * np.exp(1j * (side * 2.0 * np.pi * f * t[None, :] + phase(10 + i)[:, None]))
        sig += np.where(ssb[:, None], acc / 3.3, 0.0)
    if cw.any():
        elem = (n // int(round(0.060 * FS)))[None, :]                  # 60 ms keying elements
        on = _u01(_key(seed, np.broadcast_to(chc, (nc, ns)), np.broadcast_to(elem, (nc, ns)), 20)) < 0.55
        car = np.exp(1j * (side * 2.0 * np.pi * 700.0 * t[None, :] + phase(21)[:, None]))
        sig += np.where(cw[:, None], on * car, 0.0)
    if am.any():
        env = 1.0 + 0.5 * np.cos(2.0 * np.pi * 1000.0 * t[None, :] + phase(30)[:, None])
        sig += np.where(am[:, None], env * np.exp(1j * phase(31)[:, None]) / 1.5, 0.0)
    if interferer.any():
        het = 0.5 * np.exp(1j * (side * 2.0 * np.pi * 1500.0 * t[None, :] + phase(40)[:, None]))
        sig += np.where(interferer[:, None], het, 0.0)

    # complex white noise: SNR 10 dB in 3 kHz for a unit-power signal
    p_noise_total = (10.0 ** (-snr_db / 10.0)) * 0.5 * (FS / 3000.0)
    idx = np.broadcast_to(n[None, :], (nc, ns))
    chb = np.broadcast_to(chc, (nc, ns))
    u1 = _u01(_key(seed, chb, idx, 1))
    u2 = _u01(_key(seed, chb, idx, 2))
    r = np.sqrt(-2.0 * np.log(u1)) * np.sqrt(p_noise_total / 2.0)
    noise = r * np.exp(2j * np.pi * u2)

    x = (sig + noise) * (0.25 / 3.0)                                  # peak about -12 dBFS
    iq = np.empty((nc, ns, 2), np.float64)
    iq[..., 0] = x.real
    iq[..., 1] = x.imag
    q = np.clip(np.rint(iq * 32768.0), -32768, 32767).astype(np.int16)
    return np.ascontiguousarray(q.reshape(nc, n_blocks, BLK, 2).transpose(1, 0, 2, 3))


def snr_db(ref: np.ndarray, test: np.ndarray) -> float:
    """SNR of `test` against `ref` (same shape), in dB."""
    ref = ref.astype(np.float64)
    err = test.astype(np.float64) - ref
    pe = float(np.mean(err * err))
    ps = float(np.mean(ref * ref))
    return float("inf") if pe == 0 else 10.0 * np.log10(ps / pe)


def tone_snr_db(audio: np.ndarray, freqs, fs: float = FS) -> float:
    """Demodulated-audio SNR: power at the known tone frequencies over everything else (least squares)."""
    x = audio.astype(np.float64).reshape(-1)
    n = np.arange(x.size)
    cols = []
    for f in freqs:
        cols += [np.cos(2 * np.pi * f * n / fs), np.sin(2 * np.pi * f * n / fs)]
    A = np.stack(cols, axis=1)
    coef, *_ = np.linalg.lstsq(A, x, rcond=None)
    s = A @ coef
    r = x - s
    return 10.0 * np.log10(np.mean(s * s) / max(np.mean(r * r), 1e-30))
