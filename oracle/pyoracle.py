"""pyoracle — TEST INFRASTRUCTURE: ctypes access to the CPU oracle.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  The product package (radiodsp_sdr_rx_b200) never does.

  liboracle.so           the restatement ("port") of the chain, oracle/rdsp_oracle.c
  _ref/librdsp_ref.so    the reference's own in-tree sources compiled unmodified (one channel per
                         loaded copy: the reference keeps its state in globals)
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# RDSP_ORACLE_LIB: another build of the same sources (bench.py times a -march=native one; parity always uses the portable build)
LIB_PATH = os.environ.get("RDSP_ORACLE_LIB") or os.path.join(HERE, "liboracle.so")
REF_PATH = os.path.join(HERE, "_ref", "librdsp_ref.so")

BLK = 128

# ---- mirrors of include/rdsp_gpu.h ---------------------------------------------------------
DEMOD_LSB, DEMOD_USB, DEMOD_CW_LSB, DEMOD_CW_USB, DEMOD_AM, DEMOD_SAM = range(6)
FILTER_CW, FILTER_2100, FILTER_2700, FILTER_3100, FILTER_AM = range(5)
AGC_OFF, AGC_FAST, AGC_MEDIUM, AGC_SLOW = range(4)
NR_OFF, NR_LMS, NR_SPECTRAL = range(3)
STAGE_FRONTEND, STAGE_NOTCH, STAGE_AGC, STAGE_FFTFILT, STAGE_NR, STAGE_SPEC256, STAGE_SPEC1024 = (1 << i for i in range(7))
STAGE_ALL = 0x7F
TAPS_HILBERT_I, TAPS_HILBERT_Q, TAPS_BANDPASS = range(3)


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("n_channels", C.c_uint32), ("device", C.c_int32),
        ("stage_mask", C.c_uint32), ("max_blocks_per_call", C.c_uint32), ("io_location", C.c_uint32),
        ("async_", C.c_uint32), ("debug_f32", C.c_uint32), ("spec256_naverage", C.c_uint32),
        ("agc_target", C.c_float), ("agc_max_gain", C.c_float), ("agc_attack_ms", C.c_float),
        ("agc_decay_ms", C.c_float * 4), ("pipeline_chunks", C.c_uint32),
        ("audio_layout", C.c_uint32), ("graph_mode", C.c_uint32),
    ]


class Params(C.Structure):
    _fields_ = [
        ("demod", C.c_int32), ("audio_filter", C.c_int32), ("agc_mode", C.c_int32),
        ("notch_on", C.c_int32), ("notch_level", C.c_int32), ("nr_kind", C.c_int32),
        ("nr_level", C.c_int32), ("pbt_lo_hz", C.c_float), ("pbt_hi_hz", C.c_float),
        ("in_gain", C.c_float), ("out_gain", C.c_float), ("iq_balance", C.c_float),
        ("als_peak", C.c_int32), ("nb_on", C.c_int32), ("nb_threshold_db", C.c_float),
    ]

    def copy(self, **kw):
        p = Params.from_buffer_copy(self)
        for k, v in kw.items():
            setattr(p, k, v)
        return p


def build(quiet: bool = True) -> None:
    """Compile liboracle.so (always) and _ref/librdsp_ref.so (when /root/reference is present)."""
    r = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if not quiet:
        print(r.stdout)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.rdsp_oracle_chan_create.restype = C.c_void_p
        L.rdsp_oracle_chan_create.argtypes = [C.POINTER(Config)]
        L.rdsp_oracle_chan_destroy.argtypes = [C.c_void_p]
        L.rdsp_oracle_chan_set_mode.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.rdsp_oracle_chan_process.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p,
                                               C.c_size_t, C.c_void_p, C.c_size_t]
        L.rdsp_oracle_chan_spec256_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.rdsp_oracle_chan_dnr_f32.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.rdsp_oracle_chan_read_spectrum.argtypes = [C.c_void_p, C.c_void_p]
        L.rdsp_oracle_chan_read_audio_spectrum.argtypes = [C.c_void_p, C.c_void_p]
        L.rdsp_oracle_chan_read_panadapter.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.rdsp_oracle_chan_read_waterfall.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.rdsp_oracle_chan_get_mask.argtypes = [C.c_void_p, C.c_void_p]
        L.rdsp_oracle_chan_set_mask.argtypes = [C.c_void_p, C.c_void_p]
        L.rdsp_oracle_calc_cplx_fir.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        L.rdsp_oracle_design_mask.argtypes = [C.c_double, C.c_double, C.c_void_p]
        L.rdsp_oracle_get_taps.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.rdsp_oracle_set_taps.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.rdsp_oracle_lms_mu.restype = C.c_float
        L.rdsp_oracle_lms_mu.argtypes = [C.c_int]
        L.rdsp_oracle_default_params.argtypes = [C.POINTER(Params)]
        L.rdsp_oracle_default_config.argtypes = [C.POINTER(Config)]
        L.rdsp_oracle_bank_process.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                               C.c_void_p, C.c_void_p]
        L.oracle_sqrt_uint32_approx.restype = C.c_uint32
        L.oracle_sqrt_uint32_approx.argtypes = [C.c_uint32]
        L.oracle_hanning256.restype = C.POINTER(C.c_int16)
        L.oracle_hanning1024.restype = C.POINTER(C.c_int16)
        L.oracle_twiddle_4096_q15.restype = C.POINTER(C.c_int16)
        L.oracle_twiddle_256_f32.restype = C.POINTER(C.c_float)
        L.arm_cfft_radix4_init_q15.argtypes = [C.c_void_p, C.c_uint16, C.c_uint8, C.c_uint8]
        L.arm_cfft_radix4_q15.argtypes = [C.c_void_p, C.c_void_p]
        L.arm_float_to_q15.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.oracle_fir_q15.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.oracle_biquad_set_highpass.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.oracle_biquad_update.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def default_config(**kw) -> Config:
    cfg = Config()
    lib().rdsp_oracle_default_config(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def default_params(**kw) -> Params:
    p = Params()
    lib().rdsp_oracle_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleChan:
    """One receiver channel of the CPU oracle."""

    def __init__(self, cfg: Config, params: Params | None = None):
        self._h = lib().rdsp_oracle_chan_create(C.byref(cfg))
        if not self._h:
            raise MemoryError
        if params is not None:
            self.set_mode(params)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().rdsp_oracle_chan_destroy(self._h)
            self._h = None

    @property
    def handle(self):
        return self._h

    def set_mode(self, p: Params):
        lib().rdsp_oracle_chan_set_mode(self._h, C.byref(p))

    def process(self, iq: np.ndarray, want_f32: bool = False):
        """iq: int16 [n_blocks,128,2] -> audio int16 [n_blocks,128,2] (and f32 of the same shape)."""
        iq = np.ascontiguousarray(iq, dtype=np.int16)
        nb = iq.shape[0]
        out = np.zeros((nb, BLK, 2), np.int16)
        f32 = np.zeros((nb, BLK, 2), np.float32) if want_f32 else None
        lib().rdsp_oracle_chan_process(self._h, nb, _ptr(iq), 2 * BLK, _ptr(out), 2 * BLK,
                                       _ptr(f32) if want_f32 else None, 2 * BLK)
        return (out, f32) if want_f32 else out

    def dnr_f32(self, x: np.ndarray) -> np.ndarray:
        """K6 alone (+ the 1.1 post-NR gain) on f32 blocks [n_blocks,128]: what the chain would emit for this K5 output."""
        x = np.ascontiguousarray(x, np.float32)
        y = np.zeros_like(x)
        lib().rdsp_oracle_chan_dnr_f32(self._h, x.shape[0], _ptr(x), _ptr(y))
        return y

    def spec256_raw(self, I: np.ndarray, Q: np.ndarray):
        """K9 alone (no biquads) on int16 [n_blocks,128] I/Q -> list of (block_index, output[256])."""
        I = np.ascontiguousarray(I, np.int16); Q = np.ascontiguousarray(Q, np.int16)
        res = []
        for k in range(I.shape[0]):
            lib().rdsp_oracle_chan_spec256_raw(self._h, _ptr(I[k]), _ptr(Q[k]))
            out, ready = self.read_spectrum()
            if ready:
                res.append((k, out))
        return res

    def read_spectrum(self):
        out = np.zeros(256, np.uint16)
        ready = lib().rdsp_oracle_chan_read_spectrum(self._h, _ptr(out))
        return out, bool(ready)

    def read_audio_spectrum(self):
        out = np.zeros(512, np.uint16)
        ready = lib().rdsp_oracle_chan_read_audio_spectrum(self._h, _ptr(out))
        return out, bool(ready)

    def read_panadapter(self):
        out = np.zeros(256, np.uint16)
        s = C.c_float()
        lib().rdsp_oracle_chan_read_panadapter(self._h, _ptr(out), C.byref(s))
        return out, s.value

    def read_waterfall(self):
        rows = np.zeros((50, 128), np.uint16)
        col = np.zeros((50, 128), np.uint8)
        lib().rdsp_oracle_chan_read_waterfall(self._h, _ptr(rows), _ptr(col))
        return rows, col

    def get_mask(self):
        m = np.zeros(512, np.float32)
        lib().rdsp_oracle_chan_get_mask(self._h, _ptr(m))
        return m

    def set_mask(self, m):
        m = np.ascontiguousarray(m, np.float32)
        lib().rdsp_oracle_chan_set_mask(self._h, _ptr(m))


def process_bank(cfg: Config, params_per_channel, iq: np.ndarray, want_f32: bool = False):
    """iq int16 [n_blocks, n_channels, 128, 2]; one fresh oracle channel per column."""
    nb, nc = iq.shape[:2]
    out = np.zeros_like(iq)
    f32 = np.zeros(iq.shape, np.float32) if want_f32 else None
    chans = []
    for c in range(nc):
        ch = OracleChan(cfg, params_per_channel[c] if isinstance(params_per_channel, (list, tuple)) else params_per_channel)
        r = ch.process(iq[:, c], want_f32)
        if want_f32:
            out[:, c], f32[:, c] = r
        else:
            out[:, c] = r
        chans.append(ch)
    return (out, f32, chans) if want_f32 else (out, chans)


def get_taps(kind: int, index: int) -> np.ndarray:
    t = np.zeros(129, np.int16)
    lib().rdsp_oracle_get_taps(kind, index, _ptr(t))
    return t


def design_mask(lo: float, hi: float) -> np.ndarray:
    m = np.zeros(512, np.float32)
    lib().rdsp_oracle_design_mask(lo, hi, _ptr(m))
    return m


def calc_cplx_fir(lo: float, hi: float, n: int = 129, fs: float = 44100.0):
    ci = np.zeros(n, np.float64)
    cq = np.zeros(n, np.float64)
    lib().rdsp_oracle_calc_cplx_fir(_ptr(ci), _ptr(cq), n, lo, hi, fs)
    return ci, cq


def ref_available() -> bool:
    return os.path.exists(REF_PATH)


class RefChannel:
    """One private copy of the compiled reference (its state is global => copy the .so)."""

    def __init__(self, naverage: int = 30):
        if not ref_available():
            raise FileNotFoundError(REF_PATH)
        self._tmp = tempfile.NamedTemporaryFile(prefix="rdsp_ref_", suffix=".so", delete=False)
        self._tmp.close()
        shutil.copyfile(REF_PATH, self._tmp.name)
        L = C.CDLL(self._tmp.name)
        L.ref_setup.argtypes = [C.c_int]
        L.ref_reinit_filter.argtypes = [C.c_double, C.c_double]
        L.ref_conv_push.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_conv_loop.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.ref_conv_float_out.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_get_mask.argtypes = [C.c_void_p]
        L.ref_get_fir.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_lms_mu.restype = C.c_float
        L.ref_lms_coeffs.argtypes = [C.c_void_p]
        L.ref_fft256_update.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_fft256_output.argtypes = [C.c_void_p]
        L.ref_run_blocks.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        self.L = L
        L.ref_setup(naverage)

    def __del__(self):
        try:
            os.unlink(self._tmp.name)
        except Exception:
            pass

    def reinit_filter(self, lo, hi):
        self.L.ref_reinit_filter(lo, hi)

    def mask(self):
        m = np.zeros(512, np.float32)
        self.L.ref_get_mask(_ptr(m))
        return m

    def fir(self):
        ci = np.zeros(129); cq = np.zeros(129)
        self.L.ref_get_fir(_ptr(ci), _ptr(cq))
        return ci, cq

    def conv(self, L: np.ndarray, R: np.ndarray, nr_level):
        """L, R int16 [n_blocks,128]; nr_level int or per-block sequence.  The gate at
        RDSP_convolutional.h:231 needs one block of look-ahead, so block k+1 is queued before
        the loop pass that processes block k; the last block gets a dummy."""
        nb = L.shape[0]
        L = np.ascontiguousarray(L, np.int16); R = np.ascontiguousarray(R, np.int16)
        outL = np.zeros((nb, BLK), np.int16); outR = np.zeros((nb, BLK), np.int16)
        fL = np.zeros((nb, BLK), np.float32); fR = np.zeros((nb, BLK), np.float32)
        zero = np.zeros(BLK, np.int16)
        levels = [nr_level] * nb if np.isscalar(nr_level) else list(nr_level)
        self.L.ref_conv_push(_ptr(L[0]), _ptr(R[0]))
        for k in range(nb):
            nxtL = L[k + 1] if k + 1 < nb else zero
            nxtR = R[k + 1] if k + 1 < nb else zero
            self.L.ref_conv_push(_ptr(nxtL), _ptr(nxtR))
            got = self.L.ref_conv_loop(int(levels[k]), _ptr(outL[k]), _ptr(outR[k]))
            assert got == 1
            self.L.ref_conv_float_out(_ptr(fL[k]), _ptr(fR[k]))
        return outL, outR, fL, fR

    def run_blocks(self, iq: np.ndarray, nr_level: int, with_fft: bool = True):
        """iq int16 [n_blocks,128,2] -> audio [n_blocks,128,2]: the in-tree stages (K5+K6+K7, K9) in one C loop (bench timing)"""
        iq = np.ascontiguousarray(iq, np.int16)
        out = np.zeros_like(iq)
        played = self.L.ref_run_blocks(iq.shape[0], _ptr(iq), _ptr(out), int(nr_level), int(with_fft))
        assert played == iq.shape[0]
        return out

    def lms_mu(self):
        return float(self.L.ref_lms_mu())

    def fft256(self, I: np.ndarray, Q: np.ndarray):
        """I, Q int16 [n_blocks,128] -> list of (block_index, output[256]) whenever available()."""
        I = np.ascontiguousarray(I, np.int16); Q = np.ascontiguousarray(Q, np.int16)
        res = []
        for k in range(I.shape[0]):
            if self.L.ref_fft256_update(_ptr(I[k]), _ptr(Q[k])):
                out = np.zeros(256, np.uint16)
                self.L.ref_fft256_output(_ptr(out))
                res.append((k, out))
        return res
