"""ctypes binding of include/rdsp_gpu.h (librdsp_gpu.so, built in-tree by `make -C radiodsp_sdr_rx_b200`)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RDSP_GPU_LIB") or os.path.join(HERE, "librdsp_gpu.so")      # override: A/B builds of the same ABI

BLK = 128

DEMOD_LSB, DEMOD_USB, DEMOD_CW_LSB, DEMOD_CW_USB, DEMOD_AM, DEMOD_SAM = range(6)
FILTER_CW, FILTER_2100, FILTER_2700, FILTER_3100, FILTER_AM = range(5)
AGC_OFF, AGC_FAST, AGC_MEDIUM, AGC_SLOW = range(4)
NR_OFF, NR_LMS, NR_SPECTRAL = range(3)
STAGE_FRONTEND, STAGE_NOTCH, STAGE_AGC, STAGE_FFTFILT, STAGE_NR, STAGE_SPEC256, STAGE_SPEC1024 = (1 << i for i in range(7))
STAGE_ALL = 0x7F
IO_DEVICE, IO_HOST = 0, 1
AUDIO_STEREO, AUDIO_MONO = 0, 1
GRAPH_AUTO, GRAPH_OFF = 0, 1
TAPS_HILBERT_I, TAPS_HILBERT_Q, TAPS_BANDPASS = range(3)

#: every symbol include/rdsp_gpu.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "rdsp_gpu_default_config", "rdsp_gpu_default_params", "rdsp_gpu_create", "rdsp_gpu_destroy",
    "rdsp_gpu_set_mode", "rdsp_gpu_get_mode", "rdsp_gpu_process_block", "rdsp_gpu_process_blocks",
    "rdsp_gpu_synchronize", "rdsp_gpu_stream_join", "rdsp_gpu_set_stream", "rdsp_gpu_read_spectrum", "rdsp_gpu_read_audio_spectrum",
    "rdsp_gpu_read_panadapter", "rdsp_gpu_read_waterfall", "rdsp_gpu_set_taps", "rdsp_gpu_get_taps", "rdsp_gpu_design_bandpass", "rdsp_gpu_set_mask", "rdsp_gpu_get_mask",
    "rdsp_gpu_read_debug_f32", "rdsp_gpu_kernel_launches", "rdsp_gpu_graph_replays", "rdsp_gpu_profile", "rdsp_gpu_profile_read",
    "rdsp_gpu_last_error", "rdsp_gpu_version",
]


class RdspError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rdsp_gpu error {code}: {msg}")
        self.code = code


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


_lib = None


def lib():
    """Load librdsp_gpu.so.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {HERE}` (nvcc, sm_100a). "
                              "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
        L.rdsp_gpu_default_config.argtypes = [C.POINTER(Config)]
        L.rdsp_gpu_default_params.argtypes = [C.POINTER(Params)]
        L.rdsp_gpu_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.rdsp_gpu_destroy.argtypes = [vp]
        L.rdsp_gpu_set_mode.argtypes = [vp, u32, u32, C.POINTER(Params)]
        L.rdsp_gpu_get_mode.argtypes = [vp, u32, C.POINTER(Params)]
        L.rdsp_gpu_process_block.argtypes = [vp, vp, vp]
        L.rdsp_gpu_process_blocks.argtypes = [vp, u32, vp, vp]
        L.rdsp_gpu_synchronize.argtypes = [vp]
        L.rdsp_gpu_stream_join.argtypes = [vp]
        L.rdsp_gpu_set_stream.argtypes = [vp, vp]
        L.rdsp_gpu_read_spectrum.argtypes = [vp, u32, u32, vp, vp]
        L.rdsp_gpu_read_audio_spectrum.argtypes = [vp, u32, u32, vp, vp]
        L.rdsp_gpu_read_panadapter.argtypes = [vp, u32, u32, vp, vp]
        L.rdsp_gpu_read_waterfall.argtypes = [vp, u32, u32, vp, vp]
        L.rdsp_gpu_set_taps.argtypes = [vp, i32, i32, vp, u32]
        L.rdsp_gpu_get_taps.argtypes = [vp, i32, i32, vp, u32]
        L.rdsp_gpu_design_bandpass.argtypes = [C.c_float, C.c_float, vp, u32]
        L.rdsp_gpu_set_mask.argtypes = [vp, u32, u32, vp]
        L.rdsp_gpu_get_mask.argtypes = [vp, u32, vp]
        L.rdsp_gpu_read_debug_f32.argtypes = [vp, u32, u32, u32, vp]
        L.rdsp_gpu_kernel_launches.argtypes = [vp]
        L.rdsp_gpu_kernel_launches.restype = C.c_uint64
        L.rdsp_gpu_graph_replays.argtypes = [vp]
        L.rdsp_gpu_graph_replays.restype = C.c_uint64
        L.rdsp_gpu_profile.argtypes = [vp, i32]
        L.rdsp_gpu_profile_read.argtypes = [vp, i32, vp, vp, vp]
        L.rdsp_gpu_last_error.argtypes = [vp]
        L.rdsp_gpu_last_error.restype = C.c_char_p
        L.rdsp_gpu_version.restype = C.c_char_p
        _lib = L
    return _lib


def version() -> str:
    return lib().rdsp_gpu_version().decode()


def default_config(**kw) -> Config:
    cfg = Config()
    lib().rdsp_gpu_default_config(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, "async_" if k == "async" else k, v)
    return cfg


def default_params(**kw) -> Params:
    p = Params()
    lib().rdsp_gpu_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def design_bandpass(lo_hz: float, hi_hz: float) -> np.ndarray:
    """129 q15 band-pass taps for an arbitrary audio band (audioWSPR: 1400 .. 1600 Hz); no device needed"""
    t = np.zeros(129, np.int16)
    rc = lib().rdsp_gpu_design_bandpass(lo_hz, hi_hz, t.ctypes.data, 129)
    if rc != 0:
        raise ValueError(f"rdsp_gpu_design_bandpass({lo_hz}, {hi_hz}) failed: {rc}")
    return t


def _addr(x):
    """Device or host address of a numpy array / torch tensor / raw int."""
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


class ReceiverBank:
    """N independent receivers on one GPU: one rdsp_gpu_t handle."""

    def __init__(self, cfg: Config):
        self.cfg = Config.from_buffer_copy(cfg)
        self._h = C.c_void_p()
        rc = lib().rdsp_gpu_create(C.byref(self.cfg), C.byref(self._h))
        if rc != 0:
            self._h = None
            raise RdspError(rc, lib().rdsp_gpu_last_error(None).decode())
        self.n_channels = int(cfg.n_channels)
        self._destroy = lib().rdsp_gpu_destroy             # bound here: module globals are gone at interpreter exit

    def close(self):
        if getattr(self, "_h", None):
            self._destroy(self._h)
            self._h = None

    __del__ = close

    def _ck(self, rc):
        if rc < 0:
            raise RdspError(rc, lib().rdsp_gpu_last_error(self._h).decode())
        return rc

    # ---- parameters ------------------------------------------------------------------------
    def set_mode(self, ch_first: int, ch_count: int, p: Params):
        self._ck(lib().rdsp_gpu_set_mode(self._h, ch_first, ch_count, C.byref(p)))

    def get_mode(self, ch: int) -> Params:
        p = Params()
        self._ck(lib().rdsp_gpu_get_mode(self._h, ch, C.byref(p)))
        return p

    def set_taps(self, kind: int, index: int, taps):
        t = np.ascontiguousarray(taps, np.int16)
        self._ck(lib().rdsp_gpu_set_taps(self._h, kind, index, t.ctypes.data, t.size))

    def get_taps(self, kind: int, index: int) -> np.ndarray:
        t = np.zeros(129, np.int16)
        self._ck(lib().rdsp_gpu_get_taps(self._h, kind, index, t.ctypes.data, 129))
        return t

    def set_mask(self, ch_first: int, ch_count: int, mask):
        m = np.ascontiguousarray(mask, np.float32)
        assert m.size == 512
        self._ck(lib().rdsp_gpu_set_mask(self._h, ch_first, ch_count, m.ctypes.data))

    def get_mask(self, ch: int) -> np.ndarray:
        m = np.zeros(512, np.float32)
        self._ck(lib().rdsp_gpu_get_mask(self._h, ch, m.ctypes.data))
        return m

    # ---- block path ------------------------------------------------------------------------
    def process_block(self, iq, audio):
        self._ck(lib().rdsp_gpu_process_block(self._h, _addr(iq), _addr(audio)))

    def process_blocks(self, n_blocks: int, iq, audio):
        """iq / audio: [n_blocks, n_channels, 128, 2] int16, numpy (IO_HOST) or CUDA tensors (IO_DEVICE)."""
        self._ck(lib().rdsp_gpu_process_blocks(self._h, n_blocks, _addr(iq), 0 if audio is None else _addr(audio)))

    def process_host(self, iq: np.ndarray) -> np.ndarray:
        """Convenience for IO_HOST handles: numpy in, numpy out."""
        assert self.cfg.io_location == IO_HOST
        iq = np.ascontiguousarray(iq, np.int16)
        out = np.zeros(iq.shape[:3], np.int16) if self.cfg.audio_layout == AUDIO_MONO else np.zeros_like(iq)
        self.process_blocks(iq.shape[0], iq, out)
        if self.cfg.async_:
            self.synchronize()
        return out

    def synchronize(self):
        self._ck(lib().rdsp_gpu_synchronize(self._h))

    def stream_join(self):
        self._ck(lib().rdsp_gpu_stream_join(self._h))

    def set_stream(self, cuda_stream: int):
        self._ck(lib().rdsp_gpu_set_stream(self._h, cuda_stream))

    # ---- read-outs -------------------------------------------------------------------------
    def read_spectrum(self, ch_first=0, ch_count=None):
        n = self.n_channels - ch_first if ch_count is None else ch_count
        out = np.zeros((n, 256), np.uint16)
        ready = np.zeros(n, np.uint8)
        self._ck(lib().rdsp_gpu_read_spectrum(self._h, ch_first, n, out.ctypes.data, ready.ctypes.data))
        return out, ready

    def read_audio_spectrum(self, ch_first=0, ch_count=None):
        n = self.n_channels - ch_first if ch_count is None else ch_count
        out = np.zeros((n, 512), np.uint16)
        ready = np.zeros(n, np.uint8)
        self._ck(lib().rdsp_gpu_read_audio_spectrum(self._h, ch_first, n, out.ctypes.data, ready.ctypes.data))
        return out, ready

    def read_panadapter(self, ch_first=0, ch_count=None):
        n = self.n_channels - ch_first if ch_count is None else ch_count
        trace = np.zeros((n, 256), np.uint16)
        sm = np.zeros(n, np.float32)
        self._ck(lib().rdsp_gpu_read_panadapter(self._h, ch_first, n, trace.ctypes.data, sm.ctypes.data))
        return trace, sm

    def read_waterfall(self, ch_first=0, ch_count=None, colour=True):
        n = self.n_channels - ch_first if ch_count is None else ch_count
        rows = np.zeros((n, 50, 128), np.uint16)
        col = np.zeros((n, 50, 128), np.uint8) if colour else None
        self._ck(lib().rdsp_gpu_read_waterfall(self._h, ch_first, n, rows.ctypes.data, col.ctypes.data if colour else None))
        return rows, col

    def read_debug_f32(self, n_blocks: int, ch_first=0, ch_count=None) -> np.ndarray:
        n = self.n_channels - ch_first if ch_count is None else ch_count
        out = np.zeros((n_blocks, n, BLK, 2), np.float32)
        self._ck(lib().rdsp_gpu_read_debug_f32(self._h, n_blocks, ch_first, n, out.ctypes.data))
        return out

    # ---- instrumentation -------------------------------------------------------------------
    @property
    def kernel_launches(self) -> int:
        return int(lib().rdsp_gpu_kernel_launches(self._h))

    @property
    def graph_replays(self) -> int:
        return int(lib().rdsp_gpu_graph_replays(self._h))

    def profile(self, enable: bool):
        self._ck(lib().rdsp_gpu_profile(self._h, 1 if enable else 0))

    def profile_read(self) -> dict:
        names = (C.c_char_p * 16)()
        ms = (C.c_double * 16)()
        n = (C.c_uint64 * 16)()
        k = self._ck(lib().rdsp_gpu_profile_read(self._h, 16, names, ms, n))
        return {names[i].decode(): {"ms": ms[i], "launches": int(n[i])} for i in range(k)}
