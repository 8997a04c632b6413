"""radiodsp_sdr_rx_b200 — batched RadioDSP_SDR_RX receive chain on NVIDIA B200 (sm_100a).

The product is the C-ABI shared library ``librdsp_gpu.so`` (include/rdsp_gpu.h) built from the CUDA
sources in ``csrc/``.  This package is the thin Python binding used by the tests and by bench.py;
there is no CPU fallback — importing :mod:`radiodsp_sdr_rx_b200.native` without the built library, or
creating a bank without an sm_100-class GPU, raises.
"""
from .native import (  # noqa: F401
    RdspError, Config, Params, ReceiverBank, default_config, default_params, lib, version,
    DEMOD_LSB, DEMOD_USB, DEMOD_CW_LSB, DEMOD_CW_USB, DEMOD_AM, DEMOD_SAM,
    FILTER_CW, FILTER_2100, FILTER_2700, FILTER_3100, FILTER_AM,
    AGC_OFF, AGC_FAST, AGC_MEDIUM, AGC_SLOW, NR_OFF, NR_LMS, NR_SPECTRAL,
    STAGE_FRONTEND, STAGE_NOTCH, STAGE_AGC, STAGE_FFTFILT, STAGE_NR, STAGE_SPEC256, STAGE_SPEC1024, STAGE_ALL,
    IO_DEVICE, IO_HOST, AUDIO_STEREO, AUDIO_MONO, GRAPH_AUTO, GRAPH_OFF, TAPS_HILBERT_I, TAPS_HILBERT_Q, TAPS_BANDPASS,
)

__all__ = ["ReceiverBank", "Config", "Params", "default_config", "default_params", "RdspError"]
