"""The C-ABI library loads and exports every symbol include/rdsp_gpu.h declares; argument checking and the
no-fallback rule work without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "rdsp_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rdsp_gpu_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree(rd):
    from radiodsp_sdr_rx_b200 import native
    assert _declared_symbols() == sorted(native.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(rd):
    L = rd.lib()
    for s in _declared_symbols():
        assert hasattr(L, s), s


def test_struct_layouts_match_header(rd):
    from radiodsp_sdr_rx_b200 import native
    assert C.sizeof(native.Config) == 76 and C.sizeof(native.Params) == 60
    cfg = rd.default_config()
    assert cfg.struct_size == C.sizeof(native.Config)
    assert (cfg.n_channels, cfg.stage_mask, cfg.spec256_naverage) == (1, rd.STAGE_ALL, 30)
    p = rd.default_params()
    # setup() defaults of the sketch, RadioDSP_SDR_RX.ino:117-139,183
    assert (p.demod, p.audio_filter, p.agc_mode, p.notch_on, p.nr_kind, p.nr_level) == (
        rd.DEMOD_LSB, rd.FILTER_2700, rd.AGC_MEDIUM, 0, rd.NR_OFF, 0)
    assert (p.pbt_lo_hz, p.pbt_hi_hz, p.in_gain, p.out_gain) == (300.0, 4000.0, 1.0, 0.5)
    assert abs(p.iq_balance - 1.02) < 1e-6 and p.als_peak == 0 and p.nb_on == 0 and p.nb_threshold_db == 20.0


def test_oracle_and_product_defaults_agree(rd, po):
    a, b = rd.default_params(), po.default_params()
    assert bytes(a) == bytes(b)
    ca, cb = rd.default_config(), po.default_config()
    cb.io_location = ca.io_location
    cb.device = ca.device
    assert bytes(ca) == bytes(cb)


def test_create_rejects_bad_arguments(rd):
    L = rd.lib()
    h = C.c_void_p()
    assert L.rdsp_gpu_create(None, C.byref(h)) == -1
    cfg = rd.default_config()
    cfg.struct_size = 8
    assert L.rdsp_gpu_create(C.byref(cfg), C.byref(h)) == -1
    for kw, code in ((dict(n_channels=0), -2), (dict(stage_mask=0), -2), (dict(stage_mask=rd.STAGE_NOTCH), -5),
                     (dict(stage_mask=rd.STAGE_NR), -5), (dict(stage_mask=rd.STAGE_SPEC1024), -5),
                     (dict(spec256_naverage=0), -2), (dict(max_blocks_per_call=0), -2), (dict(io_location=7), -2)):
        assert L.rdsp_gpu_create(C.byref(rd.default_config(**kw)), C.byref(h)) == code, kw
        assert L.rdsp_gpu_last_error(None)
    assert L.rdsp_gpu_set_mode(None, 0, 1, None) == -1
    assert L.rdsp_gpu_process_block(None, None, None) == -1
    L.rdsp_gpu_destroy(None)


def test_no_cpu_fallback(rd):
    """Without a usable sm_100 device create() must fail loudly (RDSP_ERR_CUDA), never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rd.RdspError) as e:
        rd.ReceiverBank(rd.default_config(n_channels=2))
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """the product path must not route through oracle/ (sources and Python package)"""
    pkg = os.path.join(ROOT, "radiodsp_sdr_rx_b200")
    for dp, _, fs in os.walk(pkg):
        if "build" in dp or "__pycache__" in dp:
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                code = "\n".join(l for l in txt.splitlines() if not l.strip().startswith(("//", "#", "*", "/*", '"""')))
                assert "pyoracle" not in code and "liboracle" not in code and "rdsp_oracle" not in code, os.path.join(dp, f)


def test_cpp_host_mirror_compiles_and_fails_loudly_without_gpu(tmp_path, rd):
    """the C++ stub a maintainer would use (host/rdsp_sketch_api.hpp, examples/sketch_port.cpp) builds against the ABI"""
    import shutil
    import subprocess
    import torch
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "sketch_port")
    libdir = os.path.join(ROOT, "radiodsp_sdr_rx_b200")
    r = subprocess.run([gxx, "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "examples", "sketch_port.cpp"),
                        "-L" + libdir, "-lrdsp_gpu", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "4"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 2 and "no CPU fallback" in r.stderr


def test_design_bandpass_needs_no_device():
    """rdsp_gpu_design_bandpass: the designer of the default band-pass bank for any band (audioWSPR = 1400..1600 Hz,
    RDSP_controls.h:392-402); runs on the host, reproduces the preset rows of the oracle's tap bank"""
    import numpy as np
    import radiodsp_sdr_rx_b200 as rd
    from radiodsp_sdr_rx_b200 import native
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import pyoracle as po
    assert np.array_equal(native.design_bandpass(150.0, 2700.0), po.get_taps(po.TAPS_BANDPASS, po.FILTER_2700))
    w = native.design_bandpass(1400.0, 1600.0).astype(np.float64) / 32768.0
    f = np.fft.rfft(w, 8192)
    gain = lambda hz: abs(f[int(round(hz * 8192 / 44100.0))])
    assert abs(gain(1500.0) - 1.0) < 0.02 and gain(200.0) < 0.01 and gain(3000.0) < 0.01   # 129 taps: a 1.4 kHz transition band
    with __import__("pytest").raises(ValueError):
        native.design_bandpass(3000.0, 1000.0)
