"""The host+device arithmetic of the kernels (lane-distributed f32 FFT-256, q15 radix-4 butterflies, exact
division-by-naverage, generated tables, tap/mask design) executed on the CPU and checked against the oracle.
tests/host/test_device_math.cu is compiled with nvcc as a plain host program; no GPU needed."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


def test_device_math_on_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "test_device_math")
    cmd = [nvcc, "-O2", "-std=c++17", "-Xcompiler", "-Wno-unknown-pragmas", "-Wno-deprecated-gpu-targets", "-o", exe,
           os.path.join(ROOT, "tests/host/test_device_math.cu"),
           os.path.join(ROOT, "radiodsp_sdr_rx_b200/csrc/host_design.cpp"),
           "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Xlinker", "-rpath=" + os.path.join(ROOT, "oracle")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr


def test_control_plane_port(tmp_path):
    """host/rdsp_controls.hpp (mode / filter / AGC / NR cycling, PBT stepping) against the tables of RDSP_controls.h"""
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "test_controls")
    r = subprocess.run([gxx, "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests/host/test_controls.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
