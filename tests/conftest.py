"""pytest configuration: the `gpu` marker, build-once fixtures, import paths.

  python -m pytest tests -q -m "not gpu"   CPU: oracle vs golden vectors / compiled reference, host logic, ABI
  python -m pytest tests -q -m gpu         B200: parity of the CUDA path against the oracle, through the C ABI
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_HEX = "/root/reference/pre_compiled/RadioDSP_SDR_RX.ino.hex"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _make(path):
    r = subprocess.run(["make", "-C", path], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.fixture(scope="session", autouse=True)
def built_libs():
    """Build the oracle and the product library when sources are newer than the binaries (no-ops otherwise)."""
    _make(os.path.join(ROOT, "oracle"))
    _make(os.path.join(ROOT, "radiodsp_sdr_rx_b200"))
    yield


@pytest.fixture(scope="session")
def po():
    import pyoracle
    return pyoracle


@pytest.fixture(scope="session")
def rd():
    import radiodsp_sdr_rx_b200
    return radiodsp_sdr_rx_b200
