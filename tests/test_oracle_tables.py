"""Constant tables of the absent libraries: the formula-generated tables the oracle (and the product) use
must equal the bytes of the reference's shipped firmware image."""
import os

import numpy as np
import pytest

from conftest import REFERENCE_HEX


def _ihex_to_bin(path):
    mem, base, lo, hi = {}, 0, None, 0
    for line in open(path):
        line = line.strip()
        if not line.startswith(":"):
            continue
        b = bytes.fromhex(line[1:])
        n, addr, typ, data = b[0], (b[1] << 8) | b[2], b[3], b[4:4 + b[0]]
        if typ == 0:
            a = base + addr
            for i, v in enumerate(data):
                mem[a + i] = v
            lo = a if lo is None else min(lo, a)
            hi = max(hi, a + n)
        elif typ == 4:
            base = ((data[0] << 8) | data[1]) << 16
        elif typ == 2:
            base = ((data[0] << 8) | data[1]) << 4
    out = bytearray(hi - lo)
    for a, v in mem.items():
        out[a - lo] = v
    return bytes(out)


@pytest.fixture(scope="module")
def fw():
    if not os.path.exists(REFERENCE_HEX):
        pytest.skip("reference firmware image not present on this machine")
    img = _ihex_to_bin(REFERENCE_HEX)
    assert len(img) == 206012 and img[:4] == b"FCFB"
    return img


def test_hann_windows_match_firmware(fw, po):
    h256 = np.frombuffer(fw[0x1f2f4:0x1f2f4 + 512], "<i2")
    h1024 = np.frombuffer(fw[0x1eaf4:0x1eaf4 + 2048], "<i2")
    assert np.array_equal(np.ctypeslib.as_array(po.lib().oracle_hanning256(), (256,)), h256)
    assert np.array_equal(np.ctypeslib.as_array(po.lib().oracle_hanning1024(), (1024,)), h1024)


def test_q15_twiddles_match_firmware(fw, po):
    tw = np.frombuffer(fw[0x2012c:0x2012c + 6144 * 2], "<i2")
    assert np.array_equal(np.ctypeslib.as_array(po.lib().oracle_twiddle_4096_q15(), (6144,)), tw)


def test_f32_twiddles_match_firmware(fw, po):
    tw = np.frombuffer(fw[0x1f92c:0x1f92c + 512 * 4], "<f4")
    mine = np.ctypeslib.as_array(po.lib().oracle_twiddle_256_f32(), (512,))
    assert np.abs(mine - tw).max() <= 6e-8          # f32 rounding of the same angles


def test_sqrt_guess_table_matches_firmware(fw, po):
    import ctypes as C
    t = np.frombuffer(fw[0x1f558:0x1f558 + 66], "<u2")
    mine = np.array((C.c_uint16 * 33).in_dll(po.lib(), "sqrt_integer_guess_table"))
    assert np.array_equal(mine, t)


def test_literals_in_firmware(fw):
    """build facts the oracle relies on: 44100.0 and 1.1 as doubles, 1/32768, 32768, NLMS epsilon as floats"""
    assert np.frombuffer(fw[0x24b04:0x24b0c], "<f8")[0] == 44100.0
    assert np.frombuffer(fw[0x9430:0x9438], "<f8")[0] == 1.1
    assert np.frombuffer(fw[0x13838:0x1383c], "<f4")[0] == np.float32(1.0 / 32768.0)
    assert np.frombuffer(fw[0x13968:0x1396c], "<f4")[0] == np.float32(32768.0)
    assert np.frombuffer(fw[0x14998:0x1499c], "<f4")[0] == np.float32(1.19209289e-7)


def test_sqrt_approx_properties(po):
    f = po.lib().oracle_sqrt_uint32_approx
    assert f(0) == 0 and f(1) == 1 and f(4) == 2
    rng = np.random.default_rng(1)
    for v in list(rng.integers(1, 2**32 - 1, 2000, dtype=np.uint64)) + [2**31, 2**32 - 1, 65535 * 65535]:
        r = f(int(v))
        assert abs(r - np.sqrt(float(v))) <= max(2.0, 0.01 * np.sqrt(float(v)))
