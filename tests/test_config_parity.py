"""Parity of the CUDA path against the CPU oracle ON THE BENCHED WORKLOADS, at the benched sizes.

Every BASELINE.json config runs here exactly as bench.py runs it — same `channel_params`, same `make_inputs`,
same channel count per GPU, same 8 blocks per call — so the launch shapes the numbers are quoted on (tile tables,
time segments, channel lists, the notch / plain class split, the NLMS launch of 16 384 listed channels in cfg3)
are the ones that are checked.

Per config (tests/parity_util.py, check_chain):
  * the whole chain, GPU against the oracle on a sample of >= 64 channels (first tile, last tile, a stride through
    the middle; every mode / level / AGC class of the config is in it): integer outputs bit-exact, IQ spectra
    bit-exact, demodulated-audio tone SNR equal within 0.1 dB on EVERY sampled channel (NR on or off, AGC on or off);
  * each float32 stage group on identical inputs, at full size: f32 within 1e-4 relative RMS per channel, q15 within 1 LSB;
  * all channels, not just the sample: slots that carry the same parameters and the same input (the workload
    repeats with a period of 2560 channels) give bit-identical output wherever they sit in the bank.
"""
import json
import os

import numpy as np
import pytest

import bench
from parity_util import check_chain

pytestmark = pytest.mark.gpu

T = 8                       # bench.py's blocks per call
CASES = [("cfg2", 16), ("cfg3", 16), ("cfg4a", 16), ("cfg4b", 16), ("cfg5", 32)]


def sample_channels(C):
    """>= 64 channels: the first and the last ones (first / last tile, ragged ends of every list) and a stride through
    the middle that is co-prime with the periods of the configs (4, 5, 20), so every class shows up"""
    s = set(range(0, 24)) | set(range(C - 24, C)) | set(range(97, C, (C // 48) | 1))
    return np.array(sorted(s))


@pytest.mark.parametrize("wl,nb", CASES)
def test_benched_workload_matches_the_oracle(rd, po, wl, nb):
    _, stage, C = bench.WORKLOADS[wl]
    iq = bench.make_inputs(wl, 0, C, nb)
    sub = sample_channels(C)
    assert sub.size >= 64
    classes = {json.dumps(bench.channel_params(wl, c), sort_keys=True) for c in range(min(C, 20))}
    assert classes == {json.dumps(bench.channel_params(wl, int(c)), sort_keys=True) for c in sub}   # every class sampled
    report = {"workload": wl, "channels": C, "blocks": nb, "sampled": int(sub.size)}
    g_out, _, _ = check_chain(rd, po, lambda c: bench.channel_params(wl, c), stage, C, iq, sub, T, report=report)
    # every channel of the bank: same parameters + same input => same bits, wherever the slot sits
    period = 2560                                                           # lcm(512 unique inputs, parameter period 20)
    for c0 in range(0, C - period, period):
        n = min(period, C - period - c0)
        assert np.array_equal(g_out[:, c0:c0 + n], g_out[:, c0 + period:c0 + period + n]), c0
    out_dir = os.environ.get("RDSP_PARITY_REPORT_DIR")
    if out_dir:
        with open(os.path.join(out_dir, f"parity_{wl}.json"), "w") as f:
            json.dump(report, f)
    print("parity report:", json.dumps(report))
