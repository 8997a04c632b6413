import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pyoracle as po, radiodsp_sdr_rx_b200 as rd
from test_gpu_parity import run_both, rel_rms
nb, nc = 64, 4
rng = np.random.default_rng(9)
n = np.arange(nb * 128)
key = ((n // 2646) % 2 == 0).astype(float)
floor = 12.0
iq = np.zeros((nb, nc, 128, 2), np.int16)
for c in range(nc):
    amp = 12000.0 * key + floor
    z = amp * np.exp(2j * np.pi * (600.0 + 150 * c) * n / 44100.0) + rng.normal(0, floor / 2, n.size) + 1j * rng.normal(0, floor / 2, n.size)
    iq[:, c, :, 0] = np.rint(z.real).reshape(nb, 128)
    iq[:, c, :, 1] = np.rint(z.imag).reshape(nb, 128)
params = [po.default_params(nr_kind=po.NR_LMS, nr_level=(20, 30, 40, 50)[c]) for c in range(nc)]
g_out, g_f32, o_out, o_f32, _, _ = run_both(rd, po, rd.STAGE_FFTFILT | rd.STAGE_NR, params, iq, blocks_per_call=8)
for c in range(nc):
    bad = ~np.isfinite(g_f32[:, c])
    print("level", params[c].nr_level, "gpu nonfinite", int(bad.sum()), "first bad block", (np.argwhere(bad.any(axis=(1, 2)))[:1].ravel().tolist()),
          "gpu peak %.2f oracle peak %.2f" % (np.nanmax(np.abs(np.where(np.isfinite(g_f32[:, c]), g_f32[:, c], 0))), np.abs(o_f32[:, c]).max()),
          "oracle nonfinite", int((~np.isfinite(o_f32[:, c])).sum()))
