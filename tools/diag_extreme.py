import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pyoracle as po, radiodsp_sdr_rx_b200 as rd
from test_gpu_parity import run_both, rel_rms
nb = 10
iq = np.zeros((nb, 4, 128, 2), np.int16)
iq[:, 1] = 32767
iq[:, 2] = -32768
iq[:, 3] = np.where(np.arange(nb * 128).reshape(nb, 128, 1) % 2 == 0, 32767, -32768)
for sm, name in ((rd.STAGE_ALL, "all"), (rd.STAGE_FRONTEND | rd.STAGE_NOTCH | rd.STAGE_AGC, "fe+notch+agc"), (rd.STAGE_FFTFILT | rd.STAGE_NR, "conv+dnr")):
    params = [po.default_params(demod=c % 5, nr_kind=po.NR_LMS, nr_level=30, notch_on=1, in_gain=4.0) for c in range(4)]
    g_out, g_f32, o_out, o_f32, bank, chans = run_both(rd, po, sm, params, iq, blocks_per_call=3)
    d = np.abs(g_out.astype(np.int32) - o_out)
    for c in range(4):
        print(name, "ch", c, "maxdiff", d[:, c].max(), "per block", [int(d[b, c].max()) for b in range(nb)], "out rms", round(float(o_out[:, c].std()), 1), "peak", int(np.abs(o_out[:, c]).max()))
