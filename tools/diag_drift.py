import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pyoracle as po, radiodsp_sdr_rx_b200 as rd
from radiodsp_sdr_rx_b200 import synth
from test_gpu_parity import run_both, rel_rms
nb, nc, win = 10338, 4, 500          # 30 s
iq = synth.synth_iq(np.arange(700, 700 + nc), nb, [0, 1, 2, 4], interferer=True)
sm = rd.STAGE_FFTFILT | rd.STAGE_NR
params = [po.default_params(nr_kind=po.NR_LMS, nr_level=(20, 30, 40, 50)[c % 4]) for c in range(nc)]
_, g_f32, _, o_f32, _, _ = run_both(rd, po, sm, params, iq, blocks_per_call=53)
for c in range(nc):
    print("dnr level", params[c].nr_level, ["%.1e" % rel_rms(g_f32[b:b + win, c, :, 0], o_f32[b:b + win, c, :, 0]) for b in range(0, nb - win, win)])
