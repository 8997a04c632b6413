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
params = [po.default_params(demod=c % 5, nr_kind=po.NR_LMS, nr_level=30, notch_on=1, in_gain=4.0) for c in range(4)]
mid, _ = po.process_bank(po.default_config(stage_mask=po.STAGE_FRONTEND | po.STAGE_NOTCH | po.STAGE_AGC), params, iq)
print("mid peak per block ch2", [int(np.abs(mid[b, 2]).max()) for b in range(nb)])
for bpc in (3, 10, 1):
    for lvl in (30,):
        g_out, g_f32, o_out, o_f32, bank, chans = run_both(rd, po, rd.STAGE_FFTFILT | rd.STAGE_NR, params, mid, blocks_per_call=bpc)
        d = np.abs(g_out.astype(np.int32) - o_out)
        for c in (2, 3):
            print("bpc", bpc, "ch", c, "maxdiff per block", [int(d[b, c].max()) for b in range(nb)], "relrms f32 per block", ["%.1e" % rel_rms(g_f32[b, c], o_f32[b, c]) for b in range(nb)])
# NR off for reference
p2 = [p.copy(nr_kind=0, nr_level=0) for p in params]
g_out, g_f32, o_out, o_f32, bank, chans = run_both(rd, po, rd.STAGE_FFTFILT | rd.STAGE_NR, p2, mid, blocks_per_call=3)
print("NR off maxdiff", np.abs(g_out.astype(np.int32) - o_out).max(), "fft out peak per block ch2", [int(np.abs(o_out[b, 2]).max()) for b in range(nb)])
