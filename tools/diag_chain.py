import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pyoracle as po, radiodsp_sdr_rx_b200 as rd
from radiodsp_sdr_rx_b200 import synth
from test_gpu_parity import _all_mode_params, run_both, rel_rms
nc, nb = 40, 96
params, demod = _all_mode_params(po, nc)
iq = synth.synth_iq(np.arange(nc), nb, demod, interferer=[d == po.DEMOD_CW_LSB for d in demod])
g_out, g_f32, o_out, o_f32, bank, chans = run_both(rd, po, rd.STAGE_ALL, params, iq, blocks_per_call=8)
d = np.abs(g_out.astype(np.int32) - o_out)
for c in range(nc):
    p = params[c]
    print(c, "demod", p.demod, "agc", p.agc_mode, "notch", p.notch_on, "nr", p.nr_level, "maxd", d[:, c].max(), "frac", round(float((d[:, c] > 0).mean()), 5),
          "relrms_f32", "%.2e" % rel_rms(g_f32[16:, c], o_f32[16:, c]), "rms_out", round(float(o_out[16:, c].std()), 1),
          "snr", round(synth.snr_db(o_out[16:, c, :, 0], g_out[16:, c, :, 0]), 1))
