import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import bench, pyoracle as po, radiodsp_sdr_rx_b200 as rd
from radiodsp_sdr_rx_b200 import synth
nc, nb = 4, 16
iq = synth.synth_iq(np.arange(50, 50 + nc), nb, [c % 5 for c in range(nc)])
for sm, name in ((rd.STAGE_FRONTEND | rd.STAGE_SPEC1024, "fe+spec1024"), (rd.STAGE_FRONTEND | rd.STAGE_NOTCH, "fe+notch")):
    cfg = rd.default_config(n_channels=nc, stage_mask=sm, max_blocks_per_call=64, io_location=rd.IO_HOST, debug_f32=1)
    bank = rd.ReceiverBank(cfg)
    chans = []
    for c in range(nc):
        p = dict(demod=c % 5, notch_on=1)
        bank.set_mode(c, 1, rd.default_params(**p)); chans.append(po.OracleChan(po.default_config(stage_mask=sm), po.default_params(**p)))
    for b0 in range(nb):
        g = bank.process_host(iq[b0:b0 + 1])
        o = np.stack([ch.process(iq[b0:b0 + 1, c]) for c, ch in enumerate(chans)], axis=1)
        d = np.abs(g.astype(np.int32) - o).max()
        line = f"{name} graph={os.environ.get('RDSP_GRAPH','1')} call {b0}: audio maxdiff {d}"
        if sm & rd.STAGE_SPEC1024:
            spec, ready = bank.read_audio_spectrum()
            os_ = [ch.read_audio_spectrum() for ch in chans]
            line += f" ready gpu {ready.tolist()} oracle {[int(r) for _, r in os_]} spec equal {[bool(np.array_equal(spec[c], os_[c][0])) for c in range(nc)]}"
        print(line)
