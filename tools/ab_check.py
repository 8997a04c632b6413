#!/usr/bin/env python3
"""sha256 of the audio of a few cfg5 calls with the library RDSP_GPU_LIB points to (A/B of two builds; GPU box)."""
import hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import radiodsp_sdr_rx_b200 as rd
wl, C_, T = sys.argv[1] if len(sys.argv) > 1 else "cfg5", 2048, 8
iq = bench.make_inputs(wl, 0, C_, 3 * T)
cfg = rd.default_config(n_channels=C_, device=0, stage_mask=bench.WORKLOADS[wl][1], max_blocks_per_call=T, io_location=rd.IO_HOST)
b = rd.ReceiverBank(cfg)
for c in range(C_):
    b.set_mode(c, 1, rd.default_params(**bench.channel_params(wl, c)))
h = hashlib.sha256()
for k in range(3):
    out = b.process_host(iq[k * T:(k + 1) * T])
    h.update(np.ascontiguousarray(out).tobytes())
print(wl, os.environ.get("RDSP_GPU_LIB", "in-tree"), h.hexdigest()[:16])
