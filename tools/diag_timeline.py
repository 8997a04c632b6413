#!/usr/bin/env python3
"""When does every kernel of one process_blocks() call run inside the wavefront?  (GPU box; RDSP_TIMELINE=1 makes the
library bracket every launch with events and print start / end relative to the call's fork event.)"""
import os, sys
os.environ["RDSP_TIMELINE"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import radiodsp_sdr_rx_b200 as rd

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 0
T_ARG = int(sys.argv[3]) if len(sys.argv) > 3 else 8
desc, stage, C_ = bench.WORKLOADS[wl]
T = T_ARG
dev = torch.device("cuda", 0)
iq = bench.make_inputs(wl, 0, C_, T)
d_in = torch.from_numpy(iq).to(dev)
d_out = torch.zeros((T, C_, 128, 2), dtype=torch.int16, device=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
cfg = rd.default_config(n_channels=C_, device=0, stage_mask=stage, max_blocks_per_call=T, io_location=rd.IO_DEVICE, pipeline_chunks=chunks)
cfg.async_ = 1
b = rd.ReceiverBank(cfg)
for c in range(C_):
    b.set_mode(c, 1, rd.default_params(**bench.channel_params(wl, c)))
b.set_stream(stream.cuda_stream)
for i in range(4):
    print(f"--- call {i}", file=sys.stderr, flush=True)
    b.process_blocks(T, d_in, d_out)
torch.cuda.synchronize()
