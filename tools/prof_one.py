"""a few process calls of one workload, for ncu captures: python tools/prof_one.py <workload> [channels] [calls]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, radiodsp_sdr_rx_b200 as rd
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
C_ = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[wl][2]
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 3
T = 8
iq = torch.from_numpy(bench.make_inputs(wl, 0, C_, T)).cuda()
out = torch.zeros_like(iq)
cfg = rd.default_config(n_channels=C_, stage_mask=bench.WORKLOADS[wl][1], max_blocks_per_call=T, io_location=rd.IO_DEVICE, graph_mode=rd.GRAPH_OFF)
b = rd.ReceiverBank(cfg)
for c in range(C_):
    b.set_mode(c, 1, rd.default_params(**bench.channel_params(wl, c)))
for _ in range(calls):
    b.process_blocks(T, iq, out)
torch.cuda.synchronize()
print("done", wl, C_, calls, b.kernel_launches)
