#!/usr/bin/env python3
"""How long does the HOST take to enqueue one process_blocks() call, against the device time of the call?
(GPU box; development diagnostic for the launch-bound question.)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import radiodsp_sdr_rx_b200 as rd

wl, C_, T = "cfg5", 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
iq = bench.make_inputs(wl, 0, C_, T)
d_in = torch.from_numpy(iq).to(dev)
d_out = torch.zeros((T, C_, 128, 2), dtype=torch.int16, device=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
for chunks in ((1,) if len(sys.argv) > 1 else (1, 2, 4)):
    cfg = rd.default_config(n_channels=C_, device=0, stage_mask=0x7F, max_blocks_per_call=T, io_location=rd.IO_DEVICE, pipeline_chunks=chunks)
    cfg.async_ = 1
    b = rd.ReceiverBank(cfg)
    for c in range(C_):
        b.set_mode(c, 1, rd.default_params(**bench.channel_params(wl, c)))
    b.set_stream(stream.cuda_stream)
    for _ in range(5):
        b.process_blocks(T, d_in, d_out)
    torch.cuda.synchronize()
    n = 50
    # (a) host enqueue time with an idle queue in front (sync before every call)
    host = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); b.process_blocks(T, d_in, d_out); host.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    # (b) back-to-back: device time per call
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(n):
        b.process_blocks(T, d_in, d_out)
    e1.record(stream)
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"chunks {chunks}: host enqueue {1e3*np.median(host):.3f} ms/call (idle queue), {1e3*t_enq/n:.3f} ms/call back to back; "
          f"device {e0.elapsed_time(e1)/n:.3f} ms/call; launches/call {b.kernel_launches // (n * 2 + 5)}")
    del b
