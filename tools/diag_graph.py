"""device-resident step time with and without CUDA graphs, T = 1 and T = 8 (cfg5, 8192 channels); host time per call"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, radiodsp_sdr_rx_b200 as rd
C_ = 8192
for T in (1, 8):
    iq = torch.from_numpy(bench.make_inputs("cfg5", 0, C_, 4 * T)).cuda().view(4, T, C_, 128, 2)
    out = torch.zeros((T, C_, 128, 2), dtype=torch.int16, device="cuda")
    for mode in (rd.GRAPH_OFF, rd.GRAPH_AUTO):
        cfg = rd.default_config(n_channels=C_, stage_mask=0x7F, max_blocks_per_call=T, io_location=rd.IO_DEVICE, graph_mode=mode)
        cfg.async_ = 1
        b = rd.ReceiverBank(cfg)
        for c in range(C_):
            b.set_mode(c, 1, rd.default_params(**bench.channel_params("cfg5", c)))
        s = torch.cuda.Stream(); torch.cuda.set_stream(s); b.set_stream(s.cuda_stream)
        t0 = time.perf_counter()
        for i in range(16):
            b.process_blocks(T, iq[i % 4], out)
        torch.cuda.synchronize()
        t_prime = time.perf_counter() - t0
        n = 200
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        h0 = time.perf_counter()
        for i in range(n):
            b.process_blocks(T, iq[i % 4], out)
        h1 = time.perf_counter()
        e1.record(s); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"T={T} graph={'auto' if mode == rd.GRAPH_AUTO else 'off'}: {ms * 1e3:.1f} us/step = {C_ * T * 128 / ms / 1e3:.0f} MS/s; host {1e6 * (h1 - h0) / n:.1f} us/call; "
              f"16 priming calls {t_prime * 1e3:.1f} ms; replays {b.graph_replays}", flush=True)
