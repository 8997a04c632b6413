"""one block per call through host buffers: time per call and graph replays, stereo / mono (GPU box diagnostic)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, radiodsp_sdr_rx_b200 as rd
C_, T, N = 8192, 1, 400
iq = torch.from_numpy(bench.make_inputs("cfg5", 0, C_, 4)).view(4, T, C_, 128, 2).pin_memory()
for layout in (rd.AUDIO_STEREO, rd.AUDIO_MONO, rd.AUDIO_STEREO, rd.AUDIO_MONO):
    cfg = rd.default_config(n_channels=C_, stage_mask=rd.STAGE_ALL, max_blocks_per_call=T, io_location=rd.IO_HOST, audio_layout=layout)
    cfg.async_ = 1
    b = rd.ReceiverBank(cfg)
    for c in range(C_):
        b.set_mode(c, 1, rd.default_params(**bench.channel_params("cfg5", c)))
    out = torch.zeros((2, T, C_, 128) if layout == rd.AUDIO_MONO else (2, T, C_, 128, 2), dtype=torch.int16).pin_memory()
    for i in range(24):
        b.process_blocks(T, iq[i % 4], out[i % 2])
    b.synchronize()
    r0 = b.graph_replays
    t0 = time.perf_counter()
    for i in range(N):
        b.process_blocks(T, iq[i % 4], out[i % 2])
    t1 = time.perf_counter()
    b.synchronize()
    t2 = time.perf_counter()
    print("mono" if layout == rd.AUDIO_MONO else "stereo", "PDL_MAX_T", os.environ.get("RDSP_PDL_MAX_T"), "us per call: enqueue %.1f, total %.1f" % ((t1 - t0) / N * 1e6, (t2 - t0) / N * 1e6),
          "graph replays", b.graph_replays - r0, "of", N, flush=True)
    b.close()
