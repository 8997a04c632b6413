#!/usr/bin/env python3
"""Pinned-memory copy bandwidth of the box, one way and both ways at once (GPU box; the ceiling of the end-to-end metric)."""
import torch
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    e[0].record(s1); e[2].record(s2)
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    e[1].record(s1); e[3].record(s2)
    torch.cuda.synchronize()
    return (reps * n / e[0].elapsed_time(e[1]) / 1e6 if h2d else 0, reps * n / e[2].elapsed_time(e[3]) / 1e6 if d2h else 0)
run(True, True, 2)
print("H2D alone  %.1f GB/s" % run(True, False)[0])
print("D2H alone  %.1f GB/s" % run(False, True)[1])
print("both       H2D %.1f GB/s, D2H %.1f GB/s" % run(True, True))
# the sizes one call of the bank moves (8192 / 4096 channels x 8 blocks), back to back on each stream
for mb in (32, 16, 4):
    n = mb << 20
    h_in, h_out, d_in, d_out = h_in[:n], h_out[:n], d_in[:n], d_out[:n]
    print("both, %2d MiB copies: H2D %.1f GB/s, D2H %.1f GB/s" % ((mb,) + run(True, True, 40)))
