import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radiodsp_sdr_rx_b200 as rd
nc, nblocks = 1200, 12
rng = np.random.default_rng(11)
taps = {(k, i): rng.integers(-32768, 32768, 129).astype(np.int16) for k in range(3) for i in range(5)}
taps[(0, 1)] = taps[(0, 0)].copy(); taps[(1, 1)] = taps[(1, 0)].copy()
for i in range(5):
    taps[(2, i)] = (taps[(2, i)] // (1 << i)).astype(np.int16)
demod = rng.integers(0, 5, nc); filt = rng.integers(0, 5, nc); gains = rng.choice([1.0, 0.37, 2.5, 1.02], nc)
iq = rng.integers(-32768, 32768, (nblocks, nc, 128, 2)).astype(np.int16)
iq[:, ::7] = np.where(rng.random((nblocks, (nc + 6) // 7, 128, 2)) < 0.5, 32767, -32768).astype(np.int16)
print("pairable rows:", {k: bool((t < 32640).all()) for k, t in taps.items()})
def run(impl):
    os.environ.pop("RDSP_FRONT_IMPL", None)
    if impl: os.environ["RDSP_FRONT_IMPL"] = impl
    cfg = rd.default_config(n_channels=nc, stage_mask=rd.STAGE_FRONTEND, max_blocks_per_call=nblocks, io_location=rd.IO_HOST, debug_f32=int(os.environ.get("DBG", "0")))
    bank = rd.ReceiverBank(cfg)
    for (k, i), t in taps.items():
        bank.set_taps(k, i, t)
    for c in range(nc):
        bank.set_mode(c, 1, rd.default_params(demod=int(demod[c]), audio_filter=int(filt[c]), in_gain=float(gains[c])))
    return [bank.process_host(iq), bank.process_host(iq[:6]), bank.process_host(iq[:1])]
a = run(None); b = run("cuda-core")
for k, (x, y) in enumerate(zip(a, b)):
    d = (x != y)
    bad_ch = np.unique(np.nonzero(d)[1])
    print(f"call {k}: mismatching samples {int(d.sum())}, channels {bad_ch.size}", "demod/filt of bad:", sorted({(int(demod[c]), int(filt[c])) for c in bad_ch})[:12])
    if d.any():
        bl = np.unique(np.nonzero(d)[0]); sm = np.unique(np.nonzero(d)[2])
        print("   blocks", bl.tolist(), "samples", sm[:20].tolist(), "...", sm[-5:].tolist(), "maxdiff", int(np.abs(x.astype(np.int32) - y).max()))
        c0 = int(bad_ch[0]); w = np.nonzero(d[:, c0])
        print("   first bad channel", c0, "demod", int(demod[c0]), "filt", int(filt[c0]), "gain", float(gains[c0]), "first (block, sample):", int(w[0][0]), int(w[1][0]), "tc", x[w[0][0], c0, w[1][0]].tolist(), "cc", y[w[0][0], c0, w[1][0]].tolist())
        print("   bad channels head", bad_ch[:30].tolist(), "count by (demod,filt):", {k: int(sum(1 for c in bad_ch if (int(demod[c]), int(filt[c])) == k)) for k in sorted({(int(demod[c]), int(filt[c])) for c in bad_ch})})
