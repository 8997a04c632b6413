#!/usr/bin/env python3
"""bench.py — throughput of the batched receive chain on N B200s, one process per GPU.

  python bench.py --gpus 1 --steps K --warmup W                    (this framework)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...                             (the CPU chain on the host cores)

A "step" = one rdsp_gpu_process_blocks() pass of the whole hot path over one batch: `blocks_per_call`
128-sample blocks of every channel of the rank.  Channels are independent, so ranks own contiguous channel
ranges and exchange nothing (weak scaling: 8192 channels per GPU; N = 8 is configs[4] of BASELINE.json,
65,536 all-mode channels).  Rank 0 prints ONE JSON line.

  value      aggregate MS/s (complex input samples per second over all ranks), inputs resident in HBM
  e2e        the same through the C-ABI call with pinned HOST buffers, H2D and D2H inside the timed region
  roofline   the dominant kernel: algorithmic bytes per launch / its CUDA-event duration, against the
             measured HBM copy bandwidth (MEASURED_PEAKS.json); the path is FP32/INT32-pipe bound, so the
             pipe fraction is reported next to it (DESIGN.md "Rooflines")
  cpu_baseline  the CPU oracle port on the host cores, bounded sample (rank 0, N = 1)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLK = 128
FS = 44100.0
CH_PER_GPU = 8192

# ---------------------------------------------------------------------------------------------
# workloads = configs of BASELINE.json (per-GPU slice)
# ---------------------------------------------------------------------------------------------
S_FE, S_NOTCH, S_AGC, S_FF, S_NR, S_S256, S_S1024 = (1 << i for i in range(7))

WORKLOADS = {
    # name: (description, stage mask, default channels per GPU)
    "cfg2": ("USB/LSB Hilbert-FIR SSB demod + 2.7 kHz band-pass", S_FE, 4096),
    "cfg3": ("CW, 500 Hz FIR + LMS auto-notch", S_FE | S_NOTCH, 16384),
    "cfg4a": ("SSB + AGC + FFT-256 filter + NLMS DNR level 30", S_FE | S_AGC | S_FF | S_NR, 8192),
    "cfg4b": ("SSB + AGC + FFT-256 framing + spectral-subtraction NR level 2", S_FE | S_AGC | S_FF | S_NR, 8192),
    "cfg5": ("full all-mode chain (AM/SSB/CW by channel, notch + DNR + AGC) + 256-pt IQ and 1024-pt audio spectra",
             0x7F, 8192),
}


def channel_params(workload: str, ch: int) -> dict:
    """per-channel parameters as SURVEY.md 8d defines the configs (ch = absolute channel id)"""
    if workload == "cfg2":
        return dict(demod=ch % 2, audio_filter=2)
    if workload == "cfg3":
        return dict(demod=2, audio_filter=0, notch_on=1)
    if workload == "cfg4a":
        return dict(demod=ch % 2, audio_filter=2, agc_mode=2, nr_kind=1, nr_level=30)
    if workload == "cfg4b":
        return dict(demod=ch % 2, audio_filter=2, agc_mode=2, nr_kind=2, nr_level=2)
    demod = (0, 1, 2, 4)[ch % 4]                       # LSB, USB, CW, AM
    lvl = (0, 20, 30, 40, 50)[ch % 5]
    return dict(demod=demod, audio_filter={2: 0, 4: 4}.get(demod, 2), agc_mode=ch % 4, notch_on=int(demod == 2),
                nr_kind=1 if lvl else 0, nr_level=lvl)


def algorithmic_bytes(workload: str, T: int) -> dict:
    """Algorithmic HBM bytes per channel-block, per kernel and for the whole step (SURVEY.md 8d figures with
    N_h = N_b = 129; state makes one round trip per call, so it is amortised over the T blocks of a call).
    Fractions = share of channels that run the kernel in this workload."""
    st = lambda b: 2.0 * b / T                        # state read + write
    k = {}
    if workload in ("cfg2", "cfg3", "cfg4a", "cfg4b", "cfg5"):
        k["k_front"] = 512 + 256 + st(768)
    if workload == "cfg3":
        k["k_nlms_notch"] = 256 + 512 + st(1288)
        k["k_agc"] = 512 + 512 + 0.0
    if workload in ("cfg4a", "cfg4b"):
        k["k_agc"] = 256 + 256 + st(16)
        k["k_fftfilt"] = 256 + 512 + st(512)
    if workload == "cfg4a":
        k["k_nlms_dnr"] = 512 + 512 + st(1288)
    if workload == "cfg5":
        k["k_nlms_notch"] = 0.25 * (256 + 512 + st(1288))
        k["k_agc"] = 256 + 256 + st(16)
        k["k_fftfilt"] = 256 + 512 + st(512)
        k["k_nlms_dnr"] = 0.8 * (512 + 512 + st(1288))
        k["k_biquad"] = 512 + 512 + st(32)
        k["k_spec256"] = 512 + st(1540) + 512.0 / 30
        k["k_spec1024"] = 256 + 256 + 1792.0 / 4 + 1024.0 / 4
    # whole fused-ideal step (no intermediates), SURVEY.md 8d totals
    total1 = {"cfg2": 2560, "cfg3": 5136, "cfg4a": 6192, "cfg4b": 6192 - 2 * 1288 + 8, "cfg5": 14169}[workload]
    k["_step"] = 1024 + (total1 - 1024) / T
    return k


# lane-operations (FP32 FMA / INT32 IMAD class) per channel-block, for the pipe fraction
PIPE_OPS = {"k_front": 3 * 129 * 128, "k_nlms_notch": 128 * (96 * 2 + 8), "k_nlms_dnr": 128 * (96 * 2 + 8),
            "k_fftfilt": 2 * 256 * 8 * 2.5 + 256 * 4, "k_agc": 128 * 8, "k_biquad": 128 * 2 * 12, "k_spec256": 4 * 64 * 40 + 256 * 8,
            "k_spec1024": (5 * 256 * 40 + 1024 * 4) / 4.0}


# ---------------------------------------------------------------------------------------------
def ncu_traffic(kernel: str, workload: str, C_: int, T: int):
    """dram bytes per launch of `kernel` from the committed ncu --set full capture (profiles/), valid for the
    configuration it was taken on (cfg5, 8192 channels, 8 blocks per launch); None otherwise."""
    if workload != "cfg5" or C_ != 8192 or T != 8:
        return None
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files:
        return None
    k = json.load(open(files[-1]))["kernels"]
    # the capture is of the un-profiled step, where the audio chain runs as two channel classes: the bytes of one
    # bench "launch" (all channels of a stage) are the sum over the launches of that stage
    def rows_of(prefix):
        return [r for name, rows in k.items() if name.split("<")[0] == prefix for r in rows]
    if kernel in ("k_nlms_notch", "k_nlms_dnr"):
        # three k_nlms launches per step: the notch of the CW channels (a quarter of the channels, 16 per CTA) and the DNR of the
        # two channel classes; the notch is the one whose grid matches its list
        rows = rows_of("k_nlms")
        notch_grid = (C_ // 4 + 15) // 16
        notch = [r for r in rows if r["grid"] == notch_grid][:1]
        rows = notch if kernel == "k_nlms_notch" else [r for r in rows if r["grid"] != notch_grid]
        if not rows:
            return None
    elif kernel == "k_front":
        rows = rows_of("k_front_tc") or rows_of("k_front")
    elif kernel == "k_agc":
        rows = [r for name, rr in k.items() if name.startswith("k_agc<0>") for r in rr]
    else:
        rows = rows_of(kernel)
    if not rows:
        return None
    return float(sum(r["dram_bytes"] for r in rows))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        if os.environ.get("RDSP_BENCH_NO_CLOCKS"):     # experiments only: rule the sampler out as a disturbance
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t0 - 0.02 <= t <= t1 + 0.05]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}

    def n_rows(self) -> int:
        return len(self.rows)


def make_inputs(workload: str, ch0: int, C_: int, n_blocks: int, unique: int = 512) -> np.ndarray:
    """synthetic 40 m IQ for channels [ch0, ch0+C): `unique` distinct channels generated (seeded by absolute id),
    tiled over the range with the right mode for each slot (mode period 4 divides `unique`)."""
    from radiodsp_sdr_rx_b200 import synth
    u = min(unique, C_)
    ids = ch0 + np.arange(u)
    demod = [channel_params(workload, int(c))["demod"] for c in ids]
    het = [bool(channel_params(workload, int(c)).get("notch_on", 0)) for c in ids]
    base = synth.synth_iq(ids, n_blocks, demod, interferer=het)               # [nb,u,128,2]
    reps = (C_ + u - 1) // u
    return np.ascontiguousarray(np.tile(base, (1, reps, 1, 1))[:, :C_])


def rank_channel_range(rank: int, channels_per_gpu: int):
    """contiguous channel range owned by a rank (weak scaling): [first, first + count)"""
    return rank * channels_per_gpu, channels_per_gpu


def max_over_ranks(ms: float, dist=None, device=None) -> float:
    """device time of the slowest rank (the job finishes when the last rank does)"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(ms)
    import torch
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---------------------------------------------------------------------------------------------
def _timing_oracle():
    """pyoracle bound to a build of the CPU port for THIS host: oracle/liboracle_native.so (-O3 -march=native, built here on
    first use) — parity tests keep the portable x86-64-v3 build.  Returns (module, build name)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    native = os.path.join(ROOT, "oracle", "liboracle_native.so")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "native"], capture_output=True, text=True)
    name = "-O3 -march=native"
    if r.returncode == 0 and os.path.exists(native):
        os.environ["RDSP_ORACLE_LIB"] = native
    else:
        name = "-O3 -march=x86-64-v3 (native build failed)"
    import pyoracle as po
    return po, name


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_chain(workload: str, n_channels: int, T: int, n_steps: int, n_warm: int = 1):
    """The CPU oracle port on ALL host threads, the GPU arm's own shape: every step is one pass of `T` blocks over
    `n_channels` channels (contiguous channel ranges per thread, ctypes releases the GIL; the state of all channels is
    live, so it does not sit in L1 / L2 the way a one-channel loop would).  Returns (ms per step list, build name, cores)."""
    po, build = _timing_oracle()
    stage = WORKLOADS[workload][1]
    iq = make_inputs(workload, 0, n_channels, T)
    out = np.zeros_like(iq)
    cfg = po.default_config(stage_mask=stage)
    chans = [po.OracleChan(cfg, po.default_params(**channel_params(workload, c))) for c in range(n_channels)]
    arr = (C.c_void_p * n_channels)(*[c.handle for c in chans])
    L = po.lib()
    cores = host_cores()
    bounds = [n_channels * i // cores for i in range(cores + 1)]

    def work(i):
        if bounds[i + 1] > bounds[i]:
            L.rdsp_oracle_bank_process(arr, bounds[i], bounds[i + 1] - bounds[i], n_channels, T, iq.ctypes.data, out.ctypes.data)

    ms = []
    for k in range(n_warm + n_steps):
        th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        if k >= n_warm:
            ms.append((time.perf_counter() - t0) * 1e3)
    return ms, build, cores


def cpu_reference_stages(n_blocks: int = 16384):
    """kind "reference": the stages whose sources ARE in the reference tree — doConvolutionalProcessing (K5 + K6 + K7, DNR
    level 30) and AudioAnalyzeFFT256IQ::update (K9) — compiled unmodified (oracle/_ref, built in the build container: the
    reference tree does not travel), one private copy of the library per host thread because its state is file-scope
    globals.  Returns a cpu_baseline-style dict, or None when oracle/_ref is absent."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    if not po.ref_available():
        return None
    from radiodsp_sdr_rx_b200 import synth
    cores = host_cores()
    iq = synth.synth_iq(np.arange(cores), n_blocks, 0)
    refs = [po.RefChannel() for _ in range(cores)]
    ins = [np.ascontiguousarray(iq[:, i]) for i in range(cores)]
    for r, x in zip(refs, ins):
        r.run_blocks(x[:64], 30)                                  # warm the pages
    reps = 8                                                      # about 3 s: the state of a channel simply continues

    def work(r, x):
        for _ in range(reps):
            r.run_blocks(x, 30)

    th = [threading.Thread(target=work, args=(r, x)) for r, x in zip(refs, ins)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    return {"value": cores * reps * n_blocks * BLK / dt / 1e6, "unit": "MS/s", "cores": cores, "kind": "reference",
            "stages": "K5 + K6 (DNR level 30) + K7 + K9 only: the stages whose sources are in the reference tree (AudioSDR is not)",
            "sample": f"{cores} threads x 1 channel x {reps * n_blocks} blocks ({dt:.1f} s), reference sources compiled unmodified "
                      "(-O3 -march=x86-64-v3, built where /root/reference exists)"}


def workload_config(wl: str, C_: int, world: int, T: int) -> dict:
    """what both arms (this framework and --impl reference) are measured on: identical in both JSON lines"""
    return {"workload": f"{wl}: {WORKLOADS[wl][0]}", "channels_per_gpu": C_, "channels_total": world * C_, "blocks_per_call": T,
            "block_samples": BLK, "sample_rate_hz": FS, "sharding": "contiguous channel ranges, no collective on the hot path",
            "l2": "GPU arm: 256 MiB memset between timed steps; per-step CUDA events summed"}


def run_reference(args):
    """--impl reference: the reference chain on the host cores, on the GPU arm's own config — every step is one pass of
    blocks_per_call blocks over ALL channels_total channels (CPU oracle port; its in-tree stages are pinned bit-exact to the
    reference's own sources compiled unmodified, oracle/_ref; AudioSDR does not exist in the reference tree, so the whole
    chain cannot be kind "reference").  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    world = max(1, args.gpus)
    T = args.blocks_per_call
    C_ = args.channels or WORKLOADS[wl][2]
    ms, build, cores = cpu_chain(wl, world * C_, T, args.steps, n_warm=max(1, min(args.warmup, 2)))
    samples = world * C_ * T * BLK
    v = samples * len(ms) / (sum(ms) * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": "aggregate MS/s (real-time 44.1 kS/s channels sustained = value / 0.0441)",
        "value": v, "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "q15+f32", "data": "synthetic",
        "config": workload_config(wl, C_, world, T),
        "derived": {"realtime_channels": v / 0.0441},
        "cpu_baseline": {"value": v, "unit": "MS/s", "cores": cores, "kind": "port",
                         "sample": f"every step: {cores} threads x {world * C_} channels x {T} blocks of the {wl} chain (the GPU arm's step); "
                                   f"oracle port built {build}"},
        "cpu_baseline_reference": cpu_reference_stages(),
        "e2e": {"value": v, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
class Gpu:
    """one rank's device context: stream, barrier, max-over-ranks"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — this framework has no CPU fallback (use --impl reference for the CPU chain)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # NCCL's banner ("NCCL version ...") goes to stdout by default: keep stdout to the ONE JSON line
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=self.dev)
        # a real (non-default) stream: the C ABI takes NULL to mean "the handle's own stream", and CUDA events must be
        # recorded on the stream the kernels are launched on
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        assert self.stream.cuda_stream != 0
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)      # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local])
        self.torch.cuda.synchronize()

    def max_ms(self, ms: float) -> float:
        return max_over_ranks(ms, self.dist if self.world > 1 else None, self.dev)


def new_bank(gpu, rd, wl, C_, T, io, ch0, layout=0, pipeline_chunks=0):
    cfg = rd.default_config(n_channels=C_, device=gpu.local, stage_mask=WORKLOADS[wl][1], max_blocks_per_call=T, io_location=io,
                            pipeline_chunks=pipeline_chunks, audio_layout=layout)
    cfg.async_ = 1
    b = rd.ReceiverBank(cfg)
    for c in range(C_):                               # the library ignores repeats and de-duplicates masks
        b.set_mode(c, 1, rd.default_params(**channel_params(wl, ch0 + c)))
    b.set_stream(gpu.stream.cuda_stream)
    return b


def measure_device(gpu, rd, wl, C_, T, K, W, NB, iq_host, sample_clocks=False, pipeline_chunks=0):
    """device-resident throughput of one workload + per-kernel device times.  Inputs already in HBM, L2 flushed between
    timed steps (outside the per-step events), CUDA events on the launching stream, max over ranks."""
    torch = gpu.torch
    ch0, _ = rank_channel_range(gpu.rank, C_)
    d_in = torch.from_numpy(iq_host).to(gpu.dev).view(NB, T, C_, BLK, 2)
    d_out = torch.zeros((T, C_, BLK, 2), dtype=torch.int16, device=gpu.dev)
    bank = new_bank(gpu, rd, wl, C_, T, rd.IO_DEVICE, ch0, pipeline_chunks=pipeline_chunks)

    def step(i):
        bank.process_blocks(T, d_in[i % NB], d_out)

    sampler = ClockSampler(gpu.local) if sample_clocks else None     # nvidia-smi takes a moment to start: launch it before the warm-up
    # initialisation, before the warm-up: the library captures a call shape (buffers x ping-pong phase) into a CUDA graph
    # the second time it sees it; let every shape of the rotation be seen twice so that the W warm-up steps and the K
    # timed steps all run the way steady state does (one cudaGraphLaunch per call)
    for i in range(4 * NB):
        step(i)
    torch.cuda.synchronize()
    for i in range(W):
        step(i)
    gpu.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0, replays0 = bank.kernel_launches, bank.graph_replays
    t_wall0 = time.time()
    no_flush = bool(os.environ.get("RDSP_BENCH_NO_FLUSH"))      # experiments only (the reported numbers always flush)
    for i in range(K):
        if not no_flush:
            gpu.flush.zero_()                          # L2 flush between timed iterations (outside the per-step events)
        ev[i][0].record(gpu.stream)
        step(W + i)
        ev[i][1].record(gpu.stream)
    gpu.barrier()
    t_wall1 = time.time()
    res = {"launches": bank.kernel_launches - launches0, "graph_replays": bank.graph_replays - replays0}
    if sampler is not None:
        # the timed region of a short run can fall between two nvidia-smi samples (200 ms period; faster polling measurably
        # slows kernel launches): keep the SAME load up, untimed, until at least three samples were taken under it, and
        # report over [start of the timed region, end of load]
        t_load1, extended = t_wall1, 0
        if sampler.proc is not None:
            deadline = time.time() + 3.0
            i = 0
            while time.time() < deadline:
                if sum(1 for t, _ in sampler.rows if t_wall0 <= t <= time.time()) >= 3:
                    break
                step(W + K + i); i += 1
                if i % 8 == 0:
                    torch.cuda.synchronize()
            torch.cuda.synchronize()
            extended = i
            if extended:
                t_load1 = time.time()
        clocks = sampler.stop(t_wall0, t_load1)
        clocks["window"] = "timed region" if not extended else f"timed region + {extended} more of the same steps, untimed, until 3 samples"
        res["clocks"] = clocks
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_total_max = gpu.max_ms(float(sum(ms_steps)))
    res.update(ms_steps=ms_steps, ms_per_step=ms_total_max / K,
               value=gpu.world * C_ * T * BLK * K / (ms_total_max * 1e-3) / 1e6)
    # per-kernel device times (CUDA events on the launching stream, same steps, L2 flushed, kernel by kernel)
    bank.profile(True)
    for i in range(K):
        gpu.flush.zero_()
        step(W + K + i)
    torch.cuda.synchronize()
    res["prof"] = {k: v for k, v in bank.profile_read().items() if v["launches"] > 0}
    bank.profile(False)
    bank.close()
    return res


E2E_MIN_STEPS = 100


def measure_e2e(gpu, rd, wl, C_, T, K, W, NB, iq_host, layout):
    """the same metric through the C-ABI call with pinned HOST buffers: H2D of the step's input and D2H of its audio inside
    the timed region, every step.  The call is a three-phase pipeline (copy in, kernels, copy out) over staging buffers, and
    the timed region runs from the first byte in to the last byte out, so it contains one fill and one drain: with 8 blocks
    per call that is 1.1 ms next to 0.70 ms per step (r02b: 20 steps = 15.2 ms = 20 x 0.70 + 1.1 — 8 % of a 20-step window is
    ramp).  A receiver runs for hours, so the end-to-end figure is taken over at least E2E_MIN_STEPS steps (`e2e.steps` on the
    line; the copy-only ceiling runs over the same count)."""
    torch = gpu.torch
    K = max(K, E2E_MIN_STEPS)
    ch0, _ = rank_channel_range(gpu.rank, C_)
    bank = new_bank(gpu, rd, wl, C_, T, rd.IO_HOST, ch0, layout=layout)
    h_in = torch.from_numpy(iq_host).view(NB, T, C_, BLK, 2).pin_memory()
    out_shape = (2, T, C_, BLK) if layout else (2, T, C_, BLK, 2)
    h_out = torch.zeros(out_shape, dtype=torch.int16).pin_memory()
    for i in range(max(W, 1) + 8):                     # (+ 8: both staging buffers x both phases seen twice: graphs captured)
        bank.process_blocks(T, h_in[i % NB], h_out[i % 2])
    bank.synchronize()
    gpu.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(gpu.stream)
    for i in range(K):
        bank.process_blocks(T, h_in[(W + i) % NB], h_out[i % 2])
    bank.stream_join()                                 # the last device-to-host copies are part of the timed region
    e1.record(gpu.stream)
    bank.synchronize()
    gpu.barrier()
    ms = gpu.max_ms(e0.elapsed_time(e1))
    in_bytes = T * C_ * BLK * 2 * 2
    out_bytes = in_bytes // 2 if layout else in_bytes
    bank.close()
    # copy-only ceiling: the same pinned buffers and byte counts, H2D and D2H concurrently on two streams, every rank at once
    d_a = torch.empty(in_bytes, dtype=torch.uint8, device=gpu.dev)
    d_b = torch.empty(out_bytes, dtype=torch.uint8, device=gpu.dev)
    s_in, s_out = torch.cuda.Stream(device=gpu.dev), torch.cuda.Stream(device=gpu.dev)
    hv_in = h_in.view(torch.uint8).view(NB, -1)
    hv_out = h_out.view(torch.uint8).view(2, -1)

    def copies(n):
        for i in range(n):
            with torch.cuda.stream(s_in):
                d_a.copy_(hv_in[i % NB], non_blocking=True)
            with torch.cuda.stream(s_out):
                hv_out[i % 2].copy_(d_b, non_blocking=True)

    copies(3)
    gpu.barrier()
    c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    c0.record(gpu.stream)
    s_in.wait_stream(gpu.stream); s_out.wait_stream(gpu.stream)
    copies(K)
    gpu.stream.wait_stream(s_in); gpu.stream.wait_stream(s_out)
    c1.record(gpu.stream)
    gpu.barrier()
    ceil_ms = gpu.max_ms(c0.elapsed_time(c1))
    samples = gpu.world * C_ * T * BLK * K
    v, ceil_v = samples / (ms * 1e-3) / 1e6, samples / (ceil_ms * 1e-3) / 1e6
    return {"value": v, "unit": "MS/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
            "audio_layout": "mono (L)" if layout else "stereo (L,R)", "steps": K,
            "ceiling": {"value": ceil_v, "unit": "MS/s", "GB_s_each_way_all_ranks": [gpu.world * in_bytes * K / (ceil_ms * 1e-3) / 1e9, gpu.world * out_bytes * K / (ceil_ms * 1e-3) / 1e9],
                        "what": "copy only: the same pinned buffers and byte counts per step, H2D and D2H concurrently on two streams, every rank at once, max over ranks"},
            "frac_of_ceiling": v / ceil_v}


def kernel_table(wl, C_, T, K, prof):
    ab = algorithmic_bytes(wl, T)
    tot = max(sum(v["ms"] for v in prof.values()), 1e-12)
    kernels = {}
    for kname, v in prof.items():
        per_launch_ms = v["ms"] / v["launches"]
        bytes_launch = ab.get(kname, 0.0) * C_ * T
        kernels[kname] = {"ms_per_launch": per_launch_ms, "share": v["ms"] / tot,
                          "alg_gb_s": bytes_launch / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else None}
    return kernels, ab, (max(prof, key=lambda k: prof[k]["ms"]) if prof else None)


def roofline_of(dom, wl, C_, T, prof, ab):
    """the dominant kernel against the HBM roofline the contract asks for, and against the pipe that actually binds it:
    the CUDA-core FMA pipe for the recurrences and FFTs, the int8 TENSOR pipe for the tcgen05 front end"""
    peak, peak_src, sm_max = measured_peaks()
    per_launch_ms = prof[dom]["ms"] / prof[dom]["launches"]
    bytes_launch = ab.get(dom, 0.0) * C_ * T
    achieved = bytes_launch / (per_launch_ms * 1e-3) / 1e9
    frac_ch = {"k_nlms_notch": 0.25 if wl == "cfg5" else 1.0, "k_nlms_dnr": 0.8 if wl == "cfg5" else 1.0}.get(dom, 1.0)
    if dom == "k_front":
        # 4 byte-plane products per q15 product, K padded 129 -> 160: int8 MACs issued per channel-block; dense int8 peak = 2 x the
        # measured bf16 burst (same tensor pipe, half the operand width)
        ops = 3 * 128 * 160 * 4 * 2.0
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        pipe_peak = 2.0 * float(p.get("bf16_tflops", 1590.0)) * 1e12
        pipe = {"pipe": "tensor (tcgen05 kind::i8)", "achieved_Tops": ops * C_ * T / (per_launch_ms * 1e-3) / 1e12, "peak_Tops": pipe_peak / 1e12,
                "frac": ops * C_ * T / (per_launch_ms * 1e-3) / pipe_peak, "peak_source": "2 x measured bf16 burst (MEASURED_PEAKS.json)"}
        binding = "tcgen05 kind::i8 tensor pipe at small N (per-instruction cost of N = 32 MMAs), not HBM — see DESIGN.md 4a"
    else:
        pipe_peak = 148 * 128 * sm_max * 1e6                     # lane-ops/s of the FP32/INT32 FMA pipe at max clock
        pipe_ach = PIPE_OPS.get(dom, 0) * frac_ch * C_ * T / (per_launch_ms * 1e-3)
        pipe = {"pipe": "fp32/int32 FMA (CUDA cores)", "achieved_Tlaneops": pipe_ach / 1e12, "peak_Tlaneops": pipe_peak / 1e12, "frac": pipe_ach / pipe_peak}
        binding = "fp32/int32 pipe and recurrence latency (sequential NLMS / AGC, q15 FFTs), not HBM — see DESIGN.md"
    return {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": ncu_traffic(dom, wl, C_, T), "peak_source": peak_src, "alg_bytes_per_channel_block": ab.get(dom),
            "ms_per_launch": per_launch_ms, "binding": binding, "pipe": pipe}


def run_b200(args):
    import radiodsp_sdr_rx_b200 as rd
    gpu = Gpu()
    torch, world, rank = gpu.torch, gpu.world, gpu.rank
    wl = args.workload
    C_ = args.channels or WORKLOADS[wl][2]
    T, K, W, NB = args.blocks_per_call, args.steps, args.warmup, args.input_batches
    ch0, _ = rank_channel_range(rank, C_)             # contiguous channel range of this rank
    iq_host = make_inputs(wl, ch0, C_, NB * T)        # [NB*T, C, 128, 2]

    main = measure_device(gpu, rd, wl, C_, T, K, W, NB, iq_host, sample_clocks=True, pipeline_chunks=args.pipeline_chunks)
    e2e = measure_e2e(gpu, rd, wl, C_, T, K, W, NB, iq_host, layout=0)
    e2e_mono = measure_e2e(gpu, rd, wl, C_, T, K, W, NB, iq_host, layout=1)

    # the other BASELINE configs, compact (N = 1 only: the scaling runs are about the default workload)
    others = {}
    if world == 1 and not args.no_other_configs:
        for o in sorted(WORKLOADS):
            if o == wl:
                continue
            oc = WORKLOADS[o][2]
            r = measure_device(gpu, rd, o, oc, T, max(5, min(K, 10)), max(3, min(W, 5)), 2, make_inputs(o, 0, oc, 2 * T))
            kern, ab_o, dom_o = kernel_table(o, oc, T, 0, r["prof"])
            peak = measured_peaks()[0]
            step_bytes = ab_o["_step"] * oc * T
            others[o] = {"channels_per_gpu": oc, "blocks_per_call": T, "value": r["value"], "unit": "MS/s", "ms_per_step": r["ms_per_step"],
                         "realtime_channels": r["value"] / 0.0441, "dominant_kernel": dom_o,
                         "dominant_ms_per_launch": kern[dom_o]["ms_per_launch"] if dom_o else None,
                         "dominant_hbm_frac": (kern[dom_o]["alg_gb_s"] or 0.0) / peak if dom_o else None,
                         "step_hbm_frac": step_bytes / (r["ms_per_step"] * 1e-3) / 1e9 / peak,
                         "kernels_us": {k: round(v["ms_per_launch"] * 1e3, 1) for k, v in kern.items()}}

    if rank == 0:
        peak, peak_src, _ = measured_peaks()
        prof = main["prof"]
        kernels, ab, dom = kernel_table(wl, C_, T, K, prof)
        value, ms_step = main["value"], main["ms_per_step"]
        step_bytes = ab["_step"] * C_ * T
        line = {
            "metric": "aggregate MS/s (real-time 44.1 kS/s channels sustained = value / 0.0441)",
            "value": value, "unit": "MS/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "q15+f32", "data": "synthetic",
            "ms_step_min_median_max": [float(np.min(main["ms_steps"])), float(np.median(main["ms_steps"])), float(np.max(main["ms_steps"]))],
            **({"ms_steps": [round(float(x), 3) for x in main["ms_steps"]]} if os.environ.get("RDSP_BENCH_STEPS") else {}),
            "config": workload_config(wl, C_, world, T),
            "derived": {"realtime_channels": value / 0.0441, "realtime_channels_e2e": e2e["value"] / 0.0441,
                        "realtime_margin_per_gpu": (value / world) / (C_ * 0.0441), "pipeline_chunks": args.pipeline_chunks or "library default (1 channel group)"},
            "e2e": e2e,
            "e2e_mono": e2e_mono,
            "gpu_launches": int(main["launches"]),
            "graph_replays": int(main["graph_replays"]),       # timed calls that ran as one cudaGraphLaunch (the rest: kernel by kernel)
            "clocks": main["clocks"],
            "roofline": roofline_of(dom, wl, C_, T, prof, ab) if dom else None,
            "roofline_step": {"bound": "hbm", "achieved": step_bytes / (ms_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak, "alg_bytes_per_channel_block": ab["_step"]},
            "kernels": kernels,
            "profiled_step_ms": sum(v["ms"] for v in prof.values()) / max(K, 1),
            "other_configs": others,
        }
        if world == 1 and not args.no_cpu:
            # a bounded sample of the SAME step shape (all channels x T blocks), about 10 - 20 s of CPU work
            ms1, build, cores = cpu_chain(wl, C_, T, 1, n_warm=1)
            n = int(max(3, min(60, 12e3 / max(ms1[0], 1.0))))
            ms, build, cores = cpu_chain(wl, C_, T, n, n_warm=0)
            v = C_ * T * BLK * len(ms) / (sum(ms) * 1e-3) / 1e6
            line["cpu_baseline"] = {"value": v, "unit": "MS/s", "cores": cores, "kind": "port",
                                    "sample": f"{len(ms)} steps of the GPU arm's shape: {cores} threads x {C_} channels x {T} blocks of the {wl} chain "
                                              f"({sum(ms) * 1e-3:.1f} s); oracle port built {build}"}
            line["cpu_baseline_reference"] = cpu_reference_stages()
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        gpu.dist.barrier(device_ids=[gpu.local])
        gpu.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=sorted(WORKLOADS))
    ap.add_argument("--channels", type=int, default=0, help="channels per GPU (default: the workload's)")
    ap.add_argument("--blocks-per-call", type=int, default=8)
    ap.add_argument("--input-batches", type=int, default=4)
    ap.add_argument("--pipeline-chunks", type=int, default=0, help="wavefront chunks per call (0 = library default)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the compact runs of the other BASELINE configs (N = 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
