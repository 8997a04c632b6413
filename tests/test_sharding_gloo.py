"""N > 1 host-side logic on CPU: two gloo ranks shard the channel range the way bench.py does, run the chain on
their own slice (the CPU oracle stands in for the GPU library here — the CUDA path itself is covered by the
-m gpu tests), and the gathered result is byte-identical to one process running every channel.  Also checks the
max-over-ranks timing reduction."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

WORLD = 2
CH_PER_RANK = 6
NB = 10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    import bench
    import pyoracle as po
    ch0, cnt = bench.rank_channel_range(rank, CH_PER_RANK)
    iq = bench.make_inputs("cfg5", ch0, cnt, NB, unique=cnt)
    params = [po.default_params(**bench.channel_params("cfg5", ch0 + c)) for c in range(cnt)]
    out, _ = po.process_bank(po.default_config(stage_mask=bench.WORKLOADS["cfg5"][1]), params, iq)
    mine = torch.from_numpy(out.astype(np.int32))                    # gloo has no int16 collectives
    gathered = [torch.zeros_like(mine) for _ in range(WORLD)]
    dist.all_gather(gathered, mine)
    slowest = bench.max_over_ranks(10.0 + rank, dist, torch.device("cpu"))
    if rank == 0:
        np.save(out_path, np.concatenate([g.numpy().astype(np.int16) for g in gathered], axis=1))
        assert slowest == 10.0 + WORLD - 1
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_channel_sharding(tmp_path, po):
    import bench
    out_path = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(_free_port(), out_path), nprocs=WORLD, join=True)
    got = np.load(out_path)
    total = WORLD * CH_PER_RANK
    iq = np.concatenate([bench.make_inputs("cfg5", r * CH_PER_RANK, CH_PER_RANK, NB, unique=CH_PER_RANK) for r in range(WORLD)], axis=1)
    params = [po.default_params(**bench.channel_params("cfg5", c)) for c in range(total)]
    want, _ = po.process_bank(po.default_config(stage_mask=bench.WORKLOADS["cfg5"][1]), params, iq)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert [bench.rank_channel_range(r, 8192) for r in range(8)] == [(r * 8192, 8192) for r in range(8)]


def test_workload_tables_are_consistent():
    import bench
    for wl, (_, stage, ch) in bench.WORKLOADS.items():
        ab = bench.algorithmic_bytes(wl, 1)
        assert ab["_step"] == {"cfg2": 2560, "cfg3": 5136, "cfg4a": 6192, "cfg4b": 3624, "cfg5": 14169}[wl]
        assert bench.algorithmic_bytes(wl, 8)["_step"] == 1024 + (ab["_step"] - 1024) / 8
        for c in range(20):
            p = bench.channel_params(wl, c)
            assert 0 <= p["demod"] <= 4
    # config 5: mode by c mod 4, DNR level by c mod 5
    assert [bench.channel_params("cfg5", c)["demod"] for c in range(4)] == [0, 1, 2, 4]
    assert [bench.channel_params("cfg5", c)["nr_level"] for c in range(5)] == [0, 20, 30, 40, 50]


def test_roofline_traffic_lookup_reads_the_committed_capture():
    """bench.py's `roofline.traffic` comes from the newest profiles/r*_traffic.json (one `ncu --set full` capture of the
    un-profiled step, where a stage may run as two launches): every kernel of the cfg5 step resolves to a byte count,
    and nothing is claimed for configurations the capture was not taken on"""
    import bench
    for k in ("k_front", "k_nlms_notch", "k_nlms_dnr", "k_agc", "k_fftfilt", "k_biquad", "k_spec256", "k_spec1024"):
        t = bench.ncu_traffic(k, "cfg5", 8192, 8)
        assert t is not None and 1e6 < t < 1e9, (k, t)
    # the notch covers a quarter of the channels of cfg5 and reads q15: far less traffic than the DNR launches
    assert bench.ncu_traffic("k_nlms_notch", "cfg5", 8192, 8) < bench.ncu_traffic("k_nlms_dnr", "cfg5", 8192, 8)
    assert bench.ncu_traffic("k_front", "cfg3", 8192, 8) is None and bench.ncu_traffic("k_front", "cfg5", 4096, 8) is None
