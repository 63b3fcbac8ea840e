"""N>1 host logic on CPU: two gloo ranks each update their env shard (through the kernels' host
emulation), gather stats, and the concatenation equals the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import wsmgmap_b200  # noqa: F401
from wsmgmap_b200.shard import gather_stats, job_throughput, shard_range


def _frames(total, c, hf, hd, seed=0):
    rs = np.random.default_rng(seed)
    return (rs.random((total, c, hf, hf), np.float32), (rs.random((total, hd, hd), np.float32) * 0.6).astype(np.float32),
            rs.normal(size=(total, 2)).astype(np.float32), rs.uniform(-3, 3, (total, 1)).astype(np.float32))


def _worker(rank, world, port, total, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    from emul import emul_step
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = shard_range(total, world, rank)
    feat, depth, gps, compass = _frames(total, 4, 24, 32)
    sl = slice(rng.start, rng.stop)
    gmap = np.zeros((len(rng), 240, 240, 4), np.float32)
    ego, _ = emul_step(gmap, feat[sl], depth[sl], gps[sl], compass[sl], np.zeros((len(rng), 1), np.float32))
    stats = torch.tensor([float(len(rng)), 10.0 + rank, float(ego.astype(np.float64).sum()), float(gmap.astype(np.float64).sum())],
                         dtype=torch.float64)
    allst = gather_stats(stats)
    np.save(os.path.join(out_dir, f"ego{rank}.npy"), ego)
    if rank == 0:
        np.save(os.path.join(out_dir, "stats.npy"), allst.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for total, world in [(1024, 8), (1024, 1), (10, 4), (3, 8)]:
        seen = []
        for r in range(world):
            seen += list(shard_range(total, world, r))
        assert seen == list(range(total))
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_two_rank_gloo_equals_single_process(tmp_path):
    from emul import emul_step
    total, world = 5, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    feat, depth, gps, compass = _frames(total, 4, 24, 32)
    gmap = np.zeros((total, 240, 240, 4), np.float32)
    ego, _ = emul_step(gmap, feat, depth, gps, compass, np.zeros((total, 1), np.float32))
    got = np.concatenate([np.load(tmp_path / f"ego{r}.npy") for r in range(world)], 0)
    assert np.array_equal(got, ego)
    st = torch.from_numpy(np.load(tmp_path / "stats.npy"))
    assert st.shape == (2, 4) and st[:, 0].tolist() == [3.0, 2.0]
    assert abs(float(st[:, 2].sum()) - float(ego.astype(np.float64).sum())) < 1e-6 * abs(float(ego.sum()))
    assert job_throughput(st[:, 0], st[:, 1]) == pytest.approx(5 / 0.011)
