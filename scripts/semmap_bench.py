"""Batched ground-truth semantic map sensor: CUDA kernel vs the reference's torch ops (oracle port) on the GPU and
on the host cores.  Usage on the GPU box: python scripts/semmap_bench.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import wsmgmap_b200  # noqa: F401
from wsmgmap_b200 import ops
from oracle.semmap_oracle import sensor_crop, sensor_pose

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
bs, s = 1024, 480
maps = torch.from_numpy(rng.integers(0, 28, (bs, s, s)).astype(np.float32))
pose = sensor_pose(rng.uniform(150, 330, bs), rng.uniform(150, 330, bs), rng.uniform(-np.pi, np.pi, bs))
maps_d, pose_d = maps.to(dev), pose.to(dev)
for _ in range(3):
    out = ops.semantic_crop(maps_d, pose_d)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = ops.semantic_crop(maps_d, pose_d)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"k_semcrop: {bs} envs in {ms:.3f} ms -> {bs / ms * 1e3:.0f} observations/s")
n = 64
sensor_crop(maps_d[:n], pose_d[:n]); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    ref = sensor_crop(maps_d[:n], pose_d[:n])
torch.cuda.synchronize()
print(f"reference torch ops on the same GPU: {n * 5 / (time.perf_counter() - t0):.0f} observations/s")
assert (ref.cpu() != out[:n].cpu()).float().mean() < 1e-3
n = 8
t0 = time.perf_counter()
for _ in range(3):
    sensor_crop(maps[:n], pose[:n])
print(f"reference torch ops on {os.cpu_count()} host cores: {n * 3 / (time.perf_counter() - t0):.0f} observations/s "
      f"(the reference runs this per env inside its worker processes)")
