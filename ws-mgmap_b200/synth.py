"""Synthetic inputs of the map update's shapes.

Depth is what Habitat's depth sensor hands the reference: fp32 [bs,Hd,Wd,1] in
[0,1] (metres / 10, see rgb_mapping.py:37 in the reference).  Features stand in
for the UNet's post-ReLU `proj_feat` [bs,C,Hf,Wf] (unet_encoder.py:103-111).
Poses follow the VLN-CE action space: forward 0.25 m or turn 15 degrees
(habitat_extensions/config/vlnce_task.yaml:6-7).
"""
from __future__ import annotations

import math

import torch

DEPTH_KINDS = ("uniform", "near", "room2", "room4")


def make_depth(kind: str, bs: int, hd: int, wd: int, gen: torch.Generator, device="cpu") -> torch.Tensor:
    dev = torch.device(device)
    if kind == "uniform":
        d = torch.rand(bs, hd, wd, 1, generator=gen, device=dev)
        hole = torch.rand(bs, hd, wd, 1, generator=gen, device=dev) < 0.05
        return d.masked_fill_(hole, 0.0)
    if kind == "near":
        return torch.rand(bs, hd, wd, 1, generator=gen, device=dev) * 0.55 + 0.05
    if kind in ("room2", "room4"):
        wall = 2.0 if kind == "room2" else 4.0
        rows = torch.arange(hd, 0, -1, device=dev, dtype=torch.float32)
        yy = (rows - hd / 2.0) / (hd / 2.0)                      # +1 top ... ~-1 bottom
        floor = torch.where(yy < 0, 1.25 / (-yy).clamp_min(1e-6), torch.full_like(yy, 1e9))
        z = torch.minimum(floor, torch.full_like(floor, wall)).view(1, hd, 1, 1).expand(bs, hd, wd, 1)
        jitter = 1.0 + 0.01 * (torch.rand(bs, hd, wd, 1, generator=gen, device=dev) - 0.5)
        return (z * jitter / 10.0).contiguous()
    raise ValueError(f"unknown depth kind {kind!r}; expected one of {DEPTH_KINDS}")


def make_features(bs: int, c: int, hf: int, wf: int, gen: torch.Generator, device="cpu", signed=False) -> torch.Tensor:
    dev = torch.device(device)
    if signed:
        return torch.randn(bs, c, hf, wf, generator=gen, device=dev)
    return torch.rand(bs, c, hf, wf, generator=gen, device=dev)


class RandomWalk:
    """Per-env random walk: each step either forward 0.25 m along the heading or a
    +-15 degree turn; `masks` is 0 on the first step and on sparse resets."""

    def __init__(self, n_envs: int, seed: int, reset_prob: float = 0.0, far_env: int | None = None):
        self.gen = torch.Generator().manual_seed(seed)
        self.n = n_envs
        self.gps = torch.zeros(n_envs, 2)
        self.compass = (torch.rand(n_envs, 1, generator=self.gen) * 2 - 1) * math.pi
        self.reset_prob = reset_prob
        self.far_env = far_env
        self.t = 0

    def step(self):
        """Returns (gps [n,2], compass [n,1], masks [n,1]) for the next frame (CPU tensors)."""
        n = self.n
        if self.t == 0:
            masks = torch.zeros(n, 1)
        else:
            fwd = torch.rand(n, generator=self.gen) < 0.6
            turn = torch.where(torch.rand(n, generator=self.gen) < 0.5, -1.0, 1.0) * math.radians(15.0)
            self.compass[:, 0] = torch.where(fwd, self.compass[:, 0], self.compass[:, 0] + turn)
            self.compass[:, 0] = torch.remainder(self.compass[:, 0] + math.pi, 2 * math.pi) - math.pi
            stride = torch.where(fwd, 0.25, 0.0)
            if self.far_env is not None:
                stride[self.far_env] = torch.where(fwd[self.far_env], 0.6, 0.0)
            self.gps[:, 0] += stride * torch.cos(self.compass[:, 0])
            self.gps[:, 1] -= stride * torch.sin(self.compass[:, 0])
            reset = torch.rand(n, generator=self.gen) < self.reset_prob
            masks = (~reset).float().view(n, 1)
            self.gps[reset] = 0.0
        self.t += 1
        return self.gps.clone(), self.compass.clone(), masks
