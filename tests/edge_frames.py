"""Edge-case frames shared by the CPU (emulation) and GPU parity tests."""
import numpy as np
import torch


def edge_depths(hd, seed=0):
    """name -> depth [1,hd,hd,1] in the sensor's units (metres / 10)."""
    g = torch.Generator().manual_seed(seed)
    rnd = torch.rand(1, hd, hd, 1, generator=g)
    out = {
        "empty_all_zero": torch.zeros(1, hd, hd, 1),                       # no pixel writes: ego map must be zeros, map untouched
        "all_far": torch.ones(1, hd, hd, 1),                               # 10 m: every cell index out of the ego grid
        "wall_constant": torch.full((1, hd, hd, 1), 0.2),                  # whole columns collapse into single cells (max collisions)
        "one_cell": torch.full((1, hd, hd, 1), 0.001),                     # 1 cm: every valid pixel lands in the apex cell
        "half_cell_multiples": (rnd * 100).round() * 0.006,                # depth*10/0.12 exactly on rounding boundaries
    }
    bad = rnd.clone() * 0.5
    bad[0, ::3, ::5, 0] = float("nan")
    bad[0, 1::4, 2::7, 0] = float("inf")
    bad[0, 2::5, 1::3, 0] = 0.0
    out["nan_inf_holes"] = bad
    return out
