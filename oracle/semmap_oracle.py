"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the ground-truth semantic map sensor's
registration, `GtSemanticMapSensor.get_observation`,
/root/reference/habitat_extensions/sensors.py:383-410:

    st_pose  = [(grid_y - 240)/240, (grid_x - 240)/240, -heading]               (:397-401)
    rot, tra = get_grid(st_pose, map.size(), 'cpu')                              (:403, rgb_mapping.py:106-139)
    transed  = F.grid_sample(map, tra, mode='nearest')                           (:404)
    rotated  = F.grid_sample(transed, rot, mode='nearest')                       (:405)
    rotated  = F.pad(rotated, (half,)*4, 'constant', 0)                          (:406)
    return rotated.squeeze()[289-half:289+half, 289-half:289+half].long()        (:410)

Two forms, like oracle/mapping_oracle.py:
  * `sensor_crop` -- the same torch ops, batched over envs (also runs on CUDA as the
    "stock PyTorch on the same box" comparator);
  * `spec_sensor_crop` -- numpy, one output cell at a time, the arithmetic the CUDA kernel
    implements: affine_grid through MKL's K=3 bmm (gx = fma(y, -sin, x*cos), gy = fma(y, cos, x*sin);
    translation x + tx), unnormalize fma(g+1, S/2, -0.5), nearest = rint (half to even), zeros outside.
Pinned against the reference's own method body executed with a fake simulator
(oracle/reference_loader.py: run_reference_sensor) in tests/test_semmap_sensor.py.
Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this file.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .mapping_oracle import _fma, _spec_unnormalize, f32, spec_base_coords

ORIGIN = 289          # sensors.py:410 (hard-coded crop centre in the padded map)


def get_grid(pose: torch.Tensor, grid_size, device):
    """rgb_mapping.py:106-139, restated (same ops, same order)."""
    pose = pose.float()
    x, y, t = pose[:, 0], pose[:, 1], pose[:, 2]
    cos_t, sin_t = t.cos(), t.sin()
    zero = torch.zeros_like(cos_t)
    theta1 = torch.stack([torch.stack([cos_t, -sin_t, zero], 1), torch.stack([sin_t, cos_t, zero], 1)], 1)
    one = torch.ones_like(x)
    theta2 = torch.stack([torch.stack([one, -zero, x], 1), torch.stack([zero, one, y], 1)], 1)
    return (F.affine_grid(theta1, torch.Size(grid_size), align_corners=False),
            F.affine_grid(theta2, torch.Size(grid_size), align_corners=False))


def sensor_pose(grid_y, grid_x, heading, size: int = 480) -> torch.Tensor:
    """sensors.py:397-401 for a batch: float32 [bs,3]."""
    h = size // 2
    return torch.stack([(torch.as_tensor(grid_y, dtype=torch.float32) - h) / h,
                        (torch.as_tensor(grid_x, dtype=torch.float32) - h) / h,
                        -torch.as_tensor(heading, dtype=torch.float32)], 1)


def sensor_crop(maps: torch.Tensor, pose: torch.Tensor, half: int = 50, origin: int = ORIGIN) -> torch.Tensor:
    """maps [bs,S,S] float32 class ids, pose [bs,3] (sensor_pose) -> int64 [bs,2*half,2*half]  (sensors.py:403-410)."""
    m = maps.unsqueeze(1).float()
    rot, tra = get_grid(pose, m.size(), m.device)
    transed = F.grid_sample(m, tra, mode="nearest", align_corners=False)
    rotated = F.grid_sample(transed, rot, mode="nearest", align_corners=False)
    rotated = F.pad(rotated, (half, half, half, half), "constant", 0)
    return rotated[:, 0, origin - half:origin + half, origin - half:origin + half].long()


def _nearest_index(g: np.ndarray, size: int):
    """unnormalize + nearbyint; returns (index, inside)."""
    r = np.rint(_spec_unnormalize(g.astype(f32), size)).astype(f32)
    inside = (r > -1) & (r < size)
    return np.where(inside, r, 0).astype(np.int64), inside


def spec_sensor_crop(maps: np.ndarray, pose: np.ndarray, cos_t: np.ndarray, sin_t: np.ndarray, half: int = 50,
                     origin: int = ORIGIN) -> np.ndarray:
    """Elementwise spec.  maps [bs,S,S] float32, pose [bs,3] float32 (only x, y used), cos_t/sin_t [bs] float32
    (cos / sin of pose[:,2] as the host computed them)."""
    bs, s, _ = maps.shape
    base = spec_base_coords(s)
    out = np.zeros((bs, 2 * half, 2 * half), np.int64)
    r0 = origin - 2 * half                       # first row / column of the un-padded rotated map in the crop
    rr = np.arange(r0, r0 + 2 * half)
    ok_r = (rr >= 0) & (rr < s)
    rc = np.clip(rr, 0, s - 1)
    for b in range(bs):
        cs, sn = f32(cos_t[b]), f32(sin_t[b])
        bx = base[rc][None, :]                   # column coordinate of the sampled rotated cell
        by = base[rc][:, None]
        gx = _fma(by, -sn, (bx * cs).astype(f32))
        gy = _fma(by, cs, (bx * sn).astype(f32))
        x1, in_x1 = _nearest_index(gx, s)
        y1, in_y1 = _nearest_index(gy, s)
        tx, ty = f32(pose[b, 0]), f32(pose[b, 1])
        x2, in_x2 = _nearest_index((base[x1] + tx).astype(f32), s)
        y2, in_y2 = _nearest_index((base[y1] + ty).astype(f32), s)
        inside = in_x1 & in_y1 & in_x2 & in_y2 & ok_r[None, :] & ok_r[:, None]
        out[b] = np.where(inside, maps[b][y2, x2], 0).astype(np.int64)
    return out
