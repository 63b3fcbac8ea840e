"""Env sharding across ranks.  The map update has no cross-env dependency
(the batch dimension is only ever a leading index in rgb_mapping.py), so each rank owns a
contiguous block of envs and their map state for the whole trajectory -- the way each rank of the
reference owns its NUM_PROCESSES envs (config/default.py:187-189).  The only collective is a
stats gather at the end of a batch."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total_envs: int, world: int, rank: int) -> range:
    """Contiguous block of envs for `rank`; the first `total % world` ranks take one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(total_envs, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def gather_stats(stats: torch.Tensor) -> torch.Tensor:
    """all_gather a small 1-D stats vector -> [world, len] (on CPU).  Works with NCCL (device
    tensor) and gloo (CPU tensor); without an initialised process group returns [1, len]."""
    if not (dist.is_available() and dist.is_initialized()):
        return stats.detach().cpu().unsqueeze(0)
    out = [torch.zeros_like(stats) for _ in range(dist.get_world_size())]
    dist.all_gather(out, stats)
    return torch.stack(out).cpu()


def job_throughput(frames_per_rank: torch.Tensor, ms_per_rank: torch.Tensor) -> float:
    """Whole-job frames/s: all frames over the slowest rank's device time."""
    return float(frames_per_rank.sum()) / (float(ms_per_rank.max()) / 1e3)
