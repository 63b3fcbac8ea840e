"""Drop-in for the reference's vlnce_baselines/common/rgb_mapping.py.

Same names, constructor argument, call signature, tensor layouts and externally
mutated state as the reference module, so `MGMapNet`
(mg_map_policy.py:66,186), `BasePolicy.update_map` (policy.py:30-32) and the trainers'
state juggling (common_trainer.py:142-187,266-267,428-435,454-476; dagger_trainer.py:322-327,
668-678) keep working unchanged:

    from wsmgmap_b200.rgb_mapping import RGBMapping, get_grid

The per-step work is done by libwsmg.so (hand-written sm_100a kernels, C ABI in
include/wsmg.h) on the caller's current CUDA stream.  There is no CPU or eager fallback:
without a CUDA device or the built library the module raises.
"""
from __future__ import annotations

import ctypes
import math
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops


class to_grid():
    """GPS (metres) -> global-map cell, rgb_mapping.py:93-103 of the reference."""

    def __init__(self, global_map_size, coordinate_min, coordinate_max):
        self.global_map_size = global_map_size
        self.coordinate_min = coordinate_min
        self.coordinate_max = coordinate_max
        self.grid_size = (coordinate_max - coordinate_min) / global_map_size

    def get_grid_coords(self, positions):
        col = positions[:, 0]
        row = positions[:, 1]
        return ((self.coordinate_max - col) / self.grid_size).round(), ((row - self.coordinate_min) / self.grid_size).round()


def get_grid(pose, grid_size, device):
    """Rotation and translation sampling grids for `pose` = (x, y, theta) per row.

    Kept for habitat_extensions/sensors.py:22,383-410 (CPU use inside env workers);
    same contract as rgb_mapping.py:106-139: returns (rot_grid, trans_grid), each
    [bs, grid_h, grid_w, 2], built with F.affine_grid on `device`.
    """
    pose = pose.float()
    tx, ty, th = pose[:, 0], pose[:, 1], pose[:, 2]
    n = pose.shape[0]
    c, s = th.cos(), th.sin()
    rot = torch.zeros(n, 2, 3, device=device)
    rot[:, 0, 0] = c
    rot[:, 0, 1] = -s
    rot[:, 1, 0] = s
    rot[:, 1, 1] = c
    trans = torch.zeros(n, 2, 3, device=device)
    trans[:, 0, 0] = 1.0
    trans[:, 1, 1] = 1.0
    trans[:, 0, 2] = tx
    trans[:, 1, 2] = ty
    size = torch.Size(grid_size)
    return F.affine_grid(rot, size, align_corners=False), F.affine_grid(trans, size, align_corners=False)


class Mapping(nn.Module):
    """State + geometry of the reference's Mapping (rgb_mapping.py:11-30).  Holds no
    parameters and no registered buffers (the reference's state is plain attributes, so it
    is absent from state_dict and DDP never touches it)."""

    def __init__(self, model_config):
        super().__init__()
        self.device = torch.device("cuda", model_config.gpu_id)
        self.num_proc = model_config.num_proc
        self.resolution = model_config.resolution
        self.egocentric_map_size = model_config.egocentric_map_size
        self.global_map_size = model_config.global_map_size
        self.global_map_depth = model_config.map_depth
        coordinate_min = -self.global_map_size * self.resolution / 2
        coordinate_max = self.global_map_size * self.resolution / 2
        self.to_grid = to_grid(self.global_map_size, coordinate_min, coordinate_max)
        if not torch.cuda.is_available():
            raise RuntimeError("wsmgmap_b200.Mapping needs a CUDA device (no CPU fallback)")
        self._lib = _lib.load()
        g, c = self.global_map_size, self.global_map_depth
        self._env_slots = None
        self._env_slots_host = None
        # caller-owned, re-bindable state: fetched from the attribute on every call
        self.full_global_map = torch.zeros(self.num_proc, g, g, c, device=self.device)
        self.agent_view = torch.zeros(self.num_proc, c, g, g, device=self.device)
        self._scratch = None
        # Two status words in pinned (device-mapped) host memory that the kernels raise and forward() polls at the
        # NEXT call without synchronising: [0] a valid pixel fell outside the fan and was dropped (depth < 0; the
        # reference would scatter it), [1] an env slot was not a row of the map (include/wsmg.h wsmg_opts.status).
        self._status = torch.zeros(2, dtype=torch.int32).pin_memory()
        self._status_np = self._status.numpy()
        # Opt-in extras beyond the reference's contract; all default to the reference behaviour.
        self.store_half = False      # also emit an fp16 copy of the ego map into self.last_ego_half
        self.last_ego_half = None
        self.strict_inputs = False   # debugging aid: synchronise after each update and raise if a valid pixel was dropped

    # -- caller-owned state with consistency checks -------------------------------------------
    # `full_global_map` is re-bound by the trainers (zeros of a new leading size before a rollout,
    # map[state_index] when envs are paused: common_trainer.py:171-172,266,429-435).  A slot table set through
    # pause_envs()/env_slots refers to rows of the tensor it was made for: re-binding a tensor with another leading
    # dimension drops it (the re-indexed tensor is already compact), so a stale table can never address rows that
    # no longer exist.
    @property
    def full_global_map(self):
        return self._full_global_map

    @full_global_map.setter
    def full_global_map(self, value):
        old = getattr(self, "_full_global_map", None)
        self._full_global_map = value
        if self._env_slots is not None and (old is None or value.shape[0] != old.shape[0]):
            warnings.warn("full_global_map was re-bound with a different number of rows: the env slot table set by "
                          "pause_envs() is dropped (use either pause_envs() or the reference's re-indexing, not both)")
            self._env_slots = self._env_slots_host = None

    @property
    def env_slots(self):
        """int32 [bs] CUDA tensor: frame b reads / updates map row env_slots[b] (None: row b).  See pause_envs."""
        return self._env_slots

    @env_slots.setter
    def env_slots(self, value):
        if value is None:
            self._env_slots = self._env_slots_host = None
            return
        n = self._full_global_map.shape[0]
        host = [int(v) for v in (value.tolist() if isinstance(value, torch.Tensor) else value)]
        if any(v < 0 or v >= n for v in host):
            raise ValueError(f"env_slots {host} must be rows of full_global_map (0..{n - 1})")
        if len(set(host)) != len(host):
            raise ValueError(f"env_slots {host} must be distinct (two frames on one map row would race)")
        self._env_slots_host = host
        self._env_slots = torch.tensor(host, dtype=torch.int32, device=self._full_global_map.device)

    # -- helpers ---------------------------------------------------------------------------
    def _dims(self, bs, n_maps, hf, wf, hd, wd, c_in, feat_nhwc=False):
        c = self.global_map_depth
        return _lib.make_dims(bs, n_maps, c, hf, wf, hd, wd, self.egocentric_map_size, self.global_map_size,
                              self.resolution, c_in=0 if c_in == c else c_in, feat_nhwc=1 if feat_nhwc else 0)

    def _scratch_for(self, dims, device):
        need = self._lib.wsmg_scratch_bytes(ctypes.byref(dims))
        if need == 0:
            raise _lib.WsmgError("unsupported geometry for libwsmg (see include/wsmg.h WSMG_E_*)")
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != device:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=device)
        return self._scratch

    @staticmethod
    def _f32c(t, name, device):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a tensor")
        if t.device != device:
            raise ValueError(f"{name} is on {t.device}, the map lives on {device}")
        if t.dtype != torch.float32:
            t = t.float()
        return t.contiguous()

    # -- reference API ---------------------------------------------------------------------
    def project_feat_to_map(self, features, full_global_map, observations, masks, _trig=None):
        """rgb_mapping.py:32-72.  Updates `full_global_map[:bs]` in place and returns
        (final_retrieval [bs,C,E,E], full_global_map)."""
        if not full_global_map.is_cuda:
            raise ValueError("full_global_map must be a CUDA tensor")
        dev = full_global_map.device
        if full_global_map.dtype != torch.float32 or not full_global_map.is_contiguous():
            raise ValueError("full_global_map must be a contiguous fp32 [n,G,G,C] tensor")
        g, c, e = self.global_map_size, self.global_map_depth, self.egocentric_map_size
        if tuple(full_global_map.shape[1:]) != (g, g, c):
            raise ValueError(f"full_global_map has shape {tuple(full_global_map.shape)}, expected [n,{g},{g},{c}]")
        # A channels_last feature tensor (what cuDNN's NHWC convolutions hand over when the UNet runs in that memory
        # format, unet_encoder.py:103-111) is consumed as it is: no permute copy.  Anything else is made NCHW-contiguous.
        nhwc = (isinstance(features, torch.Tensor) and features.dtype == torch.float32 and ops.is_channels_last(features)
                and features.shape[1] == c and c % 4 == 0)
        if nhwc:
            if features.device != dev:
                raise ValueError(f"features is on {features.device}, the map lives on {dev}")
        else:
            features = self._f32c(features, "features", dev)
        bs, cf, hf, wf = features.shape          # cf != c: the channel pool of rgb_mapping.py:81-84 runs inside the kernel
        if bs > full_global_map.shape[0]:
            raise ValueError(f"batch {bs} larger than the map state ({full_global_map.shape[0]} envs)")
        depth = self._f32c(observations["depth"], "observations['depth']", dev)
        if depth.dim() != 4 or depth.shape[0] != bs or depth.shape[3] != 1:
            raise ValueError("observations['depth'] must be [bs,Hd,Wd,1]")
        gps = self._f32c(observations["gps"], "observations['gps']", dev)
        compass = self._f32c(observations["compass"], "observations['compass']", dev)
        masks = self._f32c(masks, "masks", dev)
        if gps.shape != (bs, 2) or compass.numel() != bs or masks.numel() != bs:
            raise ValueError("gps must be [bs,2], compass [bs,1], masks [bs,1]")
        dims = self._dims(bs, full_global_map.shape[0], hf, wf, depth.shape[1], depth.shape[2], cf, nhwc)
        scratch = self._scratch_for(dims, dev)
        slots = self._env_slots
        if slots is not None:
            if slots.numel() != bs:
                raise ValueError(f"env_slots has {slots.numel()} entries for a batch of {bs}")
            if slots.device != dev or max(self._env_slots_host) >= full_global_map.shape[0]:
                raise ValueError("env_slots does not belong to this map tensor (device or row count differ)")
        self._poll_status()
        half = torch.empty(bs, c, e, e, device=dev, dtype=torch.float16) if self.store_half else None
        ego = ops.map_update(features, depth, gps, compass, masks, full_global_map, e=e, resolution=self.resolution,
                             trig=_trig, scratch=scratch, ego_half=half, env_slots=slots, status=self._status)
        self.last_ego_half = half
        if self.strict_inputs:
            flags = ops.env_flags(scratch, dims)
            bad = (flags & _lib.FLAG_OUTSIDE_FAN).nonzero().flatten().tolist()
            if bad:
                raise ValueError(f"envs {bad}: depth < 0 produced cells behind the camera; the B200 kernel drops them "
                                 "(the reference would scatter them) -- see include/wsmg.h WSMG_FLAG_OUTSIDE_FAN")
        return ego, full_global_map

    def _poll_status(self):
        """What earlier updates reported through the pinned status words (no synchronisation: the kernels may still
        be running, a report can arrive one call late)."""
        st = self._status_np
        if st[0] != 0:
            st[0] = 0
            warnings.warn("wsmgmap_b200: an earlier map update saw depth < 0 -- pixels behind the camera are DROPPED by "
                          "the B200 kernel where the reference would scatter them (include/wsmg.h WSMG_FLAG_OUTSIDE_FAN); "
                          "set strict_inputs=True to locate the frame", RuntimeWarning)
        if st[1] != 0:
            st[1] = 0
            raise _lib.WsmgError("an earlier map update was given an env slot that is not a row of full_global_map: "
                                 "that frame was skipped (include/wsmg.h WSMG_FLAG_BAD_SLOT)")

    # -- opt-in extras ---------------------------------------------------------------------
    def pause_envs(self, envs_to_pause):
        """O(1) replacement for the trainers' `full_global_map = full_global_map[state_index]`
        (common_trainer.py:171-172,454-476): instead of re-materialising the map tensor, drop the paused
        envs from the slot table; the map tensor keeps its rows.  Batch element b then updates row
        env_slots[b].  Callers that use this must NOT also re-index full_global_map."""
        n = self.full_global_map.shape[0]
        slots = list(range(n)) if self._env_slots_host is None else list(self._env_slots_host)
        for idx in sorted(envs_to_pause, reverse=True):
            slots.pop(idx)
        self.env_slots = slots
        return self.env_slots

    def reset_slots(self):
        self.env_slots = None

    def ego_half_to_host(self, pinned_out=None, non_blocking=True):
        """Asynchronous D2H of the fp16 ego map written by the last update (store_half=True): half the bytes of
        the forward hook's `o.cpu()` (dagger_trainer.py:303-306) and no CPU-side astype (common_trainer.py:519-520)."""
        if self.last_ego_half is None:
            raise RuntimeError("store_half is off or no update has run yet")
        if pinned_out is None:
            pinned_out = torch.empty(self.last_ego_half.shape, dtype=torch.float16, pin_memory=True)
        pinned_out.copy_(self.last_ego_half, non_blocking=non_blocking)
        return pinned_out


class RGBMapping(Mapping):
    def __init__(self, model_config):
        super().__init__(model_config)

    def forward(self, rgb_features, observations, masks):
        """rgb_mapping.py:79-90: returns the cached ego map when the observations already carry
        one (LMDB replay, unet_encoder.py:65-66); otherwise one map update."""
        if 'rgb_ego_map' not in observations:
            # The channel re-binning of rgb_mapping.py:82-84 (adaptive_max_pool1d over channels; identity when
            # C_in == map_depth, the shipped config) is fused into the kernel's scatter: features go in as they are.
            final_retrieval, self.full_global_map = self.project_feat_to_map(
                rgb_features, self.full_global_map, observations, masks)
            observations['rgb_ego_map'] = final_retrieval
        else:
            final_retrieval = observations['rgb_ego_map']
        return final_retrieval
