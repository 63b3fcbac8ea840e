"""Thin torch-tensor wrappers over the C ABI (include/wsmg.h).  Each function checks that its tensors are what the
raw pointers will be read as (device, dtype, contiguity) -- libwsmg validates dims/pointers and its error code is
raised as WsmgError.  All work is enqueued on the current CUDA stream."""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _chk(t, name, device=None, dtype=torch.float32, numel=None, optional=False, channels_last_ok=False):
    """The C ABI reads raw pointers: refuse anything it would misread."""
    if t is None:
        if optional:
            return
        raise ValueError(f"{name} is required")
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a tensor")
    if device is not None and t.device != device:
        raise ValueError(f"{name} is on {t.device}, expected {device}")
    if t.dtype != dtype:
        raise ValueError(f"{name} has dtype {t.dtype}, expected {dtype}")
    if not (t.is_contiguous() or (channels_last_ok and is_channels_last(t))):
        raise ValueError(f"{name} must be contiguous")
    if numel is not None and t.numel() != numel:
        raise ValueError(f"{name} has {t.numel()} elements, expected {numel}")


def dims_for(feat_shape, depth_shape, n_maps, e=100, g=240, resolution=0.12, map_depth=None, feat_nhwc=False):
    """map_depth: channels of the map when they differ from the features' (channel pool fused in the kernel);
    feat_nhwc: the feature tensor is channels_last in memory ([bs,Hf,Wf,C]); feat_shape stays the logical NCHW shape."""
    bs, c_in, hf, wf = feat_shape
    c = c_in if map_depth is None else map_depth
    return _lib.make_dims(bs, n_maps, c, hf, wf, depth_shape[1], depth_shape[2], e, g, resolution, c_in=0 if c == c_in else c_in,
                          feat_nhwc=1 if feat_nhwc else 0)


def is_channels_last(t) -> bool:
    """A 4-D tensor whose memory is [N,H,W,C] (torch.channels_last) and not also plain contiguous."""
    return t.dim() == 4 and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last)


def scratch_bytes(dims) -> int:
    return int(_lib.load().wsmg_scratch_bytes(ctypes.byref(dims)))


def alloc_scratch(dims, device):
    n = scratch_bytes(dims)
    if n == 0:
        raise _lib.WsmgError("unsupported geometry for libwsmg (see include/wsmg.h WSMG_E_*)")
    return torch.empty(n, dtype=torch.uint8, device=device)


def map_update(feat, depth, gps, compass, mask, gmap, e=100, resolution=0.12, trig=None, scratch=None, ego=None,
               ego_half=None, env_slots=None, status=None, events=None):
    """One step.  feat [bs,C,Hf,Wf], depth [bs,Hd,Wd,1], gmap [n,G,G,C] (updated in place).  Returns ego [bs,C,E,E].
    ego_half: optional fp16 [bs,C,E,E] tensor that receives the rounded copy (rollout store);
    env_slots: optional int32 [bs] map row per frame; status: optional pinned int32[2] the kernels raise
    (dropped pixels / bad slot); events: optional (before, after) torch.cuda.Event pair recorded around k_fused
    (see include/wsmg.h wsmg_opts)."""
    lib = _lib.load()
    if not gmap.is_cuda:
        raise ValueError("gmap must be a CUDA tensor (no CPU fallback)")
    dev = gmap.device
    bs = feat.shape[0]
    _chk(gmap, "gmap", dev)
    _chk(feat, "feat", dev, channels_last_ok=True)          # a channels_last producer is consumed as it is (wsmg_dims.feat_nhwc)
    _chk(depth, "depth", dev, numel=bs * depth.shape[1] * depth.shape[2])
    _chk(gps, "gps", dev, numel=2 * bs)
    _chk(compass, "compass", dev, numel=bs)
    _chk(mask, "mask", dev, numel=bs)
    _chk(trig, "trig", dev, numel=4 * bs, optional=True)
    d = dims_for(feat.shape, depth.shape, gmap.shape[0], e, gmap.shape[1], resolution, map_depth=gmap.shape[3],
                 feat_nhwc=is_channels_last(feat))
    if scratch is None:
        scratch = alloc_scratch(d, dev)
    _chk(scratch, "scratch", dev, dtype=torch.uint8)
    if ego is None:
        ego = torch.empty(bs, gmap.shape[3], e, e, device=dev, dtype=torch.float32)
    _chk(ego, "ego", dev, numel=bs * gmap.shape[3] * e * e)
    if ego_half is not None:
        _chk(ego_half, "ego_half", dev, dtype=torch.float16, numel=ego.numel())
    _chk(env_slots, "env_slots", dev, dtype=torch.int32, numel=bs, optional=True)
    if status is not None and not (status.dtype == torch.int32 and status.numel() >= 2 and status.is_contiguous()
                                   and (status.is_pinned() or status.device == dev)):
        raise ValueError("status must be an int32[2] tensor in pinned host memory (or on the device)")
    opts = _lib.WsmgOpts(_ptr(trig), _ptr(ego_half), _ptr(env_slots),
                         None if events is None else ctypes.c_void_p(events[0].cuda_event),
                         None if events is None else ctypes.c_void_p(events[1].cuda_event), _ptr(status))
    with torch.cuda.device(dev):
        rc = lib.wsmg_map_update_ex(_ptr(feat), _ptr(depth), _ptr(gps), _ptr(compass), _ptr(mask), _ptr(gmap), _ptr(ego),
                                    ctypes.byref(opts), _ptr(scratch), scratch.numel(), ctypes.byref(d), _stream(dev))
    _lib.check(rc, "wsmg_map_update_ex")
    return ego


def env_flags(scratch, dims):
    """Per-env status words of the last update that used `scratch` (synchronises): int32 [bs], bits
    _lib.FLAG_INVALID_PIXEL / _lib.FLAG_OUTSIDE_FAN / _lib.FLAG_BAD_SLOT (see include/wsmg.h)."""
    off = int(_lib.load().wsmg_scratch_flags_offset(ctypes.byref(dims)))
    return scratch[off:off + 4 * dims.bs].view(torch.int32).cpu()


def unproject_index(depth, hf, wf, e=100, g=240, resolution=0.12):
    """depth [bs,Hd,Wd,1] -> (lin int32 [bs,hf,wf], invalid bool [bs,hf,wf])."""
    lib = _lib.load()
    if not depth.is_cuda:
        raise ValueError("depth must be a CUDA tensor (no CPU fallback)")
    dev = depth.device
    bs = depth.shape[0]
    _chk(depth, "depth", dev)
    d = _lib.make_dims(bs, bs, 4, hf, wf, depth.shape[1], depth.shape[2], e, g, resolution)
    lin = torch.empty(bs, hf, wf, dtype=torch.int32, device=dev)
    inv = torch.empty(bs, hf, wf, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.wsmg_unproject_index(_ptr(depth), _ptr(lin), _ptr(inv), ctypes.byref(d), _stream(dev))
    _lib.check(rc, "wsmg_unproject_index")
    return lin, inv.bool()


def scatter_max(feat, depth, e=100, g=240, resolution=0.12, map_depth=None):
    """-> proj_feats [bs,C,E,E] (before rotation); map_depth != feat channels applies the fused channel pool."""
    lib = _lib.load()
    if not feat.is_cuda:
        raise ValueError("feat must be a CUDA tensor (no CPU fallback)")
    dev = feat.device
    _chk(feat, "feat", dev)
    _chk(depth, "depth", dev, numel=feat.shape[0] * depth.shape[1] * depth.shape[2])
    d = dims_for(feat.shape, depth.shape, feat.shape[0], e, g, resolution, map_depth=map_depth)
    scratch = alloc_scratch(d, dev)
    proj = torch.empty(feat.shape[0], d.C, e, e, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        rc = lib.wsmg_scatter_max(_ptr(feat), _ptr(depth), _ptr(proj), _ptr(scratch), scratch.numel(),
                                  ctypes.byref(d), _stream(dev))
    _lib.check(rc, "wsmg_scatter_max")
    return proj


def register_fuse_retrieve(proj, gps, compass, mask, gmap, resolution=0.12, trig=None):
    """Everything after the projection.  proj [bs,C,E,E] must be zero outside the fan a depth >= 0 pixel can reach
    (what scatter_max returns): cells outside it are ignored."""
    lib = _lib.load()
    if not gmap.is_cuda:
        raise ValueError("gmap must be a CUDA tensor (no CPU fallback)")
    dev = gmap.device
    bs, c, e, _ = proj.shape
    for t, name, n in ((proj, "proj", None), (gmap, "gmap", None), (gps, "gps", 2 * bs), (compass, "compass", bs), (mask, "mask", bs)):
        _chk(t, name, dev, numel=n)
    _chk(trig, "trig", dev, numel=4 * bs, optional=True)
    d = _lib.make_dims(bs, gmap.shape[0], c, 4, 4, 4, 4, e, gmap.shape[1], resolution)
    scratch = alloc_scratch(d, dev)
    ego = torch.empty(bs, c, e, e, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        rc = lib.wsmg_register_fuse_retrieve(_ptr(proj), _ptr(gps), _ptr(compass), _ptr(mask), _ptr(gmap), _ptr(ego),
                                             _ptr(trig), _ptr(scratch), scratch.numel(), ctypes.byref(d), _stream(dev))
    _lib.check(rc, "wsmg_register_fuse_retrieve")
    return ego


def semantic_crop(maps, pose, half=50, origin=289, trig=None, map_index=None):
    """Batched ground-truth semantic map sensor (habitat_extensions/sensors.py:403-410, wsmg_semantic_crop).
    maps [n,S,S] fp32 CUDA class ids, pose [bs,3] fp32 CUDA = ((grid_y-S/2)/(S/2), (grid_x-S/2)/(S/2), -heading),
    trig optional [bs,2] (cos, sin of pose[:,2]), map_index optional int32 [bs].  Returns int64 [bs,2*half,2*half]."""
    lib = _lib.load()
    if not (maps.is_cuda and pose.is_cuda and maps.dtype == torch.float32 and pose.dtype == torch.float32):
        raise ValueError("semantic_crop: fp32 CUDA tensors required")
    if maps.dim() != 3 or maps.shape[1] != maps.shape[2] or pose.dim() != 2 or pose.shape[1] != 3:
        raise ValueError("semantic_crop: maps [n,S,S], pose [bs,3]")
    maps, pose = maps.contiguous(), pose.contiguous()
    bs, dev = pose.shape[0], maps.device
    out = torch.empty(bs, 2 * half, 2 * half, dtype=torch.int64, device=dev)
    trig = None if trig is None else trig.to(dev, torch.float32).contiguous()
    map_index = None if map_index is None else map_index.to(dev, torch.int32).contiguous()
    with torch.cuda.device(dev):
        rc = lib.wsmg_semantic_crop(_ptr(maps), _ptr(pose), _ptr(trig), _ptr(map_index), _ptr(out), bs, maps.shape[0],
                                    maps.shape[1], half, origin, _stream(dev))
    _lib.check(rc, "wsmg_semantic_crop")
    return out


def host_live_rows(depth_h, dims):
    """First / last feature row per frame that holds a pixel which can write (what skip_dead_rows copies); depth_h is
    the HOST depth tensor [bs,Hd,Wd,1].  Returns two int32 CPU tensors [bs]."""
    lib = _lib.load()
    depth_h = depth_h.contiguous()
    lo = torch.empty(dims.bs, dtype=torch.int32)
    hi = torch.empty(dims.bs, dtype=torch.int32)
    _lib.check(lib.wsmg_host_live_rows(_ptr(depth_h), ctypes.byref(dims), _ptr(lo), _ptr(hi)), "wsmg_host_live_rows")
    return lo, hi


class HostPipeline:
    """End-to-end step from HOST (pinned) buffers: H2D of the frame, update, D2H of the ego map,
    chunked so copies overlap the kernels (wsmg_map_update_host_ex).  zero_copy=True: the feature tensor must be
    pinned (`pin_memory()`); the scatter pulls it over the bus itself and skips the pixel groups that cannot write."""

    def __init__(self, dims, device, chunk_envs=4, zero_copy=False, skip_dead_rows=False):
        # skip_dead_rows: the host tests the depth frame first and copies only the feature rows that hold a pixel
        # which can write (indoor frames: about half)
        self.flags = (_lib.HOST_ZEROCOPY_FEATURES if zero_copy else 0) | (_lib.HOST_SKIP_DEAD_ROWS if skip_dead_rows else 0)
        self.lib = _lib.load()
        self.dims = dims
        self.device = torch.device(device)
        self.chunk = int(min(chunk_envs, dims.bs))
        n = int(self.lib.wsmg_host_staging_bytes(ctypes.byref(dims), self.chunk))
        if n == 0:
            raise _lib.WsmgError("unsupported geometry for libwsmg")
        self.staging = torch.empty(n, dtype=torch.uint8, device=self.device)

    def step(self, feat_h, depth_h, gps_h, compass_h, mask_h, gmap, ego_h):
        bs = self.dims.bs
        for t, name, n in ((feat_h, "feat_h", None), (depth_h, "depth_h", bs * self.dims.Hd * self.dims.Wd), (gps_h, "gps_h", 2 * bs),
                           (compass_h, "compass_h", bs), (mask_h, "mask_h", bs), (ego_h, "ego_h", None)):
            _chk(t, name, torch.device("cpu"), numel=n)
        _chk(gmap, "gmap", self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.wsmg_map_update_host_ex(_ptr(feat_h), _ptr(depth_h), _ptr(gps_h), _ptr(compass_h), _ptr(mask_h),
                                                  _ptr(gmap), _ptr(ego_h), _ptr(self.staging), self.staging.numel(),
                                                  self.chunk, ctypes.byref(self.dims), self.flags, _stream(self.device))
        _lib.check(rc, "wsmg_map_update_host_ex")
