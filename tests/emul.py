"""ctypes access to the host emulation of the kernels (test infrastructure)."""
import ctypes
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import wsmgmap_b200  # noqa: E402,F401
from wsmgmap_b200._lib import WsmgDims, make_dims  # noqa: E402
from wsmgmap_b200.build import build_emulation  # noqa: E402

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_emulation())
        _lib.wsmg_emul_step.restype = ctypes.c_int
        _lib.wsmg_emul_unproject_index.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def emul_cells(depth, hf, e=100, g=240, res=0.12):
    bs, hd, wd = depth.shape
    d = make_dims(bs, bs, 4, hf, hf, hd, wd, e, g, res)
    lin = np.zeros((bs, hf, hf), np.int32)
    inv = np.zeros((bs, hf, hf), np.uint8)
    codes = np.zeros((bs, hf, hf), np.uint16)
    rc = lib().wsmg_emul_unproject_index(_p(np.ascontiguousarray(depth)), _p(lin), _p(inv), _p(codes), None, ctypes.byref(d))
    assert rc == 0, rc
    return lin, inv.astype(bool), codes


def emul_step(gmap, feat, depth, gps, compass, masks, trig=None, mode=0, proj_in=None, want_proj=False, e=100, g=240, res=0.12,
              ego_half=None, env_slots=None, map_depth=None, feat_nhwc=False):
    """gmap [n,G,G,C] updated in place.  trig [bs,4] or None.  Returns (ego, proj or None).
    ego_half: optional uint16 array [bs,C,E,E] receiving the fp16 bits; env_slots: optional int32 [bs]."""
    bs, c_in, hf, wf = feat.shape if feat is not None else (proj_in.shape[0], proj_in.shape[1], 4, 4)
    c = c_in if map_depth is None else map_depth
    hd, wd = (depth.shape[1], depth.shape[2]) if depth is not None else (4, 4)
    d = make_dims(bs, gmap.shape[0] if gmap is not None else bs, c, hf, wf, hd, wd, e, g, res, c_in=0 if c == c_in else c_in,
                  feat_nhwc=1 if feat_nhwc else 0)
    if feat_nhwc:                                   # logical NCHW array in, NHWC memory to the kernel body
        feat = np.ascontiguousarray(np.transpose(feat, (0, 2, 3, 1)))
    ego = np.zeros((bs, c, e, e), np.float32)
    proj = np.zeros((bs, c, e, e), np.float32) if (want_proj or mode == 1) else None
    arrs = [None if a is None else np.ascontiguousarray(a, np.float32) for a in (feat, depth, gps, compass, masks)]
    trig_a = None if trig is None else np.ascontiguousarray(trig, np.float32)
    pin = None if proj_in is None else np.ascontiguousarray(proj_in, np.float32)
    rc = lib().wsmg_emul_step(_p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]), _p(arrs[4]), _p(gmap), _p(ego),
                              _p(trig_a), _p(proj), _p(pin), ctypes.c_int(mode), ctypes.byref(d), _p(ego_half),
                              _p(None if env_slots is None else np.ascontiguousarray(env_slots, np.int32)))
    assert rc == 0, rc
    return ego, proj


def emul_semantic_crop(maps, pose, trig=None, map_index=None, half=50, origin=289):
    maps = np.ascontiguousarray(maps, np.float32)
    pose = np.ascontiguousarray(pose, np.float32)
    bs = pose.shape[0]
    out = np.zeros((bs, 2 * half, 2 * half), np.int64)
    trig_a = None if trig is None else np.ascontiguousarray(trig, np.float32)
    mi = None if map_index is None else np.ascontiguousarray(map_index, np.int32)
    rc = lib().wsmg_emul_semantic_crop(_p(maps), _p(pose), _p(trig_a), _p(mi), _p(out), ctypes.c_int32(bs),
                                       ctypes.c_int32(maps.shape[0]), ctypes.c_int32(maps.shape[1]), ctypes.c_int32(half),
                                       ctypes.c_int32(origin))
    assert rc == 0, rc
    return out
