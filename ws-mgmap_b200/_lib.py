"""ctypes binding of libwsmg.so (C ABI in include/wsmg.h).  No fallback: if the library or
a CUDA device is missing, the product path raises."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libwsmg.so")


class WsmgDims(ctypes.Structure):
    _fields_ = [
        ("bs", ctypes.c_int32), ("n_maps", ctypes.c_int32), ("C", ctypes.c_int32),
        ("Hf", ctypes.c_int32), ("Wf", ctypes.c_int32), ("Hd", ctypes.c_int32), ("Wd", ctypes.c_int32),
        ("E", ctypes.c_int32), ("G", ctypes.c_int32), ("resolution", ctypes.c_double), ("C_in", ctypes.c_int32),
        ("feat_nhwc", ctypes.c_int32),
    ]


class WsmgOpts(ctypes.Structure):
    _fields_ = [("trig", ctypes.c_void_p), ("ego_half", ctypes.c_void_p), ("env_slots", ctypes.c_void_p),
                ("ev_before_fused", ctypes.c_void_p), ("ev_after_fused", ctypes.c_void_p), ("status", ctypes.c_void_p)]


_P = ctypes.c_void_p
_DP = ctypes.POINTER(WsmgDims)
_OP = ctypes.POINTER(WsmgOpts)

SIGNATURES = {
    "wsmg_abi_version": (ctypes.c_int, []),
    "wsmg_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "wsmg_debug_switches": (None, [ctypes.c_int, ctypes.c_int]),
    "wsmg_scratch_bytes": (ctypes.c_size_t, [_DP]),
    "wsmg_scratch_flags_offset": (ctypes.c_size_t, [_DP]),
    "wsmg_map_update": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_size_t, _DP, _P]),
    "wsmg_map_update_ex": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _OP, _P, ctypes.c_size_t, _DP, _P]),
    "wsmg_unproject_index": (ctypes.c_int, [_P, _P, _P, _DP, _P]),
    "wsmg_scatter_max": (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_size_t, _DP, _P]),
    "wsmg_register_fuse_retrieve": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_size_t, _DP, _P]),
    "wsmg_base_coords_host": (ctypes.c_int, [_P, ctypes.c_int32]),
    "wsmg_host_staging_bytes": (ctypes.c_size_t, [_DP, ctypes.c_int32]),
    "wsmg_map_update_host": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_size_t, ctypes.c_int32, _DP, _P]),
    "wsmg_semantic_crop": (ctypes.c_int, [_P, _P, _P, _P, _P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                          ctypes.c_int32, _P]),
    "wsmg_host_live_rows": (ctypes.c_int, [_P, _DP, _P, _P]),
    "wsmg_map_update_host_ex": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_size_t, ctypes.c_int32, _DP,
                                               ctypes.c_uint32, _P]),
}

ABI_VERSION = 3
FLAG_INVALID_PIXEL = 1
FLAG_OUTSIDE_FAN = 2
FLAG_BAD_SLOT = 4

_lib = None


HOST_ZEROCOPY_FEATURES = 1   # include/wsmg.h: WSMG_HOST_ZEROCOPY_FEATURES
HOST_SKIP_DEAD_ROWS = 2      # include/wsmg.h: WSMG_HOST_SKIP_DEAD_ROWS


class WsmgError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libwsmg.so.  Where nvcc exists (the build container) the library is first rebuilt in-tree if it is
    missing or was built from other sources (build.build_cuda compares a content hash of csrc/ + include/ with the
    stamp next to the .so; a no-op when they agree); on a box without nvcc the shipped .so is used as it is, and a
    missing one raises -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("WSMG_LIB_PATH", LIB_PATH)    # override: profiling builds (build.py --phase-skip)
    if path == LIB_PATH:
        import shutil
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            from .build import build_cuda
            build_cuda()
    if not os.path.exists(path):
        raise WsmgError(f"{path} is missing and cannot be built here (no nvcc): run python ws-mgmap_b200/build.py")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.wsmg_abi_version() != ABI_VERSION:
        raise WsmgError(f"libwsmg ABI {lib.wsmg_abi_version()} != {ABI_VERSION} (stale build? python ws-mgmap_b200/build.py -f)")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().wsmg_error_string(rc)
        raise WsmgError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def make_dims(bs, n_maps, c, hf, wf, hd, wd, e, g, resolution, c_in=0, feat_nhwc=0) -> WsmgDims:
    return WsmgDims(int(bs), int(n_maps), int(c), int(hf), int(wf), int(hd), int(wd), int(e), int(g), float(resolution),
                    int(c_in), int(feat_nhwc))
