"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference file by path.

The file is read from /root/reference in the build container, else from the
verbatim copy __graft_entry__.build() leaves under baseline/_ref/ (git-ignored,
never committed; gpurun ships it to the GPU box so that bench.py's reference arm
and the policy drop-in test can execute the reference's own code there).  The
golden vectors generated from it (tests/golden/, made by oracle/make_golden.py)
and the restatement in oracle/mapping_oracle.py, asserted equal to it here, pin
parity independently of that copy.

The reference module (vlnce_baselines/common/rgb_mapping.py) needs two shims
to execute without Habitat / torch_scatter / a GPU:
  * `torch_scatter.scatter_max` (third-party, pinned 2.0.6 in the reference's
    SETUP.md:55-60, not vendored): stand-in implementing its published
    semantics -- max-reduce along `dim`, positions nobody wrote are 0, the
    arg output holds src.size(dim) for those (the reference discards it,
    rgb_mapping.py:220).
  * `torch.device("cuda", id)` hard-coded at rgb_mapping.py:14 -> CPU.
"""
import importlib.util
import os
import sys
import types

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# Where the unmodified reference lives: the read-only checkout in the build container, else the copy that
# __graft_entry__.build() places under baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun).
REFERENCE_ROOTS = ("/root/reference", os.path.join(_ROOT, "baseline", "_ref"))


def reference_path(rel: str):
    """Absolute path of reference file `rel` (e.g. 'vlnce_baselines/common/rgb_mapping.py'), or None."""
    for root in REFERENCE_ROOTS:
        p = os.path.join(root, rel)
        if os.path.isfile(p):
            return p
    return None


REFERENCE_FILE = reference_path("vlnce_baselines/common/rgb_mapping.py") or "/root/reference/vlnce_baselines/common/rgb_mapping.py"


def reference_available() -> bool:
    return os.path.isfile(REFERENCE_FILE)


def _scatter_max(src, index, dim=-1, out=None, dim_size=None):
    assert out is None
    index = index.expand_as(src)
    shape = list(src.shape)
    shape[dim] = int(dim_size)
    lowest = torch.finfo(src.dtype).min
    res = torch.full(shape, lowest, dtype=src.dtype, device=src.device)
    res.scatter_reduce_(dim, index, src, reduce="amax", include_self=True)
    cnt = torch.zeros(shape, dtype=torch.int32, device=src.device)
    cnt.scatter_add_(dim, index, torch.ones_like(index, dtype=torch.int32))
    res = torch.where(cnt > 0, res, torch.zeros_like(res))
    arg = torch.full(shape, src.size(dim), dtype=torch.long, device=src.device)
    return res, arg


class _Cfg:
    def __init__(self, num_proc, resolution=0.12, egocentric_map_size=100,
                 global_map_size=240, map_depth=64, gpu_id=0):
        self.gpu_id = gpu_id
        self.num_proc = num_proc
        self.resolution = resolution
        self.egocentric_map_size = egocentric_map_size
        self.global_map_size = global_map_size
        self.map_depth = map_depth


def load_reference_module():
    """exec the reference file with the torch_scatter stand-in installed."""
    if not reference_available():
        raise FileNotFoundError(REFERENCE_FILE)
    if "torch_scatter" not in sys.modules:
        ts = types.ModuleType("torch_scatter")
        ts.scatter_max = _scatter_max
        sys.modules["torch_scatter"] = ts
    spec = importlib.util.spec_from_file_location("_ref_rgb_mapping", REFERENCE_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_reference_mapper(num_proc, device="cpu", **kw):
    """Instantiate the reference RGBMapping.  device="cpu": the hard-coded torch.device("cuda", id) of
    rgb_mapping.py:14 is redirected to the CPU while the constructor runs; device="cuda": the file runs as it
    is on cuda:gpu_id (the stock-PyTorch comparator on the GPU box)."""
    mod = load_reference_module()
    if torch.device(device).type == "cuda":
        return mod.RGBMapping(_Cfg(num_proc, gpu_id=torch.device(device).index or 0, **kw)), mod
    real_device = torch.device
    orig = mod.torch.device
    try:
        mod.torch.device = lambda *a, **k: real_device("cpu")
        m = mod.RGBMapping(_Cfg(num_proc, **kw))
    finally:
        mod.torch.device = orig
    return m, mod


# ------------------------------------------------------------------ the ground-truth semantic map sensor
SENSOR_FILE = reference_path("habitat_extensions/sensors.py") or "/root/reference/habitat_extensions/sensors.py"


def sensor_reference_available() -> bool:
    return os.path.isfile(SENSOR_FILE) and reference_available()


class ReferenceSemMapSensor:
    """Runs the UNMODIFIED body of `GtSemanticMapSensor.get_observation` (sensors.py:383-410) without Habitat:
    the method's source is cut out of the reference file with `ast` and compiled as is; the simulator, the episode
    and `np.load` are replaced by stand-ins that hand it the arrays a test chose.  Everything the method computes
    with (`get_grid`, `F.grid_sample`, `F.pad`) is the real thing."""

    def __init__(self, half_size: int = 50):
        import ast
        import numpy as np
        import torch.nn.functional as F
        src = open(SENSOR_FILE).read()
        tree = ast.parse(src)
        fn = None
        for node in ast.walk(tree):
            if isinstance(node, ast.ClassDef) and node.name == "GtSemanticMapSensor":
                for item in node.body:
                    if isinstance(item, ast.FunctionDef) and item.name == "get_observation":
                        fn = item
        assert fn is not None, "GtSemanticMapSensor.get_observation not found in the reference"
        fn.decorator_list = []
        mod = ast.Module(body=[fn], type_ignores=[])
        ast.fix_missing_locations(mod)
        ref_mod = load_reference_module()
        outer = self

        class _Np:                      # numpy, except that np.load returns the map the test supplies
            def __getattr__(self, name):
                return getattr(np, name)

            @staticmethod
            def load(path):
                return outer._map_to_load

        ns = {"torch": torch, "F": F, "np": _Np(), "os": os, "get_grid": ref_mod.get_grid, "Any": object}
        exec(compile(mod, SENSOR_FILE, "exec"), ns)
        self._method = ns["get_observation"]
        self.half_size = half_size
        self.gt_path = "unused"
        self.prev_episode_id = None
        self._map_to_load = None
        self._sim = types.SimpleNamespace(record_heading=0.0, _state=None)
        self._sim.get_agent_state = lambda: self._sim._state

    def observe(self, episode_id, position, heading, new_map=None):
        """position: (x, y, z) of the agent; heading: the simulator's record_heading; new_map: int array [S,S]
        returned by np.load when the episode id changes."""
        self._map_to_load = new_map
        self._sim.record_heading = float(heading)
        self._sim._state = types.SimpleNamespace(position=[float(p) for p in position])
        episode = types.SimpleNamespace(episode_id=episode_id)
        return self._method(self, None, episode)

