"""CPU-only checks of the host side: the C ABI library loads and exports every symbol the
header declares, host helpers agree with torch, and get_grid keeps the reference's contract."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import wsmgmap_b200  # noqa: F401
from wsmgmap_b200 import _lib
from wsmgmap_b200.rgb_mapping import get_grid, to_grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    header = open(os.path.join(ROOT, "include", "wsmg.h")).read()
    declared = set(re.findall(r"\b(wsmg_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/wsmg.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.wsmg_abi_version() == _lib.ABI_VERSION
    assert b"NULL" in lib.wsmg_error_string(-1)


def test_validation_without_gpu():
    lib = _lib.load()
    ok = _lib.make_dims(8, 8, 64, 224, 224, 256, 256, 100, 240, 0.12)
    assert lib.wsmg_scratch_bytes(ctypes.byref(ok)) >= 8 * 224 * 224 * 2
    for bad in (_lib.make_dims(0, 8, 64, 224, 224, 256, 256, 100, 240, 0.12),
                _lib.make_dims(8, 4, 64, 224, 224, 256, 256, 100, 240, 0.12),
                _lib.make_dims(8, 8, 64, 224, 224, 256, 256, 300, 240, 0.12),
                _lib.make_dims(8, 8, 64, 223, 223, 256, 256, 100, 240, 0.12)):
        assert lib.wsmg_scratch_bytes(ctypes.byref(bad)) == 0


@pytest.mark.parametrize("n", [2, 3, 7, 100, 101, 224, 240, 480])
def test_base_coords_match_torch(n):
    out = np.zeros(n, np.float32)
    assert _lib.load().wsmg_base_coords_host(out.ctypes.data_as(ctypes.c_void_p), n) == 0
    want = (torch.linspace(-1, 1, n) * (n - 1) / n).numpy()
    assert np.array_equal(out, want)


def test_get_grid_contract():
    pose = torch.tensor([[0.25, -0.5, 0.3], [0.0, 0.0, -2.0]])
    rot, trans = get_grid(pose, (2, 1, 12, 12), "cpu")
    assert rot.shape == (2, 12, 12, 2) and trans.shape == (2, 12, 12, 2)
    base = torch.linspace(-1, 1, 12) * 11 / 12
    assert torch.equal(trans[0, :, :, 0], (base + 0.25)[None, :].expand(12, 12))
    assert torch.equal(trans[0, :, :, 1], (base - 0.5)[:, None].expand(12, 12))
    from oracle.reference_loader import load_reference_module, reference_available
    if reference_available():
        r2, t2 = load_reference_module().get_grid(pose, (2, 1, 12, 12), "cpu")
        assert torch.equal(rot, r2) and torch.equal(trans, t2)
    tg = to_grid(240, -14.399999999999999, 14.399999999999999)
    gx, gy = tg.get_grid_coords(torch.tensor([[0.0, 0.0], [1.37, -2.21]]))
    assert gx.tolist() == [120.0, 109.0] and gy.tolist() == [120.0, 102.0]


def test_product_path_has_no_oracle_import():
    pkg = os.path.join(ROOT, "ws-mgmap_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_host_live_rows_match_the_oracle():
    """wsmg_host_live_rows (pure host code of libwsmg.so, what WSMG_HOST_SKIP_DEAD_ROWS copies by) against the rows
    in which the oracle finds a pixel that can write."""
    import numpy as np
    import torch
    from oracle.mapping_oracle import MapGeometry, spec_cells
    from wsmgmap_b200 import ops
    from wsmgmap_b200.synth import make_depth
    gen = torch.Generator().manual_seed(5)
    hf, hd = 224, 256
    depth = torch.cat([make_depth(k, 2, hd, hd, gen) for k in ("room2", "room4", "near", "uniform")], 0)
    depth[1] = 1.0                                   # nothing can write
    depth[2, :200] = 0.0                             # holes on top
    d = ops.dims_for((depth.shape[0], 64, hf, hf), depth.shape, depth.shape[0])
    lo, hi = ops.host_live_rows(depth, d)
    _, invalid = spec_cells(depth[..., 0].numpy(), hf, hf, MapGeometry())
    for b in range(depth.shape[0]):
        rows = np.nonzero((~invalid[b]).any(axis=1))[0]
        if rows.size == 0:
            assert lo[b] > hi[b]
        else:
            assert (int(lo[b]), int(hi[b])) == (int(rows[0]), int(rows[-1])), b


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the native one) needs no GPU and must print
    exactly one JSON line on stdout with the contract's keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    from oracle.reference_loader import reference_available
    # the unmodified reference file when it is there (this container, or baseline/_ref/ on the GPU box), else the port
    assert d["cpu_baseline"]["kind"] == ("reference" if reference_available() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert "8 envs per step" in d["cpu_baseline"]["sample"] and d["config"]["scaling"] == "strong"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
