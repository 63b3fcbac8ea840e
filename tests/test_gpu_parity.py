"""Parity of the CUDA path (through the C ABI, ws-mgmap_b200/ops.py) with the oracle and the
reference goldens.  Bars (BASELINE.json north_star):
  * cell indices, invalid flags, occupancy bits, pre-rotation channel argmax: bit-exact
    (argmax tie-break: lowest channel index; empty cell -> label 0);
  * float maps: bit-exact when the oracle's sin/cos are supplied (`trig`), and within
    |d| <= 1e-5*|ref| + 2e-5*max|feat| when sinf/cosf are evaluated on the device.
"""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle.mapping_oracle import MapGeometry, OracleMapper, spec_cells
import wsmgmap_b200  # noqa: F401
from wsmgmap_b200 import ops
from wsmgmap_b200.synth import RandomWalk, make_depth, make_features

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _trig(compass):
    c = compass[:, 0].cpu()
    return torch.stack([torch.cos(-c), torch.sin(-c), torch.cos(c), torch.sin(c)], 1).contiguous()


def _cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).to(DEV)


def test_golden_trajectory_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "traj_small.npz"))
    bs, c, steps, hf = int(g["bs"]), int(g["c"]), int(g["steps"]), int(g["hf"])
    gmap = torch.zeros(bs, 240, 240, c, device=DEV)
    for t in range(steps):
        feat, depth = _cu(g[f"feat{t}"]), _cu(g[f"depth{t}"]).unsqueeze(-1).contiguous()
        trig = _cu(np.stack([g[f"cosneg{t}"], g[f"sinneg{t}"], g[f"cospos{t}"], g[f"sinpos{t}"]], 1))
        lin, inv = ops.unproject_index(depth, hf, hf)
        assert np.array_equal(lin.cpu().numpy().astype(np.int16), g[f"lin{t}"])
        assert np.array_equal(np.packbits(inv.cpu().numpy()), g[f"invalid{t}"])
        proj = ops.scatter_max(feat, depth)
        assert np.array_equal(proj.cpu().numpy(), g[f"proj{t}"])
        ego = ops.map_update(feat, depth, _cu(g[f"gps{t}"]), _cu(g[f"compass{t}"]), _cu(g[f"masks{t}"]), gmap, trig=trig)
        assert np.array_equal(ego.cpu().numpy(), g[f"ego{t}"]), f"step {t}"
        assert _sha(gmap.cpu().numpy()) == str(g[f"mapsha{t}"]), f"step {t}"
    assert np.array_equal(gmap.cpu().numpy(), g["map_final"])


def test_golden_frame_real_shapes(golden_dir):
    g = np.load(os.path.join(golden_dir, "frame_real.npz"))
    bs, c, hf, hd = int(g["bs"]), int(g["c"]), int(g["hf"]), int(g["hd"])
    gen = torch.Generator().manual_seed(int(g["seed"]))
    feat = make_features(bs, c, hf, hf, gen)
    depth = make_depth("uniform", bs, hd, hd, gen)
    if _sha(feat.numpy()) != str(g["feat_sha"]):
        pytest.skip("torch RNG stream differs from the golden file's")
    trig = _cu(np.stack([g["cosneg"], g["sinneg"], g["cospos"], g["sinpos"]], 1))
    gmap = torch.zeros(bs, 240, 240, c, device=DEV)
    lin, inv = ops.unproject_index(depth.to(DEV), hf, hf)
    assert np.array_equal(lin.cpu().numpy().astype(np.int16), g["lin"])
    assert np.array_equal(np.packbits(inv.cpu().numpy()), g["invalid"])
    proj = ops.scatter_max(feat.to(DEV), depth.to(DEV)).cpu().numpy()
    assert _sha(proj) == str(g["proj_sha"])
    assert np.array_equal(proj[0].argmax(0).astype(np.uint8), g["argmax"])
    assert np.array_equal(np.packbits((proj[0] != 0).any(0)), g["occupied"])
    ego = ops.map_update(feat.to(DEV), depth.to(DEV), _cu(g["gps"]), _cu(g["compass"]), torch.zeros(bs, 1, device=DEV),
                         gmap, trig=trig)
    assert _sha(ego.cpu().numpy()) == str(g["ego_sha"])
    assert _sha(gmap.cpu().numpy()) == str(g["map_sha"])


def test_trajectory_100_steps_vs_oracle():
    """BASELINE config 3: 100-step random-walk trajectory, 8 envs, real shapes, checked every step."""
    bs, c, hf, hd, steps = 8, 64, 224, 256, 100
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    orc = OracleMapper(bs, c)
    gmap = torch.zeros(bs, 240, 240, c, device=DEV)
    gmap_dev_trig = torch.zeros(bs, 240, 240, c, device=DEV)
    walk = RandomWalk(bs, seed=42, far_env=7)
    resets = {b: 20 + 9 * b for b in range(bs)}
    kinds = ("uniform", "near", "room2", "room4")
    worst = 0.0
    scale = 0.0            # the map keeps values of earlier frames: tolerances scale with the largest |feature| seen so far
    for t in range(steps):
        gps, compass, masks = walk.step()
        for b, when in resets.items():
            if t == when:
                masks[b] = 0.0
        gen = torch.Generator().manual_seed(1000 * t)
        feat = make_features(bs, c, hf, hf, gen, signed=(t % 10 == 3))
        depth = torch.cat([make_depth(kinds[(t + b) % 4], 1, hd, hd, gen) for b in range(bs)], 0)
        want = orc.step(feat, depth, gps, compass, masks, keep=True)
        fd, dd = feat.to(DEV), depth.to(DEV)
        lin, inv = ops.unproject_index(dd, hf, hf)
        assert torch.equal(lin.cpu().long(), orc.last["lin"]), f"step {t}: cell indices"
        assert torch.equal(inv.cpu(), orc.last["invalid"]), f"step {t}: invalid flags"
        if t % 10 == 0:
            proj = ops.scatter_max(fd, dd).cpu()
            assert torch.equal(proj, orc.last["proj"]), f"step {t}: projected grid"
            assert torch.equal(proj.argmax(1), orc.last["proj"].argmax(1)), f"step {t}: channel argmax"
            occ = (proj != 0).any(1).reshape(bs, -1)
            assert torch.equal(occ, orc.last["occ"]), f"step {t}: occupancy"   # U(0,1)/N(0,1) features are never exactly 0
        ego = ops.map_update(fd, dd, gps.to(DEV), compass.to(DEV), masks.to(DEV), gmap, trig=_trig(compass).to(DEV))
        assert torch.equal(ego.cpu(), want), f"step {t}: ego map (oracle trig)"
        assert torch.equal(gmap.cpu(), orc.full_global_map), f"step {t}: global map (oracle trig)"
        ego2 = ops.map_update(fd, dd, gps.to(DEV), compass.to(DEV), masks.to(DEV), gmap_dev_trig)
        scale = max(scale, feat.abs().max().item())
        tol = 1e-5 * want.abs() + 2e-5 * scale
        err = (ego2.cpu() - want).abs()
        assert (err <= tol).all(), f"step {t}: ego map (device trig) max err {err.max().item()}"
        gerr = (gmap_dev_trig.cpu() - orc.full_global_map).abs()
        gtol = 1e-5 * orc.full_global_map.abs() + 2e-5 * scale
        assert (gerr <= gtol).all(), f"step {t}: global map (device trig) max err {gerr.max().item()} ratio {(gerr / gtol).max().item()}"
        worst = max(worst, err.max().item())
    print(f"device-trig worst abs deviation over the trajectory: {worst:.3e}")


@pytest.mark.parametrize("c,hf,hd,e,gl,res", [(27, 32, 32, 100, 240, 0.12), (4, 40, 48, 61, 150, 0.2), (64, 256, 256, 100, 240, 0.12),
                                                (8, 48, 64, 101, 260, 0.1), (4, 24, 24, 30, 64, 0.3), (12, 56, 64, 96, 208, 0.15),
                                                (27, 256, 256, 100, 240, 0.12), (6, 224, 256, 100, 240, 0.12)])
# 101: the largest ego grid one SM holds; the last two: C % 4 != 0 at the reference's grid sizes (compile-time-geometry builds of their own)
def test_other_geometries_vs_spec(c, hf, hd, e, gl, res):
    from oracle.mapping_oracle import spec_step
    geo = MapGeometry(resolution=res, ego=e, glob=gl)
    bs = 2
    gen = torch.Generator().manual_seed(c + e)
    walk = RandomWalk(bs, seed=e, reset_prob=0.25)
    gmap = torch.zeros(bs, gl, gl, c, device=DEV)
    g_spec = np.zeros((bs, gl, gl, c), np.float32)
    for t in range(3):
        gps, compass, masks = walk.step()
        gps = gps * (res / 0.12)
        feat = make_features(bs, c, hf, hf, gen, signed=(t == 1))
        depth = make_depth(("near", "uniform", "room2")[t], bs, hd, hd, gen) * (res / 0.12)
        trig = _trig(compass)
        lin, inv = ops.unproject_index(depth.to(DEV), hf, hf, e, gl, res)
        slin, sinv = spec_cells(depth[..., 0].numpy(), hf, hf, geo)
        assert np.array_equal(lin.cpu().numpy(), slin) and np.array_equal(inv.cpu().numpy(), sinv)
        ego = ops.map_update(feat.to(DEV), depth.to(DEV), gps.to(DEV), compass.to(DEV), masks.to(DEV), gmap, e=e,
                             resolution=res, trig=trig.to(DEV))
        tr = trig.numpy()
        sego, _ = spec_step(g_spec, feat.numpy(), depth[..., 0].numpy(), gps.numpy(), compass.numpy(), masks[:, 0].numpy(),
                            dict(neg=(tr[:, 0], tr[:, 1]), pos=(tr[:, 2], tr[:, 3])), geo)
        assert np.array_equal(ego.cpu().numpy(), sego)
        assert np.array_equal(gmap.cpu().numpy(), g_spec)


def test_stage_composition_and_host_pipeline():
    bs, c, hf, hd = 5, 64, 224, 256
    gen = torch.Generator().manual_seed(77)
    feat = make_features(bs, c, hf, hf, gen)
    depth = make_depth("room4", bs, hd, hd, gen)
    gps = torch.randn(bs, 2, generator=gen)
    compass = torch.rand(bs, 1, generator=gen) * 6 - 3
    masks = torch.zeros(bs, 1)
    orc = OracleMapper(bs, c)
    want = orc.step(feat, depth, gps, compass, masks)        # every leg below is held against the oracle (device sin / cos: tolerance)

    def close(a, b):
        return ((a - b).abs() <= 1e-5 * b.abs() + 2e-5).all()

    g1 = torch.zeros(bs + 2, 240, 240, c, device=DEV)        # bs < n_maps: rows [bs:] stay untouched
    g1[bs:] = 3.0
    ego1 = ops.map_update(feat.to(DEV), depth.to(DEV), gps.to(DEV), compass.to(DEV), masks.to(DEV), g1)
    assert (g1[bs:] == 3.0).all()
    assert close(ego1.cpu(), want) and close(g1[:bs].cpu(), orc.full_global_map)
    proj = ops.scatter_max(feat.to(DEV), depth.to(DEV))
    g2 = torch.zeros(bs, 240, 240, c, device=DEV)
    ego2 = ops.register_fuse_retrieve(proj, gps.to(DEV), compass.to(DEV), masks.to(DEV), g2)
    assert torch.equal(ego1, ego2) and torch.equal(g1[:bs], g2)
    orc.step(feat, depth, gps, compass, masks, keep=True)    # (masks == 0: the same step again)
    assert torch.equal(proj.cpu(), orc.last["proj"])         # the stage output itself is bit-exact: no trigonometry involved
    # host-buffer entry (pinned), chunked: same result
    d = ops.dims_for(feat.shape, depth.shape, bs)
    pipe = ops.HostPipeline(d, DEV, chunk_envs=2)
    g3 = torch.zeros(bs, 240, 240, c, device=DEV)
    ego_h = torch.empty(bs, c, 100, 100).pin_memory()
    pipe.step(feat.pin_memory(), depth.pin_memory(), gps.pin_memory(), compass.pin_memory(), masks.pin_memory(), g3, ego_h)
    torch.cuda.synchronize()
    assert torch.equal(ego_h, ego1.cpu()) and torch.equal(g3, g2)
    assert close(ego_h, want) and close(g3.cpu(), orc.full_global_map)
    # zero-copy features: the scatter reads the pinned host tensor itself; pageable memory is refused
    pipe0 = ops.HostPipeline(d, DEV, chunk_envs=3, zero_copy=True)
    g4 = torch.zeros(bs, 240, 240, c, device=DEV)
    ego_h.zero_()
    pipe0.step(feat.pin_memory(), depth.pin_memory(), gps.pin_memory(), compass.pin_memory(), masks.pin_memory(), g4, ego_h)
    torch.cuda.synchronize()
    assert torch.equal(ego_h, ego1.cpu()) and torch.equal(g4, g2)
    assert close(ego_h, want) and close(g4.cpu(), orc.full_global_map)
    # only the feature rows that hold a pixel which can write cross the bus; the rest of the staging buffer is
    # poisoned with NaN bit patterns to prove the kernel never looks at it
    pipe1 = ops.HostPipeline(d, DEV, chunk_envs=2, skip_dead_rows=True)
    pipe1.staging.fill_(0xFF)
    g5 = torch.zeros(bs, 240, 240, c, device=DEV)
    ego_h.zero_()
    pipe1.step(feat.pin_memory(), depth.pin_memory(), gps.pin_memory(), compass.pin_memory(), masks.pin_memory(), g5, ego_h)
    torch.cuda.synchronize()
    assert torch.equal(ego_h, ego1.cpu()) and torch.equal(g5, g2)
    assert close(ego_h, want) and close(g5.cpu(), orc.full_global_map)
    dead = depth.clone()
    dead[1] = 1.0                                            # 10 m everywhere: this frame copies no feature row at all
    pipe1.staging.fill_(0xFF)
    g6 = torch.zeros(bs, 240, 240, c, device=DEV)
    pipe1.step(feat.pin_memory(), dead.pin_memory(), gps.pin_memory(), compass.pin_memory(), masks.pin_memory(), g6, ego_h)
    g7 = torch.zeros(bs, 240, 240, c, device=DEV)
    ego7 = ops.map_update(feat.to(DEV), dead.to(DEV), gps.to(DEV), compass.to(DEV), masks.to(DEV), g7)
    torch.cuda.synchronize()
    assert torch.equal(ego_h, ego7.cpu()) and torch.equal(g6, g7)
    orc_dead = OracleMapper(bs, c)
    want_dead = orc_dead.step(feat, dead, gps, compass, masks)
    assert close(ego_h, want_dead) and close(g6.cpu(), orc_dead.full_global_map)
    from wsmgmap_b200._lib import WsmgError
    with pytest.raises(WsmgError):
        pipe0.step(feat.clone(), depth.pin_memory(), gps.pin_memory(), compass.pin_memory(), masks.pin_memory(), g4, ego_h)


def test_host_pipeline_without_batched_copies():
    """The host-buffer entry submits a chunk's large copies as one cudaMemcpyBatchAsync when the runtime has it; an older
    runtime (or WSMG_HOST_NO_BATCHCOPY=1, read once per process -- hence the subprocess) takes one call per copy.  Both ways,
    in all three copy modes, the ego map and the global map are the direct update's, bit for bit."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, torch
sys.path.insert(0, %r)
import wsmgmap_b200
from wsmgmap_b200 import ops
from wsmgmap_b200.synth import make_depth, make_features
dev = torch.device('cuda', 0)
bs, c, hf, hd = 7, 64, 224, 256
gen = torch.Generator().manual_seed(5)
feat = make_features(bs, c, hf, hf, gen); depth = make_depth('room2', bs, hd, hd, gen)
depth[3] = make_depth('uniform', 1, hd, hd, gen)[0]
gps = torch.randn(bs, 2, generator=gen); compass = torch.rand(bs, 1, generator=gen) * 6 - 3; masks = torch.zeros(bs, 1)
g0 = torch.zeros(bs, 240, 240, c, device=dev)
ego0 = ops.map_update(feat.to(dev), depth.to(dev), gps.to(dev), compass.to(dev), masks.to(dev), g0)
d = ops.dims_for(feat.shape, depth.shape, bs)
for kw in (dict(), dict(skip_dead_rows=True), dict(zero_copy=True)):
    pipe = ops.HostPipeline(d, dev, chunk_envs=3, **kw)
    g = torch.zeros(bs, 240, 240, c, device=dev)
    ego_h = torch.empty(bs, c, 100, 100).pin_memory()
    pipe.step(feat.pin_memory(), depth.pin_memory(), gps.pin_memory(), compass.pin_memory(), masks.pin_memory(), g, ego_h)
    torch.cuda.synchronize()
    assert torch.equal(ego_h, ego0.cpu()) and torch.equal(g, g0), kw
print('ok')
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for env in ({}, {"WSMG_HOST_NO_BATCHCOPY": "1"}):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (env, r.stdout[-500:], r.stderr[-1500:])


def test_full_size_properties():
    """Size-independent properties at a batch the oracle would need minutes for (256 envs):
    max-fusion is idempotent, the map is monotone, a reset forgets everything, batch order is irrelevant."""
    bs, c, hf, hd = 256, 64, 224, 256
    gen = torch.Generator(device=DEV).manual_seed(5)
    feat = torch.rand(bs, c, hf, hf, generator=gen, device=DEV)
    depth = torch.rand(bs, hd, hd, 1, generator=gen, device=DEV) * 0.6
    gps = torch.randn(bs, 2, generator=gen, device=DEV) * 3
    compass = torch.rand(bs, 1, generator=gen, device=DEV) * 6 - 3
    ones, zeros = torch.ones(bs, 1, device=DEV), torch.zeros(bs, 1, device=DEV)
    gmap = torch.zeros(bs, 240, 240, c, device=DEV)
    ego_a = ops.map_update(feat, depth, gps, compass, zeros, gmap)
    snap = gmap.clone()
    ego_b = ops.map_update(feat, depth, gps, compass, ones, gmap)          # same frame again
    assert torch.equal(gmap, snap) and torch.equal(ego_a, ego_b)            # idempotent
    feat2 = torch.rand(bs, c, hf, hf, generator=gen, device=DEV)
    ops.map_update(feat2, depth, gps + 0.3, compass + 0.2, ones, gmap)
    assert (gmap >= snap).all() and (gmap >= 0).all()                         # monotone, non-negative
    ego_c = ops.map_update(feat, depth, gps, compass, zeros, gmap)          # reset forgets
    assert torch.equal(gmap, snap) and torch.equal(ego_c, ego_a)
    perm = torch.randperm(bs, device=DEV)
    g2 = torch.zeros(bs, 240, 240, c, device=DEV)
    ego_p = ops.map_update(feat[perm].contiguous(), depth[perm].contiguous(), gps[perm].contiguous(),
                           compass[perm].contiguous(), zeros, g2)
    assert torch.equal(ego_p, ego_a[perm]) and torch.equal(g2, snap[perm])   # envs independent
    # linearity of the bilinear stages in the features: scaling features by 2 scales everything by 2 exactly
    g3 = torch.zeros(bs, 240, 240, c, device=DEV)
    ego_2 = ops.map_update(feat * 2, depth, gps, compass, zeros, g3)
    assert torch.equal(ego_2, ego_a * 2) and torch.equal(g3, snap * 2)


def test_error_codes():
    import ctypes
    from wsmgmap_b200 import _lib
    lib = _lib.load()
    d = _lib.make_dims(1, 1, 64, 224, 224, 256, 256, 100, 240, 0.12)
    assert lib.wsmg_map_update(None, None, None, None, None, None, None, None, None, 0, ctypes.byref(d), None) == -1
    bad = _lib.make_dims(1, 1, 64, 224, 224, 256, 256, 300, 240, 0.12)
    assert lib.wsmg_scratch_bytes(ctypes.byref(bad)) == 0
    x = torch.zeros(16, device=DEV)
    assert lib.wsmg_unproject_index(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(x.data_ptr()),
                                    ctypes.c_void_p(x.data_ptr()), ctypes.byref(bad), None) == -3
    feat = torch.zeros(1, 64, 224, 224, device=DEV)                         # a 120 x 120 crop does not fit one SM's shared memory
    depth = torch.zeros(1, 256, 256, 1, device=DEV)
    with pytest.raises(_lib.WsmgError, match="shared memory"):
        ops.map_update(feat, depth, torch.zeros(1, 2, device=DEV), torch.zeros(1, 1, device=DEV),
                       torch.zeros(1, 1, device=DEV), torch.zeros(1, 300, 300, 64, device=DEV), e=120)


def test_kernel_variants_agree():
    """The four builds of the fused kernel (TMA vs cp.async window path, compile-time vs run-time
    geometry) are bit-identical over a short trajectory that includes a border-clipped window."""
    import os
    bs, c, hf, hd = 6, 64, 224, 256
    gen = torch.Generator().manual_seed(21)
    frames = []
    walk = RandomWalk(bs, seed=5, far_env=0)
    for t in range(4):
        gps, compass, masks = walk.step()
        if t >= 2:
            gps[0] += torch.tensor([10.0, -9.0])
        frames.append((make_features(bs, c, hf, hf, gen).to(DEV), make_depth(("room2", "uniform", "near", "room4")[t], bs, hd, hd, gen).to(DEV),
                       gps.to(DEV), compass.to(DEV), masks.to(DEV)))
    results = {}
    from wsmgmap_b200 import _lib
    try:
        for no_tma in ("0", "1"):
            for generic in ("0", "1"):
                _lib.load().wsmg_debug_switches(int(generic), int(no_tma))
                gmap = torch.zeros(bs, 240, 240, c, device=DEV)
                egos = [ops.map_update(f, d, g, cp, m, gmap).clone() for f, d, g, cp, m in frames]
                torch.cuda.synchronize()
                results[(no_tma, generic)] = (egos, gmap)
    finally:
        _lib.load().wsmg_debug_switches(-1, -1)
    ref_egos, ref_map = results[("0", "0")]
    for key, (egos, gmap) in results.items():
        assert torch.equal(gmap, ref_map), key
        for a, b in zip(egos, ref_egos):
            assert torch.equal(a, b), key


def test_edge_frames_match_oracle():
    """Empty, far, total-collision, boundary-rounding and NaN/inf depth frames at the real shapes, plus bs = 1."""
    from edge_frames import edge_depths
    c, hf, hd = 64, 224, 256
    gen = torch.Generator().manual_seed(8)
    for name, depth in edge_depths(hd).items():
        feat = make_features(1, c, hf, hf, gen, signed=(name == "wall_constant"))
        gps = torch.tensor([[0.4, -0.3]])
        compass = torch.tensor([[0.9]])
        orc = OracleMapper(1, c)
        orc.full_global_map += 0.25
        gmap = orc.full_global_map.clone().to(DEV)
        want = orc.step(feat, depth, gps, compass, torch.ones(1, 1), keep=True)
        lin, inv = ops.unproject_index(depth.to(DEV), hf, hf)
        assert torch.equal(lin.cpu().long(), orc.last["lin"]) and torch.equal(inv.cpu(), orc.last["invalid"]), name
        proj = ops.scatter_max(feat.to(DEV), depth.to(DEV))
        assert torch.equal(proj.cpu(), orc.last["proj"]), name
        ego = ops.map_update(feat.to(DEV), depth.to(DEV), gps.to(DEV), compass.to(DEV), torch.ones(1, 1, device=DEV), gmap,
                             trig=_trig(compass).to(DEV))
        assert torch.equal(ego.cpu(), want), name
        assert torch.equal(gmap.cpu(), orc.full_global_map), name


def test_window_touching_the_map_border():
    """Windows whose last row / column is exactly the map's last one (the TMA boxes then reach past the tensor and
    are zero-filled there), one cell inside, and one cell outside (clipped: cp.async path) -- all bit-identical to
    the oracle; and a repeated identical frame changes no byte of the map (max-fusion is idempotent, the kernel
    then stores nothing)."""
    c, hf, hd = 64, 224, 256
    # gps -> window origin: u0 = 51 - 1 + round((14.4 - gps0)/0.12) - 120 ... u0 = 138 <=> gps0 = -8.28, u0 = 0 <=> gps0 = 8.28
    g0 = [8.28, 8.16, -8.28, -8.16, -8.40, 8.40, 0.0, -8.28]
    g1 = [-8.28, 8.28, 8.28, -8.16, 8.40, 0.0, -8.40, -8.28]
    bs = len(g0)
    gen = torch.Generator().manual_seed(77)
    gps = torch.tensor(list(zip(g0, g1)), dtype=torch.float32)
    compass = torch.rand(bs, 1, generator=gen) * 6 - 3
    orc = OracleMapper(bs, c)
    orc.full_global_map += 0.125
    gmap = orc.full_global_map.clone().to(DEV)
    ones = torch.ones(bs, 1)
    for t in range(2):
        feat = make_features(bs, c, hf, hf, gen)
        depth = make_depth(("room4", "near")[t], bs, hd, hd, gen)
        want = orc.step(feat, depth, gps, compass, ones, keep=True)
        ego = ops.map_update(feat.to(DEV), depth.to(DEV), gps.to(DEV), compass.to(DEV), ones.to(DEV), gmap,
                             trig=_trig(compass).to(DEV))
        assert torch.equal(ego.cpu(), want), t
        assert torch.equal(gmap.cpu(), orc.full_global_map), t
    before = gmap.clone()
    ego2 = ops.map_update(feat.to(DEV), depth.to(DEV), gps.to(DEV), compass.to(DEV), ones.to(DEV), gmap,
                          trig=_trig(compass).to(DEV))
    assert torch.equal(gmap, before) and torch.equal(ego2.cpu(), want)


def test_random_small_geometries_vs_spec():
    """A dozen random geometries (ego sizes around the band size, odd sizes, maps barely larger than the ego grid,
    channel counts around the slab size, windows over the map border) through the run-time-geometry kernels,
    bit-exact against the elementwise spec."""
    from oracle.mapping_oracle import spec_step
    rng = np.random.default_rng(2024)
    for case in range(12):
        e = int(rng.integers(6, 64)); gl = e + int(rng.integers(0, 60)); c = int(rng.integers(1, 10))
        hf = 4 * int(rng.integers(2, 11)); hd = int(round(hf * float(rng.choice([1.0, 1.25, 2.0])))); res = 0.2
        geo = MapGeometry(resolution=res, ego=e, glob=gl)
        gen = torch.Generator().manual_seed(case)
        bs = 2
        gmap = torch.zeros(bs, gl, gl, c, device=DEV)
        g_spec = np.zeros((bs, gl, gl, c), np.float32)
        span = gl * res / 2
        for t in range(2):
            gps = torch.from_numpy(rng.uniform(-1.2 * span, 1.2 * span, size=(bs, 2)).astype(np.float32))
            compass = torch.from_numpy(rng.uniform(-np.pi, np.pi, size=(bs, 1)).astype(np.float32))
            masks = torch.from_numpy((rng.uniform(size=(bs, 1)) > 0.3).astype(np.float32)) if t else torch.zeros(bs, 1)
            feat = make_features(bs, c, hf, hf, gen, signed=bool(case & 1))
            depth = make_depth(("near", "room2")[t], bs, hd, hd, gen) * (e * res / 12.0)
            trig = _trig(compass)
            ego = ops.map_update(feat.to(DEV), depth.to(DEV), gps.to(DEV), compass.to(DEV), masks.to(DEV), gmap, e=e,
                                 resolution=res, trig=trig.to(DEV))
            tr = trig.numpy()
            sego, _ = spec_step(g_spec, feat.numpy(), depth[..., 0].numpy(), gps.numpy(), compass.numpy(), masks[:, 0].numpy(),
                                dict(neg=(tr[:, 0], tr[:, 1]), pos=(tr[:, 2], tr[:, 3])), geo)
            assert np.array_equal(ego.cpu().numpy(), sego), (case, e, gl, c, hf, hd, t)
            assert np.array_equal(gmap.cpu().numpy(), g_spec), (case, e, gl, c, hf, hd, t)
