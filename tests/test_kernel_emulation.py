"""The CUDA kernels' CTA body and pixel math, compiled for the host (csrc/wsmg_emul.cpp),
against the reference goldens and the elementwise spec.  Catches index / layout / schedule
bugs here, where there is no GPU; the -m gpu tests then check the real kernels."""
import hashlib
import os

import numpy as np
import pytest
import torch

from emul import emul_cells, emul_step, lib
from oracle.mapping_oracle import MapGeometry, spec_cells, spec_step
import wsmgmap_b200  # noqa: F401
from wsmgmap_b200._lib import make_dims
from wsmgmap_b200.synth import RandomWalk, make_depth, make_features


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _trig(compass):
    c = torch.as_tensor(compass)[:, 0]
    return np.stack([torch.cos(-c).numpy(), torch.sin(-c).numpy(), torch.cos(c).numpy(), torch.sin(c).numpy()], 1)


def _trig_dict(tr):
    return dict(neg=(tr[:, 0], tr[:, 1]), pos=(tr[:, 2], tr[:, 3]))


def test_golden_trajectory_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "traj_small.npz"))
    bs, c, steps, hf = int(g["bs"]), int(g["c"]), int(g["steps"]), int(g["hf"])
    gmap = np.zeros((bs, 240, 240, c), np.float32)
    for t in range(steps):
        trig = np.stack([g[f"cosneg{t}"], g[f"sinneg{t}"], g[f"cospos{t}"], g[f"sinpos{t}"]], 1)
        lin, inv, codes = emul_cells(g[f"depth{t}"], hf)
        assert np.array_equal(lin.astype(np.int16), g[f"lin{t}"])
        assert np.array_equal(np.packbits(inv), g[f"invalid{t}"])
        assert not (codes == 0xFFFE).any()
        ego, proj = emul_step(gmap, g[f"feat{t}"], g[f"depth{t}"], g[f"gps{t}"], g[f"compass{t}"], g[f"masks{t}"],
                              trig=trig, want_proj=True)
        assert np.array_equal(proj, g[f"proj{t}"])
        assert np.array_equal(ego, g[f"ego{t}"])
        assert _sha(gmap) == str(g[f"mapsha{t}"])


def test_golden_frame_real_shapes(golden_dir):
    g = np.load(os.path.join(golden_dir, "frame_real.npz"))
    bs, c, hf, hd = int(g["bs"]), int(g["c"]), int(g["hf"]), int(g["hd"])
    gen = torch.Generator().manual_seed(int(g["seed"]))
    feat = make_features(bs, c, hf, hf, gen).numpy()
    depth = make_depth("uniform", bs, hd, hd, gen)[..., 0].numpy()
    if _sha(feat) != str(g["feat_sha"]):
        pytest.skip("torch RNG stream differs from the golden file's")
    trig = np.stack([g["cosneg"], g["sinneg"], g["cospos"], g["sinpos"]], 1)
    gmap = np.zeros((bs, 240, 240, c), np.float32)
    ego, proj = emul_step(gmap, feat, depth, g["gps"], g["compass"], np.zeros((bs, 1), np.float32), trig=trig, want_proj=True)
    assert _sha(proj) == str(g["proj_sha"])
    assert np.array_equal(proj[0].argmax(0).astype(np.uint8), g["argmax"])
    assert _sha(ego) == str(g["ego_sha"])
    assert _sha(gmap) == str(g["map_sha"])


@pytest.mark.parametrize("c,hf,hd,e,gl,res", [
    (27, 32, 32, 100, 240, 0.12),     # channels not a multiple of 4: scalar map path, ragged last slab
    (4, 40, 48, 61, 150, 0.2),        # odd ego size, different resolution
    (6, 24, 24, 30, 64, 0.3),         # small even geometry
])
def test_general_geometry_matches_spec(c, hf, hd, e, gl, res):
    geo = MapGeometry(resolution=res, ego=e, glob=gl)
    bs = 2
    gen = torch.Generator().manual_seed(c * 7 + e)
    walk = RandomWalk(bs, seed=e, reset_prob=0.25)
    g_emul = np.zeros((bs, gl, gl, c), np.float32)
    g_spec = np.zeros((bs, gl, gl, c), np.float32)
    for t in range(4):
        gps, compass, masks = walk.step()
        gps = gps * (res / 0.12)
        if t == 3:
            gps[0] += torch.tensor([0.45 * gl * res, -0.48 * gl * res])     # window clipped by the map border
        feat = make_features(bs, c, hf, hf, gen, signed=(t % 2 == 0)).numpy()
        depth = make_depth(("near", "uniform", "room2", "near")[t], bs, hd, hd, gen)[..., 0].numpy() * (res / 0.12)
        trig = _trig(compass)
        lin, inv, _ = emul_cells(depth, hf, e, gl, res)
        slin, sinv = spec_cells(depth, hf, hf, geo)
        assert np.array_equal(lin, slin) and np.array_equal(inv, sinv)
        ego, proj = emul_step(g_emul, feat, depth, gps.numpy(), compass.numpy(), masks.numpy(), trig=trig,
                              want_proj=True, e=e, g=gl, res=res)
        sego, inter = spec_step(g_spec, feat, depth, gps.numpy(), compass.numpy(), masks[:, 0].numpy(), _trig_dict(trig), geo)
        assert np.array_equal(proj, inter["proj"])
        assert np.array_equal(ego, sego)
        assert np.array_equal(g_emul, g_spec)


def test_stage_entry_points_compose():
    """scatter-only then registration-only equals the whole step."""
    bs, c, hf, hd = 2, 8, 48, 64
    gen = torch.Generator().manual_seed(3)
    feat = make_features(bs, c, hf, hf, gen, signed=True).numpy()
    depth = make_depth("near", bs, hd, hd, gen)[..., 0].numpy()
    gps = np.array([[0.3, -0.2], [1.5, 2.5]], np.float32)
    compass = np.array([[0.4], [-2.0]], np.float32)
    masks = np.zeros((bs, 1), np.float32)
    trig = _trig(compass)
    g1 = np.zeros((bs, 240, 240, c), np.float32)
    ego1, _ = emul_step(g1, feat, depth, gps, compass, masks, trig=trig)
    _, proj = emul_step(None, feat, depth, gps, compass, masks, mode=1)
    g2 = np.zeros((bs, 240, 240, c), np.float32)
    ego2, _ = emul_step(g2, None, None, gps, compass, masks, trig=trig, mode=2, proj_in=proj)
    assert np.array_equal(ego1, ego2) and np.array_equal(g1, g2)


def test_shared_memory_plan_fits_b200():
    import ctypes
    d = make_dims(8, 8, 64, 224, 224, 256, 256, 100, 240, 0.12)
    total = lib().wsmg_emul_smem_bytes(ctypes.byref(d))
    assert 0 < total <= 227 * 1024, total
    assert lib().wsmg_emul_fan_cells(ctypes.byref(d)) == 2748


def test_device_trig_tolerance_statement():
    """With sinf/cosf evaluated in-kernel (trig=None) instead of the oracle's values the result moves
    by a few ulp of the rotation matrix; the documented bound is |d| <= 1e-5*|ref| + 2e-5*max|feat|."""
    bs, c, hf, hd = 2, 4, 56, 64
    gen = torch.Generator().manual_seed(9)
    feat = make_features(bs, c, hf, hf, gen).numpy()
    depth = make_depth("room2", bs, hd, hd, gen)[..., 0].numpy()
    gps = np.array([[0.7, 0.1], [-1.2, 0.9]], np.float32)
    compass = np.array([[1.234], [-0.777]], np.float32)
    masks = np.zeros((bs, 1), np.float32)
    ga = np.zeros((bs, 240, 240, c), np.float32)
    gb = np.zeros((bs, 240, 240, c), np.float32)
    ea, _ = emul_step(ga, feat, depth, gps, compass, masks, trig=_trig(compass))
    eb, _ = emul_step(gb, feat, depth, gps, compass, masks, trig=None)       # libm sinf/cosf
    tol = 1e-5 * np.abs(ea) + 2e-5 * np.abs(feat).max()
    assert (np.abs(ea - eb) <= tol).all()


def test_half_output_and_env_slots():
    """wsmg_opts extras: fp16 copy == numpy astype(float16) of the ego map (what
    common_trainer.py:519-520 stores); env_slots == the reference's map[state_index] re-indexing."""
    n, c, hf, hd = 4, 4, 40, 48
    gen = torch.Generator().manual_seed(5)
    def frame(bs, signed=False):
        return (make_features(bs, c, hf, hf, gen, signed=signed).numpy(), make_depth("near", bs, hd, hd, gen)[..., 0].numpy(),
                torch.randn(bs, 2, generator=gen).numpy(), (torch.rand(bs, 1, generator=gen) * 6 - 3).numpy())
    g_ref = np.zeros((n, 240, 240, c), np.float32)
    g_slot = np.zeros((n, 240, 240, c), np.float32)
    f = frame(n)
    emul_step(g_ref, *f, np.zeros((n, 1), np.float32))
    half = np.zeros((n, c, 100, 100), np.uint16)
    ego, _ = emul_step(g_slot, *f, np.zeros((n, 1), np.float32), ego_half=half, env_slots=np.arange(n))
    assert np.array_equal(g_ref, g_slot)
    assert np.array_equal(half.view(np.float16), ego.astype(np.float16))
    keep = [0, 2, 3]                                   # env 1 finished: reference re-materialises map[keep]
    g_ref2 = np.ascontiguousarray(g_ref[keep])
    f = frame(3, signed=True)
    ego_ref, _ = emul_step(g_ref2, *f, np.ones((3, 1), np.float32))
    ego_slot, _ = emul_step(g_slot, *f, np.ones((3, 1), np.float32), env_slots=np.array(keep))
    assert np.array_equal(ego_ref, ego_slot)
    assert np.array_equal(g_slot[keep], g_ref2) and np.array_equal(g_slot[1], g_ref[1])   # paused env untouched


def test_edge_frames_match_oracle():
    """Empty, far, total-collision, boundary-rounding and NaN/inf depth frames (SURVEY 8c edge cases)."""
    from edge_frames import edge_depths
    from oracle.mapping_oracle import OracleMapper
    c, hf, hd = 4, 56, 64
    gen = torch.Generator().manual_seed(8)
    for name, depth in edge_depths(hd).items():
        feat = make_features(1, c, hf, hf, gen, signed=True)
        gps = torch.tensor([[0.4, -0.3]])
        compass = torch.tensor([[0.9]])
        orc = OracleMapper(1, c)
        orc.full_global_map += 0.25                                   # pre-existing map content must survive
        gmap = orc.full_global_map.numpy().copy()
        want = orc.step(feat, depth, gps, compass, torch.ones(1, 1), keep=True)
        lin, inv, _ = emul_cells(depth[..., 0].numpy(), hf)
        assert np.array_equal(lin, orc.last["lin"].numpy()), name
        assert np.array_equal(inv, orc.last["invalid"].numpy()), name
        ego, proj = emul_step(gmap, feat.numpy(), depth[..., 0].numpy(), gps.numpy(), compass.numpy(), np.ones((1, 1), np.float32),
                              trig=_trig(compass), want_proj=True)
        assert np.array_equal(proj, orc.last["proj"].numpy()), name
        assert np.array_equal(ego, want.numpy()), name
        assert np.array_equal(gmap, orc.full_global_map.numpy()), name
        if name in ("empty_all_zero", "all_far"):
            assert not proj.any() and inv.all(), name


@pytest.mark.parametrize("c_in,c_out", [(8, 4), (10, 4), (7, 3), (3, 5), (64, 27)])
def test_fused_channel_pool(c_in, c_out):
    """RGBMapping.forward's adaptive_max_pool1d over channels (rgb_mapping.py:81-84), fused into the scatter."""
    from oracle.mapping_oracle import OracleMapper
    bs, hf, hd = 2, 24, 32
    gen = torch.Generator().manual_seed(c_in * 100 + c_out)
    feat = make_features(bs, c_in, hf, hf, gen, signed=True)
    depth = make_depth("near", bs, hd, hd, gen)
    gps, compass = torch.randn(bs, 2, generator=gen), torch.rand(bs, 1, generator=gen) * 6 - 3
    pooled = torch.nn.functional.adaptive_max_pool1d(feat.permute(0, 2, 3, 1).reshape(bs, -1, c_in), c_out)
    pooled = pooled.reshape(bs, hf, hf, c_out).permute(0, 3, 1, 2).contiguous()
    orc = OracleMapper(bs, c_out)
    want = orc.step(pooled, depth, gps, compass, torch.zeros(bs, 1), keep=True)
    gmap = np.zeros((bs, 240, 240, c_out), np.float32)
    ego, proj = emul_step(gmap, feat.numpy(), depth[..., 0].numpy(), gps.numpy(), compass.numpy(), np.zeros((bs, 1), np.float32),
                          trig=_trig(compass), want_proj=True, map_depth=c_out)
    assert np.array_equal(proj, orc.last["proj"].numpy())
    assert np.array_equal(ego, want.numpy()) and np.array_equal(gmap, orc.full_global_map.numpy())


@pytest.mark.parametrize("e,gl", [(100, 240), (61, 150), (30, 64)])
def test_first_rotation_column_bounds_cut_nothing(e, gl):
    """The kernel skips the cells of the first rotation that cannot see the fan (per-row column bounds).  With
    EVERY fan cell occupied, any cell it wrongly skipped would come out 0 instead of > 0: sweep headings
    (multiples of 15 and 45 degrees, +-pi, tiny angles, random ones) against the elementwise spec."""
    from oracle.mapping_oracle import spec_rotate, spec_translate, spec_gps_cell
    geo = MapGeometry(resolution=0.12, ego=e, glob=gl)
    rng = np.random.default_rng(e)
    proj = np.zeros((1, 4, e, e), np.float32)
    for y in range(e // 2 + 1):                      # the whole packed fan (csrc/wsmg_math.h: fan_x_lo / fan_x_hi)
        proj[0, :, y, max(y - 2, 0):min(e - y + 1, e - 1) + 1] = rng.uniform(0.5, 1.0, size=(4, 1)).astype(np.float32)
    angles = [k * np.pi / 12 for k in range(-12, 13)] + [1e-7, -1e-7, 1e-3, np.pi - 1e-6, -np.pi + 1e-6]
    angles += list(rng.uniform(-np.pi, np.pi, 40))
    lo, hi = geo.paste_lo, geo.paste_hi
    for a in angles:
        compass = np.array([[a]], np.float32)
        gps = rng.uniform(-1.0, 1.0, size=(1, 2)).astype(np.float32)
        trig = _trig(torch.from_numpy(compass))
        gmap = np.zeros((1, gl, gl, 4), np.float32)
        ego, _ = emul_step(gmap, None, None, gps, compass, np.zeros((1, 1), np.float32), trig=trig, mode=2, proj_in=proj,
                           e=e, g=gl)
        t = _trig_dict(trig)
        rot = spec_rotate(proj[0], t["neg"][0][0], t["neg"][1][0])
        canvas = np.zeros((4, gl, gl), np.float32)
        canvas[:, lo:hi, lo:hi] = rot
        gxc, gyc = spec_gps_cell(gps, geo)
        half = np.float32(gl // 2)
        tx = -(((gyc[0] - half).astype(np.float32)) / half).astype(np.float32)
        ty = -(((gxc[0] - half).astype(np.float32)) / half).astype(np.float32)
        fused = np.maximum(0, spec_translate(canvas, tx, ty).transpose(1, 2, 0))
        assert np.array_equal(gmap[0], fused), a
        back = spec_translate(np.ascontiguousarray(fused.transpose(2, 0, 1)), -tx, -ty)
        assert np.array_equal(ego[0], spec_rotate(back[:, lo:hi, lo:hi], t["pos"][0][0], t["pos"][1][0])), a


def test_random_small_geometries_match_spec():
    """Property test over geometry: ego sizes below, at and above the band size, odd sizes, global maps barely larger
    than the ego grid, channel counts around the slab size, poses that push the window over the map border --
    the emulated CTA body (run-time geometry build) against the elementwise spec, two steps each."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=80, deadline=None)
    @given(e=st.integers(6, 64), extra=st.integers(0, 60), c=st.integers(1, 9), hq=st.integers(2, 10),
           kd=st.sampled_from([1.0, 1.25, 2.0]), seed=st.integers(0, 10_000))
    def run(e, extra, c, hq, kd, seed):
        gl = e + extra
        hf = 4 * hq                                   # Hf*Wf must be a multiple of 4
        hd = int(round(hf * kd))
        res = 0.2
        geo = MapGeometry(resolution=res, ego=e, glob=gl)
        rng = np.random.default_rng(seed)
        gen = torch.Generator().manual_seed(seed)
        bs = 2
        g_emul = np.zeros((bs, gl, gl, c), np.float32)
        g_spec = np.zeros((bs, gl, gl, c), np.float32)
        span = gl * res / 2
        for t in range(2):
            gps = rng.uniform(-1.2 * span, 1.2 * span, size=(bs, 2)).astype(np.float32)
            compass = rng.uniform(-np.pi, np.pi, size=(bs, 1)).astype(np.float32)
            masks = (rng.uniform(size=(bs, 1)) > 0.3).astype(np.float32) if t else np.zeros((bs, 1), np.float32)
            feat = make_features(bs, c, hf, hf, gen, signed=bool(seed & 1)).numpy()
            depth = (make_depth(("near", "room2")[t], bs, hd, hd, gen)[..., 0].numpy() * (e * res / 12.0)).astype(np.float32)
            trig = _trig(torch.from_numpy(compass))
            ego, proj = emul_step(g_emul, feat, depth, gps, compass, masks, trig=trig, want_proj=True, e=e, g=gl, res=res)
            sego, inter = spec_step(g_spec, feat, depth, gps, compass, masks[:, 0], _trig_dict(trig), geo)
            assert np.array_equal(proj, inter["proj"])
            assert np.array_equal(ego, sego)
            assert np.array_equal(g_emul, g_spec)

    run()



@pytest.mark.parametrize("c,e,gl,hf,hd", [(64, 100, 240, 224, 256), (8, 31, 64, 32, 40)])
def test_channels_last_features_match_nchw(c, e, gl, hf, hd):
    """a channels_last producer ([bs,Hf,Wf,C] in memory, wsmg_dims.feat_nhwc) gives bit-identical maps to
    the NCHW path -- the emulated CTA body at the reference shapes and at a small run-time geometry."""
    bs = 2
    gen = torch.Generator().manual_seed(c + e)
    res = 0.12 * 100 / e
    feat = make_features(bs, c, hf, hf, gen, signed=True).numpy()
    depth = make_depth("near", bs, hd, hd, gen)[..., 0].numpy()
    gps = (torch.randn(bs, 2, generator=gen) * (res / 0.12)).numpy()
    compass = (torch.rand(bs, 1, generator=gen) * 6 - 3).numpy()
    trig = _trig(torch.from_numpy(compass))
    masks = np.zeros((bs, 1), np.float32)
    g0 = np.zeros((bs, gl, gl, c), np.float32)
    g1 = np.zeros((bs, gl, gl, c), np.float32)
    ego0, _ = emul_step(g0, feat, depth, gps, compass, masks, trig=trig, e=e, g=gl, res=res)
    ego1, _ = emul_step(g1, feat, depth, gps, compass, masks, trig=trig, e=e, g=gl, res=res, feat_nhwc=True)
    assert np.array_equal(ego0, ego1) and np.array_equal(g0, g1) and np.abs(ego0).sum() > 0
