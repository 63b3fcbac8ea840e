"""The oracle (both layers) against golden vectors produced by the reference
file itself (oracle/make_golden.py).  Runs anywhere -- this is what pins the
oracle on the GPU box."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle.mapping_oracle import MapGeometry, OracleMapper, spec_cells, spec_scatter, spec_step
import wsmgmap_b200  # noqa: F401
from wsmgmap_b200.synth import make_depth, make_features


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def traj(golden_dir):
    return np.load(os.path.join(golden_dir, "traj_small.npz"))


def test_trajectory_small_oracle(traj):
    bs, c, steps = int(traj["bs"]), int(traj["c"]), int(traj["steps"])
    orc = OracleMapper(bs, c)
    for t in range(steps):
        feat = torch.from_numpy(traj[f"feat{t}"])
        depth = torch.from_numpy(traj[f"depth{t}"]).unsqueeze(-1)
        ego = orc.step(feat, depth, torch.from_numpy(traj[f"gps{t}"]), torch.from_numpy(traj[f"compass{t}"]),
                       torch.from_numpy(traj[f"masks{t}"]), keep=True)
        assert np.array_equal(orc.last["lin"].numpy().astype(np.int16), traj[f"lin{t}"])
        assert np.array_equal(np.packbits(orc.last["invalid"].numpy()), traj[f"invalid{t}"])
        assert np.array_equal(orc.last["proj"].numpy(), traj[f"proj{t}"])
        assert np.array_equal(ego.numpy(), traj[f"ego{t}"])
        assert _sha(orc.full_global_map.numpy()) == str(traj[f"mapsha{t}"])
    assert np.array_equal(orc.full_global_map.numpy(), traj["map_final"])


def test_trajectory_small_spec(traj):
    bs, c, steps = int(traj["bs"]), int(traj["c"]), int(traj["steps"])
    gmap = np.zeros((bs, 240, 240, c), np.float32)
    for t in range(steps):
        trig = dict(neg=(traj[f"cosneg{t}"], traj[f"sinneg{t}"]), pos=(traj[f"cospos{t}"], traj[f"sinpos{t}"]))
        ego, inter = spec_step(gmap, traj[f"feat{t}"], traj[f"depth{t}"], traj[f"gps{t}"], traj[f"compass{t}"],
                               traj[f"masks{t}"][:, 0], trig)
        assert np.array_equal(inter["lin"].astype(np.int16), traj[f"lin{t}"])
        assert np.array_equal(np.packbits(inter["invalid"]), traj[f"invalid{t}"])
        assert np.array_equal(inter["proj"], traj[f"proj{t}"])
        assert np.array_equal(ego, traj[f"ego{t}"])
        assert _sha(gmap) == str(traj[f"mapsha{t}"])


def test_frame_real_shapes(golden_dir):
    g = np.load(os.path.join(golden_dir, "frame_real.npz"))
    bs, c, hf, hd = int(g["bs"]), int(g["c"]), int(g["hf"]), int(g["hd"])
    gen = torch.Generator().manual_seed(int(g["seed"]))
    feat = make_features(bs, c, hf, hf, gen)
    depth = make_depth("uniform", bs, hd, hd, gen)
    if _sha(feat.numpy()) != str(g["feat_sha"]) or _sha(depth.numpy()) != str(g["depth_sha"]):
        pytest.skip("torch RNG stream differs from the one the golden file was made with")
    geo = MapGeometry()
    lin, invalid = spec_cells(depth[..., 0].numpy(), hf, hf, geo)
    assert np.array_equal(lin.astype(np.int16), g["lin"])
    assert np.array_equal(np.packbits(invalid), g["invalid"])
    proj, occ = spec_scatter(feat.numpy(), lin, invalid, geo)
    assert _sha(proj) == str(g["proj_sha"])
    assert np.array_equal(proj[0].argmax(0).astype(np.uint8), g["argmax"])       # ties -> lowest channel
    assert np.array_equal(np.packbits((proj[0] != 0).any(0)), g["occupied"])
    orc = OracleMapper(bs, c)
    ego = orc.step(feat, depth, torch.from_numpy(g["gps"]), torch.from_numpy(g["compass"]), torch.zeros(bs, 1))
    assert _sha(ego.numpy()) == str(g["ego_sha"])
    assert _sha(orc.full_global_map.numpy()) == str(g["map_sha"])
    assert np.array_equal(ego.numpy().reshape(-1)[g["ego_idx"]], g["ego_val"])
    assert np.array_equal(orc.full_global_map.numpy().reshape(-1)[g["map_idx"]], g["map_val"])
