"""Generate tests/golden/*.npz from the UNMODIFIED reference file.

Run in the build container only (needs /root/reference):
    python oracle/make_golden.py
The reference has no tests / fixtures of its own, so these
vectors are outputs of the reference module itself (loaded by
oracle/reference_loader.py) on seeded inputs.  They are what pins the oracle --
and through it the CUDA path -- on the GPU box, where /root/reference is absent.
"""
import hashlib
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle.reference_loader import make_reference_mapper  # noqa: E402
import wsmgmap_b200  # noqa: E402,F401
from wsmgmap_b200.synth import RandomWalk, make_depth, make_features  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_intermediates(mod, ref, feat, depth):
    """Re-run the reference's own stage objects to expose linear_locs / invalid / proj_feats
    (they are locals of ProjectToGroundPlane.forward, rgb_mapping.py:184-232)."""
    proj = ref.projection
    locs, valid = proj.compute_spatial_locs.forward(depth * 10)
    captured = {}
    import torch_scatter
    orig = torch_scatter.scatter_max

    def spy(src, index, dim=-1, out=None, dim_size=None):
        captured["lin"] = index[:, 0].clone()
        captured["src"] = src
        return orig(src, index, dim=dim, out=out, dim_size=dim_size)

    mod.torch_scatter.scatter_max = spy
    try:
        p = proj.project_to_ground_plane.forward(feat, locs, valid)
    finally:
        mod.torch_scatter.scatter_max = orig
    invalid = captured["src"][:, 0] == -1e16     # sentinel marks invalid writes (:212); no real feature equals it
    shape = feat.shape[0], feat.shape[2], feat.shape[3]
    return captured["lin"].reshape(shape), invalid.reshape(shape), p


def trajectory_small():
    bs, c, hf, hd, steps = 2, 4, 56, 64, 6
    ref, mod = make_reference_mapper(bs, map_depth=c)
    walk = RandomWalk(bs, seed=7, far_env=1)
    gen = torch.Generator().manual_seed(11)
    rec = dict(bs=bs, c=c, hf=hf, hd=hd, steps=steps)
    for t in range(steps):
        gps, compass, masks = walk.step()
        if t == 3:
            masks[0] = 0.0
        if t >= 2:
            gps[1] = gps[1] + torch.tensor([9.0, -10.5])     # beyond 8.4 m: window clips at the map border
        feat = make_features(bs, c, hf, hf, gen, signed=(t % 2 == 1))
        depth = make_depth(("uniform", "near", "room2")[t % 3], bs, hd, hd, gen)
        if t == 4:
            depth = (depth * 200).round() / 200                  # exact multiples of half a cell
        lin, invalid, proj = ref_intermediates(mod, ref, feat, depth)
        obs = dict(depth=depth.clone(), gps=gps.clone(), compass=compass.clone())
        ego = ref(feat.clone(), obs, masks.clone())
        rec[f"feat{t}"] = feat.numpy()
        rec[f"depth{t}"] = depth[..., 0].numpy()
        rec[f"gps{t}"] = gps.numpy()
        rec[f"compass{t}"] = compass.numpy()
        rec[f"masks{t}"] = masks.numpy()
        rec[f"cosneg{t}"] = torch.cos(-compass[:, 0]).numpy()
        rec[f"sinneg{t}"] = torch.sin(-compass[:, 0]).numpy()
        rec[f"cospos{t}"] = torch.cos(compass[:, 0]).numpy()
        rec[f"sinpos{t}"] = torch.sin(compass[:, 0]).numpy()
        rec[f"lin{t}"] = lin.numpy().astype(np.int16)
        rec[f"invalid{t}"] = np.packbits(invalid.numpy())
        rec[f"proj{t}"] = proj.numpy()
        rec[f"ego{t}"] = ego.numpy()
        rec[f"mapsha{t}"] = sha(ref.full_global_map.numpy())
    rec["map_final"] = ref.full_global_map.numpy()
    np.savez_compressed(os.path.join(OUT, "traj_small.npz"), **rec)


def frame_real():
    """One frame at the real shapes (C=64, 224x224 features, 256x256 depth); inputs are
    regenerated from the seed, outputs stored as hashes + samples + integer planes."""
    bs, c, hf, hd, seed = 1, 64, 224, 256, 2024
    ref, mod = make_reference_mapper(bs, map_depth=c)
    gen = torch.Generator().manual_seed(seed)
    feat = make_features(bs, c, hf, hf, gen)
    depth = make_depth("uniform", bs, hd, hd, gen)
    gps = torch.tensor([[1.37, -2.21]])
    compass = torch.tensor([[0.8123]])
    masks = torch.zeros(bs, 1)
    lin, invalid, proj = ref_intermediates(mod, ref, feat, depth)
    obs = dict(depth=depth.clone(), gps=gps, compass=compass)
    ego = ref(feat.clone(), obs, masks).numpy()
    gmap = ref.full_global_map.numpy()
    rs = np.random.default_rng(5)
    ei = rs.integers(0, ego.size, 8192)
    nz = np.flatnonzero(gmap.reshape(-1))
    mi = np.concatenate([rs.choice(nz, 6144), rs.integers(0, gmap.size, 2048)])
    p = proj.numpy()[0]
    np.savez_compressed(
        os.path.join(OUT, "frame_real.npz"),
        seed=seed, bs=bs, c=c, hf=hf, hd=hd, gps=gps.numpy(), compass=compass.numpy(),
        feat_sha=sha(feat.numpy()), depth_sha=sha(depth.numpy()),
        cosneg=torch.cos(-compass[:, 0]).numpy(), sinneg=torch.sin(-compass[:, 0]).numpy(),
        cospos=torch.cos(compass[:, 0]).numpy(), sinpos=torch.sin(compass[:, 0]).numpy(),
        lin=lin.numpy().astype(np.int16), invalid=np.packbits(invalid.numpy()),
        argmax=p.argmax(0).astype(np.uint8), occupied=np.packbits((p != 0).any(0)),
        proj_sha=sha(proj.numpy()), ego_sha=sha(ego), map_sha=sha(gmap),
        ego_idx=ei, ego_val=ego.reshape(-1)[ei], map_idx=mi, map_val=gmap.reshape(-1)[mi],
    )


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    trajectory_small()
    frame_real()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
