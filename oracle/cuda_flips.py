"""TEST / BASELINE INFRASTRUCTURE -- the kernels against the reference's PyTorch ops run ON CUDA (what users of the
reference actually execute).  Used by tests/test_cuda_reference.py and by bench.py's `cuda_reference_flips` record; never by
the product path."""
import torch

from oracle.mapping_oracle import OracleMapper


def flips_against_cuda_reference(n_envs=8, seed=0):
    """Returns {kind: (flipped pixels, pixels, max |ego difference|, max |ego|)} of the kernels against the reference's
    ops run on the GPU (oracle port, device=cuda; the unmodified reference file when it is staged)."""
    import wsmgmap_b200  # noqa: F401
    from wsmgmap_b200 import ops
    from wsmgmap_b200.synth import DEPTH_KINDS, make_depth, make_features
    DEV = torch.device("cuda", 0)
    c, hf, hd = 64, 224, 256
    gen = torch.Generator().manual_seed(seed)
    out = {}
    for kind in DEPTH_KINDS:
        depth = make_depth(kind, n_envs, hd, hd, gen).to(DEV)
        feat = make_features(n_envs, c, hf, hf, gen).to(DEV)
        gps = torch.randn(n_envs, 2, generator=gen).to(DEV)
        compass = (torch.rand(n_envs, 1, generator=gen) * 6 - 3).to(DEV)
        masks = torch.zeros(n_envs, 1, device=DEV)
        orc = OracleMapper(n_envs, c, device=DEV)
        want = orc.step(feat, depth, gps, compass, masks, keep=True)
        lin, inv = ops.unproject_index(depth, hf, hf)
        flipped = int(((lin.long() != orc.last["lin"]) | (inv != orc.last["invalid"])).sum())
        gmap = torch.zeros(n_envs, 240, 240, c, device=DEV)
        ego = ops.map_update(feat, depth, gps, compass, masks, gmap)
        torch.cuda.synchronize()
        out[kind] = (flipped, lin.numel(), float((ego - want).abs().max()), float(want.abs().max()))
    return out
