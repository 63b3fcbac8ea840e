"""HARNESS ONLY -- BASELINE.json configs[4]: the reference's full policy forward (BasePolicy.act -> MGMapNet.forward:
instruction LSTM, UNet, depth ResNet, map update, map encoder/decoder, two GRUs, cross-modal attention) at batch 64,
once with the reference's own `vlnce_baselines.common.rgb_mapping` and once with the drop-in module swapped in at the
import path the reference uses (mg_map_policy.py:16).

The policy code is the UNMODIFIED reference (read from /root/reference or the verbatim copy under baseline/_ref/);
third-party packages this image lacks are stand-ins (baseline/habitat_shims.py); weights are random (no network for
checkpoints), identical for both variants (same seed, same construction order; the mapping module owns no parameters).
Used by tests/test_policy_dropin.py and bench.py's `policy_forward` record.  Not on the product path.
"""
from __future__ import annotations

import contextlib
import os
import sys
import tempfile
import types

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from baseline import habitat_shims  # noqa: E402

REFERENCE_ROOTS = ("/root/reference", os.path.join(ROOT, "baseline", "_ref"))
_PKGS = ("vlnce_baselines", "vlnce_baselines.models", "vlnce_baselines.models.encoders", "vlnce_baselines.common")
_REIMPORT = ("vlnce_baselines.models.mg_map_policy", "vlnce_baselines.models.policy",
             "vlnce_baselines.models.encoders.unet_encoder", "vlnce_baselines.models.encoders.map_encoder",
             "vlnce_baselines.models.encoders.resnet_encoders", "vlnce_baselines.models.encoders.instruction_encoder")


def reference_root():
    for r in REFERENCE_ROOTS:
        if os.path.isfile(os.path.join(r, "vlnce_baselines", "models", "mg_map_policy.py")):
            return r
    return None


def available() -> bool:
    return reference_root() is not None


def model_config(num_proc: int, unet_ckpt: str, gpu_id: int = 0):
    """The MODEL node of the reference's config (config/default.py:81-137) with the file-backed pieces switched off
    (pretrained word embeddings, DD-PPO depth checkpoint) and RGBMAPPING as default.py:188-189 fills it."""
    return habitat_shims.Config.of({
        "INSTRUCTION_ENCODER": dict(vocab_size=2504, max_length=200, use_pretrained_embeddings=False, embedding_file="",
                                    dataset_vocab="", fine_tune_embeddings=False, embedding_size=50, hidden_size=128,
                                    rnn_type="LSTM", final_state_only=False, bidirectional=True, backbone="lstm"),
        "RGB_ENCODER": dict(output_size=256, backbone="unet", pretrain_model=unet_ckpt),
        "DEPTH_ENCODER": dict(output_size=128, backbone="resnet50", ddppo_checkpoint="NONE"),
        "MAP_ENCODER": dict(ego_map_size=100, output_size=256),
        "STATE_ENCODER": dict(hidden_size=512, rnn_type="GRU", input_type=["rgb", "depth", "map"]),
        "PROGRESS_MONITOR": dict(use=True, alpha=1.0),
        "CONTRASTIVE_MONITOR": dict(target_tau=0.07, use=True, alpha=1.0),
        "PREDICTION_MONITOR": dict(use=True, alpha=1.0),
        "RGBMAPPING": dict(map_depth=64, global_map_size=240, egocentric_map_size=100, resolution=0.12, gpu_id=gpu_id,
                           num_proc=num_proc),
    })


@contextlib.contextmanager
def _patched():
    """Two compatibility patches for running 2020-era code on this image's torch / torchvision, both outside the map
    path: resnet18(pretrained=True) would download weights (no network) -> random init; pack_padded_sequence wants its
    lengths on the CPU since torch 1.7 (instruction_encoder.py:83-86 passes a CUDA tensor)."""
    import torchvision.models as tvm
    orig_r18, orig_pack = tvm.resnet18, nn.utils.rnn.pack_padded_sequence
    tvm.resnet18 = lambda pretrained=False, **kw: orig_r18(weights=None)
    nn.utils.rnn.pack_padded_sequence = lambda x, lengths, **kw: orig_pack(
        x, lengths.cpu() if isinstance(lengths, torch.Tensor) else lengths, **kw)
    try:
        yield
    finally:
        tvm.resnet18, nn.utils.rnn.pack_padded_sequence = orig_r18, orig_pack


def _bind_packages(which: str):
    """Make `vlnce_baselines.*` importable from the reference tree WITHOUT executing vlnce_baselines/__init__.py (it
    imports the trainers, hence Habitat), and bind `vlnce_baselines.common.rgb_mapping` to the reference file or to
    the drop-in -- the one line a maintainer changes."""
    root = reference_root()
    if root is None:
        raise FileNotFoundError("reference policy files not found (run __graft_entry__.build() in the build container)")
    habitat_shims.install()
    for name in _PKGS:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(root, *name.split("."))]
        sys.modules[name] = m
    for name in _REIMPORT:
        sys.modules.pop(name, None)
    if which == "reference":
        from oracle.reference_loader import load_reference_module
        mapping_mod = load_reference_module()
    elif which == "dropin":
        import wsmgmap_b200  # noqa: F401
        from wsmgmap_b200 import rgb_mapping as mapping_mod
    else:
        raise ValueError(which)
    sys.modules["vlnce_baselines.common.rgb_mapping"] = mapping_mod
    return mapping_mod


def _unet_checkpoint(path: str, seed: int):
    """A checkpoint file of the shape UNet.__init__ loads (unet_encoder.py:19-22): random ResNetUNet(3, 27) weights."""
    import importlib
    ue = importlib.import_module("vlnce_baselines.models.encoders.unet_encoder")
    torch.manual_seed(seed)
    net = ue.ResNetUNet(3, 27)
    state = {"models": {"img_segm_model": {"module.segm_model." + k: v for k, v in net.state_dict().items()}}}
    torch.save(state, path)


def build_policy(which: str, num_proc: int, device, seed: int = 0):
    """which = "reference" | "dropin".  Returns the reference's BasePolicy on `device`, in eval mode."""
    import importlib
    dev = torch.device(device)
    mapping_mod = _bind_packages(which)
    with _patched():
        tmp = tempfile.mkdtemp(prefix="wsmg_unet_")
        ckpt = os.path.join(tmp, "unet.pt")
        _unet_checkpoint(ckpt, seed)
        pol_mod = importlib.import_module("vlnce_baselines.models.policy")
        assert pol_mod.MGMapNet.__module__ == "vlnce_baselines.models.mg_map_policy"
        obs_space = habitat_shims.Dict({"depth": habitat_shims.Box(0.0, 1.0, (256, 256, 1)),
                                        "rgb": habitat_shims.Box(0, 255, (224, 224, 3))})
        act_space = habitat_shims.Box(-1.0, 1.0, (2,))
        torch.manual_seed(seed + 1)
        cfg = model_config(num_proc, ckpt, dev.index or 0)
        if dev.type == "cuda":
            with torch.cuda.device(dev):
                policy = pol_mod.BasePolicy(obs_space, act_space, cfg)
        else:
            # CPU smoke runs of the harness (reference mapping only): rgb_mapping.py:14 hard-codes torch.device("cuda", id)
            real_device = torch.device
            torch.device = lambda *a, **k: real_device("cpu")
            try:
                policy = pol_mod.BasePolicy(obs_space, act_space, cfg)
            finally:
                torch.device = real_device
        os.remove(ckpt)
        os.rmdir(tmp)
    policy = policy.to(dev).eval()
    assert type(policy.net.rgb_mapping_module).__module__ == mapping_mod.__name__
    return policy


def make_observations(batch: int, steps: int, seed: int, device):
    """`steps` observation dicts of the shapes batch_obs hands the policy (instruction tokens, rgb 224, depth 256, gps,
    compass) and the masks (0 on the first step)."""
    import wsmgmap_b200  # noqa: F401
    from wsmgmap_b200.synth import DEPTH_KINDS, RandomWalk, make_depth
    gen = torch.Generator().manual_seed(seed)
    tokens = torch.randint(1, 2504, (batch, 200), generator=gen)
    lengths = torch.randint(8, 80, (batch,), generator=gen)
    tokens[torch.arange(200)[None, :] >= lengths[:, None]] = 0
    walk = RandomWalk(batch, seed=seed + 1)
    out = []
    for t in range(steps):
        gps, compass, masks = walk.step()
        depth = torch.cat([make_depth(DEPTH_KINDS[(b + t) % 4], 1, 256, 256, gen) for b in range(batch)], 0)
        rgb = torch.rand(batch, 224, 224, 3, generator=gen) * 255.0
        obs = dict(instruction=tokens.to(device), rgb=rgb.to(device), depth=depth.to(device), gps=gps.to(device),
                   compass=compass.to(device))
        out.append((obs, masks.to(device)))
    return out


class MapTimer:
    """CUDA events around every call of the mapping module (forward pre-hook / hook, the mechanism the reference's
    own trainer uses to observe the module, dagger_trainer.py:322-327)."""

    def __init__(self, module):
        self.pairs = []
        self._h = [module.register_forward_pre_hook(self._pre), module.register_forward_hook(self._post)]

    def _pre(self, mod, inp):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.pairs.append([e, None])

    def _post(self, mod, inp, out):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.pairs[-1][1] = e

    def total_ms(self):
        return sum(a.elapsed_time(b) for a, b in self.pairs)

    def close(self):
        for h in self._h:
            h.remove()


@torch.no_grad()
def rollout(policy, frames, time_it: bool = False):
    """Step the policy over `frames` (BasePolicy.act, deterministic).  Returns per step (action, value, ego map) on the
    CPU, plus (total ms, map-update ms) when time_it."""
    dev = next(policy.parameters()).device
    bs = frames[0][1].shape[0]
    hidden = torch.zeros(policy.net.num_recurrent_layers, bs, policy.net.output_size, device=dev)
    prev = torch.zeros(bs, 2, device=dev)
    timer = MapTimer(policy.net.rgb_mapping_module) if time_it else None
    outs = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with _patched():                                  # (instruction_encoder.py:83-86 packs with CUDA lengths)
        for obs, masks in frames:
            obs = dict(obs)                           # the module adds 'rgb_ego_map' to the dict it is given
            value, action, _, hidden = policy.act(obs, hidden, prev, masks, deterministic=True)
            prev = action
            outs.append((action, value, obs["rgb_ego_map"]))
    e1.record()
    torch.cuda.synchronize(dev)
    res = [(a.cpu(), v.cpu(), m.cpu()) for a, v, m in outs]
    if not time_it:
        return res
    total, share = e0.elapsed_time(e1), timer.total_ms()
    timer.close()
    return res, total, share
