"""Drop-in contract of the RGBMapping module: every interaction the reference's
policy and trainers have with the module is replayed here against the oracle."""
import types

import pytest
import torch

from oracle.mapping_oracle import OracleMapper
import wsmgmap_b200  # noqa: F401
from wsmgmap_b200.rgb_mapping import RGBMapping
from wsmgmap_b200.synth import make_depth, make_features

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _cfg(num_proc, c=8):
    return types.SimpleNamespace(gpu_id=0, num_proc=num_proc, resolution=0.12, egocentric_map_size=100,
                                 global_map_size=240, map_depth=c)


def _close(a, b, scale):
    return ((a - b).abs() <= 1e-5 * b.abs() + 2e-5 * scale).all()


def test_module_surface_and_state():
    m = RGBMapping(_cfg(4))
    assert isinstance(m, torch.nn.Module)
    assert len(list(m.parameters())) == 0 and len(list(m.buffers())) == 0 and len(m.state_dict()) == 0
    assert tuple(m.full_global_map.shape) == (4, 240, 240, 8) and m.full_global_map.device == DEV
    assert tuple(m.agent_view.shape) == (4, 8, 240, 240)


def test_forward_hook_rebinding_pause_and_cache():
    c, hf, hd, n = 8, 64, 64, 4
    m = RGBMapping(_cfg(n, c))
    orc = OracleMapper(n, c)
    seen = []
    m.register_forward_hook(lambda mod, i, o: seen.append(o.cpu()))           # dagger_trainer.py:303-306,325-327
    gen = torch.Generator().manual_seed(1)

    def frame(bs):
        return (make_features(bs, c, hf, hf, gen), make_depth("near", bs, hd, hd, gen),
                torch.randn(bs, 2, generator=gen), torch.rand(bs, 1, generator=gen) * 6 - 3)

    # trainers re-bind the state before a rollout (common_trainer.py:266-267, dagger_trainer.py:669-678)
    m.full_global_map = torch.zeros([n] + list(m.full_global_map.shape[1:]), device=DEV)
    m.agent_view = torch.zeros([n] + list(m.agent_view.shape[1:]), device=DEV)
    feat, depth, gps, compass = frame(n)
    obs = dict(depth=depth.to(DEV), gps=gps.to(DEV), compass=compass.to(DEV))
    masks = torch.zeros(n, 1)
    out = m(feat.to(DEV), obs, masks.to(DEV))
    want = orc.step(feat, depth, gps, compass, masks)
    assert obs["rgb_ego_map"] is out and tuple(out.shape) == (n, c, 100, 100)
    assert _close(out.cpu(), want, 1.0) and _close(m.full_global_map.cpu(), orc.full_global_map, 1.0)
    assert len(seen) == 1 and torch.equal(seen[0], out.cpu())
    # cached path (rgb_mapping.py:80,87-88): features may be None
    assert m(None, obs, masks.to(DEV)) is out
    # _pause_envs: fancy-index the state to fewer envs and assign it back (common_trainer.py:171-172,476)
    keep = [0, 2, 3]
    m.full_global_map = m.full_global_map[keep]
    orc.full_global_map = orc.full_global_map[keep]
    feat, depth, gps, compass = frame(3)
    obs = dict(depth=depth.to(DEV), gps=gps.to(DEV), compass=compass.to(DEV))
    masks = torch.ones(3, 1)
    out = m(feat.to(DEV), obs, masks.to(DEV))
    want = orc.step(feat, depth, gps, compass, masks)
    assert _close(out.cpu(), want, 1.0) and _close(m.full_global_map.cpu(), orc.full_global_map, 1.0)
    # batch smaller than the state (rgb_mapping.py:35,56 touch rows [:bs] only)
    before = m.full_global_map[2].clone()
    feat, depth, gps, compass = frame(2)
    obs = dict(depth=depth.to(DEV), gps=gps.to(DEV), compass=compass.to(DEV))
    out = m(feat.to(DEV), obs, torch.ones(2, 1, device=DEV))
    want = orc.step(feat, depth, gps, compass, torch.ones(2, 1))
    assert _close(out.cpu(), want, 1.0) and torch.equal(m.full_global_map[2], before)
    # secondary API returns (final_retrieval, the very same map tensor)
    ego, gm = m.project_feat_to_map(feat.to(DEV), m.full_global_map, obs, torch.ones(2, 1, device=DEV))
    assert gm is m.full_global_map and ego.shape == out.shape


def test_channel_rebinning_and_input_checks():
    m = RGBMapping(_cfg(2, 4))
    gen = torch.Generator().manual_seed(2)
    feat = make_features(2, 8, 32, 32, gen)                      # 8 channels pooled to map_depth 4 (rgb_mapping.py:82-84)
    depth = make_depth("near", 2, 32, 32, gen)
    obs = dict(depth=depth.to(DEV), gps=torch.zeros(2, 2, device=DEV), compass=torch.zeros(2, 1, device=DEV))
    out = m(feat.to(DEV), obs, torch.zeros(2, 1, device=DEV))
    orc = OracleMapper(2, 4)
    pooled = torch.nn.functional.adaptive_max_pool1d(feat.permute(0, 2, 3, 1).reshape(2, -1, 8), 4)
    pooled = pooled.reshape(2, 32, 32, 4).permute(0, 3, 1, 2)
    want = orc.step(pooled, depth, torch.zeros(2, 2), torch.zeros(2, 1), torch.zeros(2, 1))
    assert _close(out.cpu(), want, 1.0)
    with pytest.raises(ValueError):
        m(feat.to(DEV), dict(depth=depth, gps=obs["gps"], compass=obs["compass"]), torch.zeros(2, 1, device=DEV))  # CPU depth
    # a wider producer at the reference's shapes (96 -> 64 channels, uneven bins of 1 and 2): the compile-time-geometry build
    m2 = RGBMapping(_cfg(2, 64))
    feat2 = make_features(2, 96, 224, 224, gen, signed=True)
    depth2 = make_depth("room2", 2, 256, 256, gen)
    gps2, comp2 = torch.randn(2, 2, generator=gen), torch.rand(2, 1, generator=gen) * 6 - 3
    obs2 = dict(depth=depth2.to(DEV), gps=gps2.to(DEV), compass=comp2.to(DEV))
    out2 = m2(feat2.to(DEV), obs2, torch.zeros(2, 1, device=DEV))
    pooled2 = torch.nn.functional.adaptive_max_pool1d(feat2.permute(0, 2, 3, 1).reshape(2, -1, 96), 64)
    pooled2 = pooled2.reshape(2, 224, 224, 64).permute(0, 3, 1, 2).contiguous()
    orc2 = OracleMapper(2, 64)
    want2 = orc2.step(pooled2, depth2, gps2, comp2, torch.zeros(2, 1))
    assert _close(out2.cpu(), want2, float(feat2.abs().max())) and _close(m2.full_global_map.cpu(), orc2.full_global_map, float(feat2.abs().max()))


def test_channels_last_producer_is_consumed_as_is():
    """SURVEY 8f rank 1 (unet_encoder.py:103-111): features in torch.channels_last memory format go to the kernel without a
    permute copy (wsmg_dims.feat_nhwc) and give the same bits as the NCHW path; the oracle confirms the values.  Real
    shapes (compile-time geometry build), two steps, plus the host-buffer pipeline with an NHWC pinned tensor."""
    from wsmgmap_b200 import ops
    c, hf, hd, n = 64, 224, 256, 3
    gen = torch.Generator().manual_seed(12)
    a, b = RGBMapping(_cfg(n, c)), RGBMapping(_cfg(n, c))
    orc = OracleMapper(n, c)
    for t in range(2):
        feat = make_features(n, c, hf, hf, gen, signed=(t == 1))
        depth = make_depth(("room4", "uniform")[t], n, hd, hd, gen)
        gps, compass = torch.randn(n, 2, generator=gen), torch.rand(n, 1, generator=gen) * 6 - 3
        masks = torch.full((n, 1), float(t))
        obs = lambda: dict(depth=depth.to(DEV), gps=gps.to(DEV), compass=compass.to(DEV))  # noqa: E731
        f_cl = feat.to(DEV).contiguous(memory_format=torch.channels_last)
        assert ops.is_channels_last(f_cl) and f_cl.shape == feat.shape
        oa = a(feat.to(DEV), obs(), masks.to(DEV))
        ob = b(f_cl, obs(), masks.to(DEV))
        want = orc.step(feat, depth, gps, compass, masks)
        assert torch.equal(oa, ob) and torch.equal(a.full_global_map, b.full_global_map)
        assert _close(ob.cpu(), want, float(want.abs().max())) and _close(b.full_global_map.cpu(), orc.full_global_map, float(want.abs().max()))
    # host-buffer entry with an NHWC pinned tensor, dead rows skipped: the span of live rows is one contiguous copy
    d = ops.dims_for(feat.shape, depth.shape, n, feat_nhwc=True)
    pipe = ops.HostPipeline(d, DEV, chunk_envs=2, skip_dead_rows=True)
    pipe.staging.fill_(0xFF)
    g3 = torch.zeros(n, 240, 240, c, device=DEV)
    ego_h = torch.empty(n, c, 100, 100).pin_memory()
    f_h = feat.permute(0, 2, 3, 1).contiguous().pin_memory()             # [n,Hf,Wf,C] in memory
    pipe.step(f_h, depth.pin_memory(), gps.pin_memory(), compass.pin_memory(), torch.zeros(n, 1).pin_memory(), g3, ego_h)
    g4 = torch.zeros(n, 240, 240, c, device=DEV)
    ego4 = ops.map_update(feat.to(DEV), depth.to(DEV), gps.to(DEV), compass.to(DEV), torch.zeros(n, 1, device=DEV), g4)
    torch.cuda.synchronize()
    assert torch.equal(ego_h, ego4.cpu()) and torch.equal(g3, g4)


def test_optin_extras_env_slots_and_half_store():
    """slot table instead of map[state_index]; fp16 ego map for the rollout store.  Both are
    checked against the ORACLE stepped the reference's way (re-indexed state), not only against each other."""
    import numpy as np
    c, hf, hd, n = 8, 64, 64, 4
    gen = torch.Generator().manual_seed(3)

    def frame(bs):
        return (make_features(bs, c, hf, hf, gen), make_depth("room2", bs, hd, hd, gen),
                torch.randn(bs, 2, generator=gen), torch.rand(bs, 1, generator=gen) * 6 - 3)

    a, b = RGBMapping(_cfg(n, c)), RGBMapping(_cfg(n, c))        # a: reference-style re-indexing, b: slot table
    orc = OracleMapper(n, c)
    b.store_half = True
    feat, depth, gps, compass = frame(n)
    obs = lambda: dict(depth=depth.to(DEV), gps=gps.to(DEV), compass=compass.to(DEV))  # noqa: E731
    oa = a(feat.to(DEV), obs(), torch.zeros(n, 1, device=DEV))
    ob = b(feat.to(DEV), obs(), torch.zeros(n, 1, device=DEV))
    want = orc.step(feat, depth, gps, compass, torch.zeros(n, 1))
    assert torch.equal(oa, ob)
    assert _close(ob.cpu(), want, 1.0) and _close(b.full_global_map.cpu(), orc.full_global_map, 1.0)
    host = b.ego_half_to_host()
    torch.cuda.synchronize()
    assert np.array_equal(host.numpy(), oa.cpu().numpy().astype(np.float16))     # common_trainer.py:519-520
    assert np.allclose(host.numpy().astype(np.float32), want.numpy(), rtol=1e-3, atol=1e-3)   # fp16 of the oracle's map
    keep = [0, 1, 3]
    a.full_global_map = a.full_global_map[keep]                                  # what _pause_envs does
    orc.full_global_map = orc.full_global_map[keep]
    b.pause_envs([2])
    assert b.env_slots.tolist() == keep and b.full_global_map.shape[0] == n
    paused_row = b.full_global_map[2].clone()
    feat, depth, gps, compass = frame(3)
    oa = a(feat.to(DEV), obs(), torch.ones(3, 1, device=DEV))
    ob = b(feat.to(DEV), obs(), torch.ones(3, 1, device=DEV))
    want = orc.step(feat, depth, gps, compass, torch.ones(3, 1))
    assert torch.equal(oa, ob)
    assert torch.equal(a.full_global_map, b.full_global_map[keep])
    assert _close(ob.cpu(), want, 1.0) and _close(b.full_global_map[keep].cpu(), orc.full_global_map, 1.0)
    assert torch.equal(b.full_global_map[2], paused_row)                         # the paused env's map is untouched
    b.pause_envs([0])                                                            # slots compose: [1, 3]
    assert b.env_slots.tolist() == [1, 3]
    feat, depth, gps, compass = frame(2)
    ob = b(feat.to(DEV), obs(), torch.tensor([[1.0], [0.0]], device=DEV))        # second frame starts a new episode
    orc.full_global_map = orc.full_global_map[[1, 2]]
    want = orc.step(feat, depth, gps, compass, torch.tensor([[1.0], [0.0]]))
    assert _close(ob.cpu(), want, 1.0) and _close(b.full_global_map[[1, 3]].cpu(), orc.full_global_map, 1.0)


def test_env_slot_validation_and_rebinding():
    """a slot table must address distinct, existing rows of the map it was made for; re-binding the map
    with another leading dimension (what the unmodified trainer does) drops it instead of leaving stale rows."""
    import ctypes
    import warnings
    from wsmgmap_b200 import _lib, ops
    c, hf, hd, n = 4, 32, 32, 4
    m = RGBMapping(_cfg(n, c))
    with pytest.raises(ValueError, match="rows of full_global_map"):
        m.env_slots = [0, 4]
    with pytest.raises(ValueError, match="distinct"):
        m.env_slots = torch.tensor([1, 1], dtype=torch.int32)
    m.pause_envs([1])
    assert m.env_slots.tolist() == [0, 2, 3] and m.env_slots.dtype == torch.int32 and m.env_slots.device == DEV
    gen = torch.Generator().manual_seed(9)
    feat, depth = make_features(2, c, hf, hf, gen).to(DEV), make_depth("near", 2, hd, hd, gen).to(DEV)
    obs = dict(depth=depth, gps=torch.zeros(2, 2, device=DEV), compass=torch.zeros(2, 1, device=DEV))
    with pytest.raises(ValueError, match="entries for a batch"):
        m(feat, dict(obs), torch.ones(2, 1, device=DEV))                         # 3 slots, batch of 2
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m.full_global_map = m.full_global_map[[0, 2, 3]]                         # the trainer re-indexes as well: table dropped
        assert any("slot table" in str(x.message) for x in w)
    assert m.env_slots is None
    # the C ABI on its own: a slot outside the map skips that frame, raises the flag and the status word
    lib = _lib.load()
    gmap = torch.zeros(2, 240, 240, c, device=DEV)
    ego = torch.full((2, c, 100, 100), -7.0, device=DEV)
    d = ops.dims_for(feat.shape, depth.shape, 2)
    scratch = ops.alloc_scratch(d, DEV)
    slots = torch.tensor([1, 5], dtype=torch.int32, device=DEV)
    status = torch.zeros(2, dtype=torch.int32).pin_memory()
    P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    opts = _lib.WsmgOpts(None, None, P(slots), None, None, P(status))
    z = torch.zeros(2, 2, device=DEV)
    rc = lib.wsmg_map_update_ex(P(feat), P(depth), P(z), P(z), P(z), P(gmap), P(ego), ctypes.byref(opts), P(scratch),
                                scratch.numel(), ctypes.byref(d), None)
    torch.cuda.synchronize()
    assert rc == 0 and status[1].item() == 1
    flags = ops.env_flags(scratch, d)
    assert (flags[1] & _lib.FLAG_BAD_SLOT) and not (flags[0] & _lib.FLAG_BAD_SLOT)
    assert (ego[1] == -7.0).all() and not (ego[0] == -7.0).all()                 # frame 1 skipped, frame 0 done (on row 1)
    assert gmap[1].abs().sum() > 0 and gmap[0].abs().sum() == 0


def test_negative_depth_warns_lazily_without_sync():
    """Default path (strict_inputs off): pixels behind the camera are dropped -- the one documented divergence from
    the reference -- and the module says so at its next call, from a pinned status word, without synchronising."""
    import warnings
    m = RGBMapping(_cfg(2, 4))
    gen = torch.Generator().manual_seed(4)
    feat = make_features(2, 4, 32, 32, gen).to(DEV)
    depth = make_depth("near", 2, 32, 32, gen)
    depth[1, :16, :, 0] = -0.2
    obs = dict(depth=depth.to(DEV), gps=torch.zeros(2, 2, device=DEV), compass=torch.zeros(2, 1, device=DEV))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m(feat, dict(obs), torch.zeros(2, 1, device=DEV))
        torch.cuda.synchronize()                                                 # (only so that the test is deterministic)
        assert not any("depth < 0" in str(x.message) for x in w)
        obs2 = dict(depth=depth.clamp_min(0).to(DEV), gps=obs["gps"], compass=obs["compass"])
        m(feat, obs2, torch.ones(2, 1, device=DEV))
        assert any("depth < 0" in str(x.message) for x in w)
        n = len(w)
        m(feat, dict(obs2), torch.ones(2, 1, device=DEV))                        # reported once, not again
        torch.cuda.synchronize()
        m(feat, dict(obs2), torch.ones(2, 1, device=DEV))
        assert len([x for x in w if "depth < 0" in str(x.message)]) == len([x for x in w[:n] if "depth < 0" in str(x.message)])


def test_strict_inputs_flags_negative_depth():
    m = RGBMapping(_cfg(2, 4))
    m.strict_inputs = True
    gen = torch.Generator().manual_seed(4)
    feat = make_features(2, 4, 32, 32, gen)
    depth = make_depth("near", 2, 32, 32, gen)
    obs = dict(depth=depth.to(DEV), gps=torch.zeros(2, 2, device=DEV), compass=torch.zeros(2, 1, device=DEV))
    m(feat.to(DEV), obs, torch.zeros(2, 1, device=DEV))                     # fine
    depth[1, :16, :, 0] = -0.2                                              # behind the camera (rows above the horizon pass the height test): not representable
    obs = dict(depth=depth.to(DEV), gps=torch.zeros(2, 2, device=DEV), compass=torch.zeros(2, 1, device=DEV))
    with pytest.raises(ValueError, match="envs \\[1\\]"):
        m(feat.to(DEV), obs, torch.ones(2, 1, device=DEV))
