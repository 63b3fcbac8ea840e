"""BASELINE.json configs[4]: the reference's full policy forward (BasePolicy.act -> MGMapNet.forward, unmodified files
from /root/reference or baseline/_ref) with the map module swapped at the import path the reference uses
(mg_map_policy.py:16).  Third-party packages this image lacks are stand-ins (baseline/habitat_shims.py), weights are
random and identical for both variants."""
import pytest
import torch

from baseline import policy_harness as ph

needs_ref = pytest.mark.skipif(not ph.available(), reason="reference policy files not staged (baseline/_ref)")


@needs_ref
def test_harness_runs_reference_policy_on_cpu():
    """The harness itself (stand-ins, package binding, checkpoint shim) with the reference's own mapping module, CPU,
    two envs, two steps: the forward runs and the ego map it leaves in the observations is the oracle's."""
    from oracle.mapping_oracle import OracleMapper
    policy = ph.build_policy("reference", 2, "cpu", seed=3)
    assert type(policy.net.rgb_mapping_module).__name__ == "RGBMapping"
    frames = ph.make_observations(2, 2, seed=5, device="cpu")
    hidden = torch.zeros(policy.net.num_recurrent_layers, 2, 512)
    prev = torch.zeros(2, 2)
    orc = OracleMapper(2, 64)
    with torch.no_grad():
        for obs, masks in frames:
            obs = dict(obs)
            _, proj = policy.net.rgb_encoder(obs)
            value, action, _, hidden = policy.act(obs, hidden, prev, masks, deterministic=True)
            want = orc.step(proj, obs["depth"], obs["gps"], obs["compass"], masks)
            assert torch.equal(obs["rgb_ego_map"], want)
            assert action.shape == (2, 2) and value.shape == (2, 1) and torch.isfinite(action).all()
            prev = action


@needs_ref
@pytest.mark.gpu
def test_policy_forward_batch64_dropin_matches_reference():
    """Batch 64, three steps (reset, then two accumulating steps) of the reference's policy, once with its own mapping
    module (its torch ops ON CUDA) and once with the drop-in bound at the same import path.
      * strict: the ego map the drop-in leaves in observations['rgb_ego_map'] against the CPU oracle stepped on the very
        features the policy's UNet handed the module -- north_star's bar (1e-5 relative, + the atol the device sin / cos needs);
      * against the reference-on-CUDA run the same bar holds except around the odd pixel that the reference's own CUDA
        arithmetic puts into the neighbouring cell (tests/test_cuda_reference.py: ~2e-6 of the pixels), so there the
        share of elements outside the bar is bounded instead, and actions / values agree to float noise."""
    from oracle.mapping_oracle import OracleMapper
    dev = torch.device("cuda", 0)
    bs, steps = 64, 3
    frames = ph.make_observations(bs, steps, seed=11, device=dev)
    ref = ph.build_policy("reference", bs, dev, seed=7)
    out_ref = ph.rollout(ref, frames)
    ref_map = ref.net.rgb_mapping_module.full_global_map.cpu()
    del ref
    torch.cuda.empty_cache()
    new = ph.build_policy("dropin", bs, dev, seed=7)
    mod = new.net.rgb_mapping_module
    assert type(mod).__module__.startswith("wsmgmap_b200")
    fed = []
    mod.register_forward_pre_hook(lambda m, inp: fed.append((inp[0].cpu(), {k: inp[1][k].cpu() for k in ("depth", "gps", "compass")}, inp[2].cpu())))
    out_new = ph.rollout(new, frames)
    new_map = mod.full_global_map.cpu()

    def outside(a, b):
        scale = float(b.abs().max())
        return ((a - b).abs() > 1e-5 * b.abs() + 2e-5 * scale).float().mean().item()

    orc = OracleMapper(bs, 64)
    for (feat, obs, masks), (_, _, m_n) in zip(fed, out_new):
        want = orc.step(feat, obs["depth"], obs["gps"], obs["compass"], masks)
        assert outside(m_n, want) == 0.0
    assert outside(new_map, orc.full_global_map) == 0.0
    for (a_r, v_r, m_r), (a_n, v_n, m_n) in zip(out_ref, out_new):
        assert outside(m_n, m_r) < 1e-3
        assert torch.allclose(a_n, a_r, rtol=1e-2, atol=5e-3) and torch.allclose(v_n, v_r, rtol=1e-2, atol=5e-3)
    assert outside(new_map, ref_map) < 1e-3
