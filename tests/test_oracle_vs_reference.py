"""Pins the oracle against the reference FILE executed in this container
(skipped on the GPU box, where /root/reference does not exist -- there the
golden vectors in tests/golden/ take over, see test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

from oracle.mapping_oracle import OracleMapper, spec_step
from oracle.reference_loader import make_reference_mapper, reference_available
import wsmgmap_b200  # noqa: F401
from wsmgmap_b200.synth import RandomWalk, make_depth, make_features

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


def _trig(compass):
    c = compass[:, 0]
    return dict(neg=(torch.cos(-c).numpy(), torch.sin(-c).numpy()), pos=(torch.cos(c).numpy(), torch.sin(c).numpy()))


@pytest.mark.parametrize("c,hf,hd", [(8, 224, 256), (5, 64, 64), (4, 56, 64)])
def test_trajectory_bit_identical(c, hf, hd):
    bs, steps = 3, 5
    ref, _ = make_reference_mapper(bs, map_depth=c)
    orc = OracleMapper(bs, c)
    spec_map = np.zeros((bs, 240, 240, c), np.float32)
    walk = RandomWalk(bs, seed=3, reset_prob=0.2, far_env=2)
    gen = torch.Generator().manual_seed(c * 1000 + hf)
    for t in range(steps):
        gps, compass, masks = walk.step()
        if t >= 2:
            gps[2] += torch.tensor([-9.5, 11.0])
        feat = make_features(bs, c, hf, hf, gen, signed=(t % 2 == 0))
        depth = make_depth(("uniform", "room2", "near", "room4")[t % 4], bs, hd, hd, gen)
        obs = dict(depth=depth.clone(), gps=gps.clone(), compass=compass.clone())
        want = ref(feat.clone(), obs, masks.clone())
        got = orc.step(feat, depth, gps, compass, masks, keep=True)
        assert torch.equal(want, got)
        assert torch.equal(ref.full_global_map, orc.full_global_map)
        ego, inter = spec_step(spec_map, feat.numpy(), depth[..., 0].numpy(), gps.numpy(), compass.numpy(),
                               masks[:, 0].numpy(), _trig(compass))
        assert np.array_equal(inter["lin"], orc.last["lin"].numpy())
        assert np.array_equal(inter["invalid"], orc.last["invalid"].numpy())
        assert np.array_equal(inter["proj"], orc.last["proj"].numpy())
        assert np.array_equal(ego, want.numpy())
        assert np.array_equal(spec_map, ref.full_global_map.numpy())


def test_forward_short_circuit_and_state_rebinding():
    """rgb_mapping.py:80,87-88 (cached ego map) and the trainers' re-binding of
    full_global_map to a smaller tensor (common_trainer.py:171-172)."""
    ref, _ = make_reference_mapper(4, map_depth=4)
    gen = torch.Generator().manual_seed(0)
    feat = make_features(2, 4, 32, 32, gen)
    obs = dict(depth=make_depth("near", 2, 32, 32, gen), gps=torch.zeros(2, 2), compass=torch.zeros(2, 1))
    out = ref(feat, obs, torch.zeros(2, 1))
    assert obs["rgb_ego_map"] is out
    assert ref(None, obs, torch.zeros(2, 1)) is out
    ref.full_global_map = ref.full_global_map[[0, 1]]
    obs2 = dict(depth=obs["depth"], gps=obs["gps"], compass=obs["compass"])
    ref(feat, obs2, torch.ones(2, 1))
    assert ref.full_global_map.shape[0] == 2
