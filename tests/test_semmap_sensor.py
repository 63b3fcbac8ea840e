"""Ground-truth semantic map sensor registration (reference
habitat_extensions/sensors.py:383-410): oracle pinned against the reference's own method body (build
container only), the elementwise spec against the oracle, and -- GPU -- the CUDA kernel against both."""
import numpy as np
import pytest
import torch

from oracle.reference_loader import ReferenceSemMapSensor, sensor_reference_available
from oracle.semmap_oracle import sensor_crop, sensor_pose, spec_sensor_crop


def _random_semmap(rng, s=480, classes=27):
    """Blocky class-id map (rooms of constant label) with an empty margin, like the reference's .npy maps."""
    m = np.zeros((s, s), np.int64)
    for _ in range(160):
        y, x = rng.integers(60, s - 60, 2)
        h, w = rng.integers(4, 70, 2)
        m[y:y + h, x:x + w] = rng.integers(1, classes + 1)
    return m


def _episodes(rng, n):
    """(position deltas in metres, headings) of n observations of one episode."""
    pos = np.cumsum(rng.normal(0, 0.6, size=(n, 3)), 0)
    pos[0] = 0
    head = rng.uniform(-np.pi, np.pi, n)
    return pos, head


@pytest.mark.skipif(not sensor_reference_available(), reason="/root/reference not present")
def test_oracle_equals_reference_method():
    rng = np.random.default_rng(0)
    for ep in range(3):
        ref = ReferenceSemMapSensor(half_size=50)
        raw = _random_semmap(rng)
        pos, head = _episodes(rng, 6)
        start = rng.normal(0, 3.0, 3)
        for t in range(len(pos)):
            want = ref.observe(ep, start + pos[t], head[t], new_map=raw)
            # the reference pre-rotates the map once per episode (sensors.py:390-392) and keeps it
            held = ref.global_gt_semmap[0, 0]
            grid_y = (start[0] + pos[t][0] - ref.init_agent_state.position[0]) / 0.12 + 240      # sensors.py:395
            grid_x = (start[2] + pos[t][2] - ref.init_agent_state.position[2]) / 0.12 + 240      # sensors.py:396
            pose = sensor_pose([grid_y], [grid_x], [head[t]])
            got = sensor_crop(held[None], pose)[0]
            assert want.dtype == torch.int64 and tuple(want.shape) == (100, 100)
            assert torch.equal(want, got), (ep, t)


def test_spec_equals_oracle():
    rng = np.random.default_rng(1)
    for s, half, origin in ((480, 50, 289), (200, 30, 120), (96, 50, 60)):
        bs = 5
        maps = np.stack([_random_semmap(rng, s) if s >= 200 else rng.integers(0, 28, (s, s)) for _ in range(bs)]).astype(np.float32)
        gy = rng.uniform(s / 2 - 40, s / 2 + 40, bs)
        gx = rng.uniform(s / 2 - 40, s / 2 + 40, bs)
        head = rng.uniform(-np.pi, np.pi, bs)
        head[0] = 0.0
        head[1] = np.pi / 2
        pose = sensor_pose(gy, gx, head, size=s)
        want = sensor_crop(torch.from_numpy(maps), pose, half, origin).numpy()
        got = spec_sensor_crop(maps, pose.numpy(), torch.cos(pose[:, 2]).numpy(), torch.sin(pose[:, 2]).numpy(), half, origin)
        assert np.array_equal(want, got), s


def test_kernel_emulation_equals_oracle():
    """The host mirror of k_semcrop (same source as the device function) against the torch oracle, with the
    oracle's cos/sin; shared maps through map_index."""
    from emul import emul_semantic_crop
    rng = np.random.default_rng(2)
    for s, half, origin in ((480, 50, 289), (200, 30, 120), (96, 50, 60)):
        bs = 6
        maps = np.stack([_random_semmap(rng, s) if s >= 200 else rng.integers(0, 28, (s, s)) for _ in range(3)]).astype(np.float32)
        index = rng.integers(0, 3, bs).astype(np.int32)
        pose = sensor_pose(rng.uniform(s / 2 - 40, s / 2 + 40, bs), rng.uniform(s / 2 - 40, s / 2 + 40, bs),
                           rng.uniform(-np.pi, np.pi, bs), size=s)
        trig = torch.stack([torch.cos(pose[:, 2]), torch.sin(pose[:, 2])], 1).numpy()
        want = sensor_crop(torch.from_numpy(maps[index]), pose, half, origin).numpy()
        got = emul_semantic_crop(maps, pose.numpy(), trig, index, half, origin)
        assert np.array_equal(want, got), s


@pytest.mark.gpu
def test_gpu_semantic_crop():
    """CUDA kernel through the C ABI: bit-exact labels with the oracle's cos/sin; with sinf/cosf on the device at
    most a handful of cells sitting exactly on a rounding boundary may take the neighbouring label."""
    import wsmgmap_b200  # noqa: F401
    from wsmgmap_b200 import ops
    rng = np.random.default_rng(3)
    bs, s = 64, 480
    maps = np.stack([_random_semmap(rng, s) for _ in range(8)]).astype(np.float32)
    index = rng.integers(0, 8, bs).astype(np.int32)
    pose = sensor_pose(rng.uniform(150, 330, bs), rng.uniform(150, 330, bs), rng.uniform(-np.pi, np.pi, bs))
    trig = torch.stack([torch.cos(pose[:, 2]), torch.sin(pose[:, 2])], 1)
    want = sensor_crop(torch.from_numpy(maps[index]), pose)
    dev = "cuda:0"
    got = ops.semantic_crop(torch.from_numpy(maps).to(dev), pose.to(dev), trig=trig, map_index=torch.from_numpy(index))
    assert got.dtype == torch.int64 and torch.equal(got.cpu(), want)
    got2 = ops.semantic_crop(torch.from_numpy(maps).to(dev), pose.to(dev), map_index=torch.from_numpy(index))
    assert (got2.cpu() != want).float().mean().item() < 1e-3
