"""Property tests (hypothesis) of the scatter semantics the reference relies on (rgb_mapping.py:207-230 plus
torch_scatter.scatter_max): untouched cells -> 0, invalid pixels send -1e16 to cell 0, cells equal to -1e16
become 0, negative-only cells stay negative, -0.0 -> +0.0.  Kernel emulation vs the elementwise spec vs the
torch oracle."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from emul import emul_cells, emul_step
from oracle.mapping_oracle import MapGeometry, OracleMapper, spec_cells, spec_scatter

HF = HD = 8
SPECIAL = [0.0, -0.0, -1e16, 1e-30, -1e-30, 3.0, -3.0, -2e16, 65504.0]


@st.composite
def frames(draw):
    depth = draw(st.lists(st.sampled_from([0.0, 0.004, 0.05, 0.11, 0.23, 0.37, 0.52, 0.9]), min_size=HD * HD, max_size=HD * HD))
    vals = draw(st.lists(st.one_of(st.sampled_from(SPECIAL), st.floats(-8, 8, width=32)), min_size=4 * HF * HF, max_size=4 * HF * HF))
    return (np.array(depth, np.float32).reshape(1, HD, HD), np.array(vals, np.float32).reshape(1, 4, HF, HF))


@settings(max_examples=60, deadline=None)
@given(frames())
def test_scatter_semantics(fr):
    depth, feat = fr
    geo = MapGeometry()
    lin, inv, _ = emul_cells(depth, HF)
    slin, sinv = spec_cells(depth, HF, HF, geo)
    assert np.array_equal(lin, slin) and np.array_equal(inv, sinv)
    _, proj = emul_step(None, feat, depth, np.zeros((1, 2), np.float32), np.zeros((1, 1), np.float32), np.zeros((1, 1), np.float32), mode=1)
    want, occ = spec_scatter(feat, slin, sinv, geo)
    assert np.array_equal(proj, want)
    assert not np.signbit(proj[proj == 0]).any()                         # -0.0 never survives
    orc = OracleMapper(1, 4)
    from oracle.mapping_oracle import project_to_ego
    tp, tocc = project_to_ego(torch.from_numpy(feat), torch.from_numpy(slin), torch.from_numpy(sinv), geo)
    assert np.array_equal(tp.numpy(), want) and np.array_equal(tocc.numpy(), occ)
    untouched = ~occ.reshape(1, 1, 100, 100).repeat(4, 1)
    untouched[:, :, 0, 0] = False                                        # cell 0 also receives the sentinel
    assert not proj[untouched].any()
