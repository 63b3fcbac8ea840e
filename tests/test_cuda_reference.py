"""The drop-in against what users of the reference actually run: the reference's PyTorch ops ON CUDA.

Parity (bit-exact indices, 1e-5 maps) is defined and tested against the reference's CPU result.  On CUDA the same torch
ops differ from that in the last bit here and there (`x / 0.12` becomes a multiplication by the reciprocal,
rgb_mapping.py:173-174; sin / cos come from another library), which can move a pixel that sits on a cell boundary into
the neighbouring cell.  This test measures how often, on the four synthetic depth kinds at the real shapes, so that
the number is known (bench.py reports the same figures as `cuda_reference_flips`)."""
import pytest
import torch

from oracle.cuda_flips import flips_against_cuda_reference
from oracle.mapping_oracle import OracleMapper
import wsmgmap_b200  # noqa: F401
from wsmgmap_b200 import ops
from wsmgmap_b200.synth import DEPTH_KINDS, make_depth, make_features

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def test_index_flips_against_the_reference_on_cuda_are_rare():
    res = flips_against_cuda_reference()
    cpu = OracleMapper(8, 64)
    for kind, (flipped, total, dmax, scale) in res.items():
        print(f"{kind}: {flipped} of {total} sampled pixels change cell vs reference-on-CUDA ({flipped / total:.2e}); "
              f"max |ego difference| {dmax:.3e} (max |ego| {scale:.3f})")
        assert flipped / total < 1e-3, (kind, flipped, total)
    # and none at all against the reference on the CPU, the parity target (also pinned by the golden vectors)
    gen = torch.Generator().manual_seed(0)
    depth = make_depth("uniform", 8, 256, 256, gen)
    lin, inv = ops.unproject_index(depth.to(DEV), 224, 224)
    lin_c, inv_c = cpu.stage_cells(depth, 224, 224)
    assert torch.equal(lin.cpu().long(), lin_c) and torch.equal(inv.cpu(), inv_c)
