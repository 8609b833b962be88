"""VQ pixel-decoder oracle vs the reference modules' own outputs (tests/golden/vq_*.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import vq_oracle as VO
from tests._cases import vq_codes
from umgen_b200 import synth


@pytest.mark.parametrize("kind", ["map", "image"])
def test_vq_oracle_matches_reference(kind, golden_dir):
    g = np.load(os.path.join(golden_dir, f"vq_{kind}.npz"))
    sd = synth.make_vq_state_dict(kind, seed=1)
    with torch.no_grad():
        out = VO.decode_code(sd, kind, vq_codes(kind)[:1])
    np.testing.assert_allclose(out[:, :, ::4, ::4].numpy(), g["out"][:1], rtol=0, atol=1e-4)


def test_to_rgb_weights_do_not_touch_global_rng():
    from umgen_b200.vq import rgb_weights
    state = torch.random.get_rng_state()
    torch.manual_seed(0)
    want = torch.randn(3, 5, 1, 1).view(3, 5)
    torch.random.set_rng_state(state)
    before = torch.random.get_rng_state()
    got = rgb_weights(5, 0)
    assert torch.equal(got, want) and torch.equal(before, torch.random.get_rng_state())
