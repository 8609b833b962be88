"""GPU parity of the VQ pixel decoders against the reference modules' outputs (tests/golden/vq_*.npz)."""
import os

import numpy as np
import pytest
import torch

from tests._cases import vq_codes
from umgen_b200 import synth

pytestmark = pytest.mark.gpu
ATOL = 3e-2        # fp16 activations through ~30 residual blocks vs the fp32 reference; outputs are O(0.3)


@pytest.mark.parametrize("kind", ["map", "image"])
def test_vq_decoder_matches_reference(kind, golden_dir):
    from umgen_b200.vq import VQDecoder
    g = np.load(os.path.join(golden_dir, f"vq_{kind}.npz"))
    dec = VQDecoder(synth.make_vq_state_dict(kind, seed=1), kind)
    out = dec.decode_code(vq_codes(kind))
    got = out[:, :, ::4, ::4].cpu().numpy()
    err = np.abs(got - g["out"])
    print(f"{kind}: max abs err {err.max():.3e}, mean abs err {err.mean():.3e}, reference mean |x| {g['absmean']:.3f}")
    assert err.max() < ATOL and err.mean() < ATOL / 10


def test_map_decoder_rgb_matches_reference(golden_dir):
    from umgen_b200.vq import Mapdecoder
    g = np.load(os.path.join(golden_dir, "vq_map.npz"))
    md = Mapdecoder(synth.make_vq_state_dict("map", seed=1))
    rgb = md.decode_maps(vq_codes("map").reshape(2, 1024))
    assert rgb.shape == (2, 3, 256, 256)
    got = rgb[:, :, ::4, ::4].cpu().numpy()
    assert float(rgb.min()) == pytest.approx(-1.0, abs=1e-5) and float(rgb.max()) == pytest.approx(1.0, abs=1e-5)
    assert np.abs(got - g["rgb"]).max() < 5e-2


def test_image_decoder_shape_and_chunking(golden_dir):
    from umgen_b200.vq import Imagedecoder
    idec = Imagedecoder(synth.make_vq_state_dict("image", seed=1))
    toks = vq_codes("image").reshape(2, 512)
    a = idec.decode_images(toks)
    b = torch.cat([idec.decode_images(toks[i:i + 1]) for i in range(2)])
    assert a.shape == (2, 3, 256, 512)
    assert torch.equal(a, b)       # chunking does not change results
