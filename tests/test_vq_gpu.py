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


# ---- implicit-GEMM convolution (csrc/gemm_sm100.cu, 4-D TMA boxes) ------------------------------------------------------------
CONV_SHAPES = [      # (B, H, W, Cin, Cout): every box geometry the decoders use (W = 32 .. 512) plus the narrowest one
    (2, 16, 32, 512, 512), (1, 32, 64, 256, 256), (2, 64, 128, 256, 128), (1, 4, 256, 128, 128), (3, 2, 512, 128, 128), (1, 16, 8, 64, 128),
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
@pytest.mark.parametrize("with_resid", [False, True])
def test_conv3x3_implicit_gemm(shape, with_resid):
    """umgen_conv3x3_f16 against (a) im2col + GEMM through the same tensor-core kernel: same products in the same K order, so bit-identical;
    (b) torch's fp32 conv2d on the same fp16 values (nn.Conv2d(.., 3, 1, 1), vq_modules.py:98-107): within fp16 output rounding."""
    import torch.nn.functional as F
    from umgen_b200 import ops
    B, H, W, Cin, Cout = shape
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B * 1000 + W)
    x = (torch.randn(B, H, W, Cin, generator=g) * 0.5).to(dev, torch.float16)
    wt = (torch.randn(Cout, Cin, 3, 3, generator=g) * (9 * Cin) ** -0.5).to(dev, torch.float16)
    w = wt.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    bias = torch.randn(Cout, generator=g).to(dev)
    resid = torch.randn(B * H * W, Cout, generator=g).to(dev, torch.float16) if with_resid else None
    epi = ops.EPI_RESID_F16 if with_resid else ops.EPI_BIAS_F16
    assert ops.conv3x3_supported(H, W, Cin, Cout)
    got = ops.conv3x3(x, w, bias, torch.empty(B * H * W, Cout, dtype=torch.float16, device=dev), B, H, W, Cin, epi, resid)
    a = torch.empty(B * H * W, 9 * Cin, dtype=torch.float16, device=dev)
    ops.im2col3x3(x, a, B, H, W, Cin, 9 * Cin, False)
    via_im2col = ops.gemm(a, w, bias, torch.empty_like(got), epi, resid)
    assert torch.equal(got, via_im2col)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(B * H * W, Cout)
    if with_resid:
        ref = ref + resid.float()
    err = (got.float() - ref).abs().max().item()
    assert err < 4e-3 * max(1.0, ref.abs().max().item()), err


def test_conv3x3_rejects_unsupported_geometry():
    from umgen_b200 import capi, ops
    dev = torch.device("cuda:0")
    assert not ops.conv3x3_supported(16, 32, 16, 512) and not ops.conv3x3_supported(6, 32, 64, 128) and not ops.conv3x3_supported(16, 48, 64, 128)
    x = torch.zeros(1, 6, 32, 64, dtype=torch.float16, device=dev)
    w = torch.zeros(128, 9 * 64, dtype=torch.float16, device=dev)
    with pytest.raises(capi.UmgenError):
        ops.conv3x3(x, w, None, torch.empty(6 * 32, 128, dtype=torch.float16, device=dev), 1, 6, 32, 64)


def test_upsample_then_conv_matches_the_folded_im2col():
    """Upsample.forward (vq_modules.py:34-40): nearest 2x + conv.  Materialised upsample + implicit GEMM == im2col with the upsample folded in."""
    from umgen_b200 import ops
    dev = torch.device("cuda:0")
    B, H, W, Cc = 2, 32, 64, 256      # output size
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, H // 2, W // 2, Cc, generator=g).to(dev, torch.float16)
    w = (torch.randn(Cc, 9 * Cc, generator=g) * (9 * Cc) ** -0.5).to(dev, torch.float16)
    bias = torch.randn(Cc, generator=g).to(dev)
    up = ops.upsample2x(x, torch.empty(B, H, W, Cc, dtype=torch.float16, device=dev), B, H // 2, W // 2, Cc)
    assert torch.equal(up, x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2))
    got = ops.conv3x3(up, w, bias, torch.empty(B * H * W, Cc, dtype=torch.float16, device=dev), B, H, W, Cc)
    a = torch.empty(B * H * W, 9 * Cc, dtype=torch.float16, device=dev)
    ops.im2col3x3(x, a, B, H, W, Cc, 9 * Cc, True)
    assert torch.equal(got, ops.gemm(a, w, bias, torch.empty_like(got), ops.EPI_BIAS_F16))


def test_conv_out_planes():
    """conv_out (vq_modules.py:330-334): Conv2d(128, 3 | 5, 3, 1, 1) on the tensor cores with the fp32-plane epilogue, against torch's fp32 conv2d."""
    import torch.nn.functional as F
    from umgen_b200 import ops
    dev = torch.device("cuda:0")
    for (B, H, W, Cin, n_out) in [(2, 8, 256, 128, 5), (1, 4, 512, 128, 3), (3, 32, 32, 128, 3)]:
        g = torch.Generator().manual_seed(n_out + W)
        x = (torch.randn(B, H, W, Cin, generator=g) * 0.5).to(dev, torch.float16)
        wt = (torch.randn(n_out, Cin, 3, 3, generator=g) * (9 * Cin) ** -0.5).to(dev, torch.float16)
        w = torch.zeros(128, 9 * Cin, dtype=torch.float16, device=dev)
        w[:n_out] = wt.permute(0, 2, 3, 1).reshape(n_out, 9 * Cin)
        bias = torch.zeros(128, device=dev)
        bias[:n_out] = torch.randn(n_out, generator=g).to(dev)
        out = torch.full((B, n_out, H, W), float("nan"), device=dev)
        ops.conv3x3_nchw(x, w, bias, out, B, H, W, Cin, n_out)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias[:n_out], padding=1)
        assert (out - ref).abs().max().item() < 2e-4          # fp32 accumulation and fp32 output: only the summation order differs


@pytest.mark.parametrize("shape", [(2, 512, 512), (4, 2048, 256), (1, 8192, 256), (2, 32768, 128), (1, 131072, 128), (3, 1000, 128)])
@pytest.mark.parametrize("swish", [False, True])
def test_groupnorm_slab_statistics(shape, swish):
    """GroupNorm(32, eps 1e-6) + swish (vq_modules.py:14-22) with the coalesced statistics pass: against torch in fp32, against the
    one-CTA-per-group kernel, and bit-identical when repeated (the slab partials are added in a fixed order)."""
    import torch.nn.functional as F
    from umgen_b200 import ops
    B, HW, Cc = shape
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(HW + Cc)
    x = (torch.randn(B, HW, Cc, generator=g) * 1.5 + 0.3).to(dev, torch.float16)
    gamma, beta = torch.randn(Cc, generator=g).to(dev), torch.randn(Cc, generator=g).to(dev)
    sc = torch.zeros(ops.groupnorm_scratch_floats(B, HW), dtype=torch.float32, device=dev)
    y1 = ops.groupnorm_slab(x, gamma, beta, torch.empty_like(x), sc, B, HW, Cc, swish)
    y2 = ops.groupnorm_slab(x, gamma, beta, torch.empty_like(x), sc, B, HW, Cc, swish)      # the tickets were left at zero
    assert torch.equal(y1, y2)
    assert int(sc[-B:].view(torch.int32).abs().sum()) == 0
    old = ops.groupnorm(x, gamma, beta, torch.empty_like(x), torch.empty(B * 64, dtype=torch.float32, device=dev), B, HW, Cc, swish)
    ref = F.group_norm(x.float().permute(0, 2, 1), 32, gamma, beta, eps=1e-6).permute(0, 2, 1)
    if swish:
        ref = ref * torch.sigmoid(ref)
    tol = 2e-3 * max(1.0, ref.abs().max().item())        # fp16 outputs: half an ulp at the largest value, plus the statistics' rounding
    assert (y1.float() - ref).abs().max().item() < tol
    assert (y1.float() - old.float()).abs().max().item() < tol


@pytest.mark.parametrize("kind", ["map", "image"])
def test_vq_decoder_paths_agree(kind):
    """CUDA-graph replay == eager launches (bit-identical); implicit-GEMM convolutions == im2col convolutions (bit-identical, same products in
    the same order); slab GroupNorm == per-group GroupNorm up to fp32 summation order."""
    from umgen_b200.vq import VQDecoder
    dec = VQDecoder(synth.make_vq_state_dict(kind, seed=1), kind)
    code = vq_codes(kind)
    a = dec.decode_code(code)
    a2 = dec.decode_code(code)            # second call: pure replay
    dec.use_graph = False
    b = dec.decode_code(code)
    assert torch.equal(a, b) and torch.equal(a, a2)
    dec.conv_out_tc = False             # conv_out with fp32 weights on the FMA pipe instead of fp16 weights on the tensor cores
    b32 = dec.decode_code(code)
    assert (b - b32).abs().max().item() < 3e-3
    dec.implicit_conv = False
    c = dec.decode_code(code)
    assert torch.equal(b32, c)
    dec.slab_groupnorm = False
    d = dec.decode_code(code)
    assert (c - d).abs().max().item() < 2e-2


def test_decode_tokens_whole_scene(golden_dir):
    """UMGen_PL.decode_tokens (model_pl.py:357-457) through umgen_b200.postprocess.decode_tokens: boxes / poses equal to the reference's values,
    maps and images in 6-frame pieces from the GPU decoders (the reference-module goldens of the first frames)."""
    from umgen_b200 import postprocess as P
    from umgen_b200.vq import Imagedecoder, Mapdecoder
    g = np.load(os.path.join(golden_dir, "postprocess.npz"))
    gm, gi = np.load(os.path.join(golden_dir, "vq_map.npz")), np.load(os.path.join(golden_dir, "vq_image.npz"))
    T = 8                                                            # two pieces: 6 + 2 frames
    rs = np.random.RandomState(2)
    mtok = rs.randint(0, 8192, size=(1, T, 1024))
    itok = rs.randint(0, 8192, size=(1, T, 512))
    mtok[0, :2] = vq_codes("map").reshape(2, 1024).numpy()
    itok[0, :2] = vq_codes("image").reshape(2, 512).numpy()
    pred = {"pose": np.resize(g["pose_tokens"], (1, T, 3)), "map": mtok, "bbox3d": np.resize(g["bbox_tokens"], (1, T, 660)), "image": itok}
    md, idec = Mapdecoder(synth.make_vq_state_dict("map", seed=1)), Imagedecoder(synth.make_vq_state_dict("image", seed=1))
    bboxes, anno, pose, real_pose, maps, image, map_tr = P.decode_tokens(pred, None, md, idec)
    assert len(bboxes) == T and pose.shape == (T, 3) and anno is None and real_pose is None and map_tr is None
    np.testing.assert_array_equal(np.stack(bboxes[:6]), g["bboxes"])
    assert maps.shape == (T, 3, 256, 256) and image.shape == (T, 3, 256, 512) and maps.device.type == "cpu" and image.device.type == "cpu"
    assert np.abs(image[:2, :, ::4, ::4].numpy() - gi["out"]).max() < ATOL
    for piece in (maps[:6], maps[6:]):                               # to_rgb's min-max is taken per piece
        assert float(piece.min()) == pytest.approx(-1.0, abs=1e-5) and float(piece.max()) == pytest.approx(1.0, abs=1e-5)
    # the same frames decoded alone (one piece of 2) differ from the 6-frame piece only by that piece-wide affine map
    alone = md.decode_maps(mtok[:, :2]).cpu()
    a, b = alone.flatten().double(), maps[:2].flatten().double()
    A = torch.stack([a, torch.ones_like(a)], dim=1)
    sol = torch.linalg.lstsq(A, b[:, None]).solution.flatten()
    assert (A @ sol - b).abs().max().item() < 1e-4
