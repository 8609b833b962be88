"""CPU oracle of the VQ pixel decoders.  TEST INFRASTRUCTURE ONLY (same rules as umgen_oracle.py).

Functional fp32 restatement of NormVQModel.decode_code (reference tokenizer/vq_model.py:87-101,123-145) and of
vq_modules.Decoder / ResnetBlock / AttnBlock / Upsample (tokenizer/vq_modules.py:14-176, 293-415) plus to_rgb
(tools/decode_map.py:25-30), consuming a reference-keyed state_dict.  Pinned against the unmodified reference
modules by tests/golden/vq_*.npz (oracle/make_golden.py)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

CONFIGS = {
    "map": dict(ch=128, ch_mult=(1, 2, 2, 4), num_res_blocks=2, attn_resolutions=(16,), resolution=256, post_quant_pad=0),
    "image": dict(ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolutions=(32,), resolution=512, post_quant_pad=1),
}


def swish(x):                                   # vq_modules.py:14-16
    return x * torch.sigmoid(x)


def gnorm(P, pre, x):                           # vq_modules.py:19-22
    return F.group_norm(x, 32, P[pre + ".weight"], P[pre + ".bias"], 1e-6)


def conv(P, pre, x, pad):
    return F.conv2d(x, P[pre + ".weight"], P[pre + ".bias"], padding=pad)


def resblock(P, pre, x):                        # vq_modules.py:108-127
    h = conv(P, pre + ".conv1", swish(gnorm(P, pre + ".norm1", x)), 1)
    h = conv(P, pre + ".conv2", swish(gnorm(P, pre + ".norm2", h)), 1)
    if pre + ".nin_shortcut.weight" in P:
        x = conv(P, pre + ".nin_shortcut", x, 0)
    return x + h


def attnblock(P, pre, x):                       # vq_modules.py:149-176
    h = gnorm(P, pre + ".norm", x)
    q, k, v = (conv(P, f"{pre}.{n}", h, 0) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    w_ = torch.bmm(q.reshape(b, c, -1).permute(0, 2, 1), k.reshape(b, c, -1)) * (int(c) ** -0.5)
    w_ = torch.softmax(w_, dim=2)
    h = torch.bmm(v.reshape(b, c, -1), w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + conv(P, pre + ".proj_out", h, 0)


def decode_code(P, kind: str, code: torch.Tensor) -> torch.Tensor:
    """code int [B,h,w] -> [B,out_ch,H,W] fp32."""
    cfg = CONFIGS[kind]
    quant = P["quantize.embedding.weight"][code.long()].permute(0, 3, 1, 2)           # vq_model.py:93-94
    x = conv(P, "post_quant_conv", quant, cfg["post_quant_pad"])                       # vq_model.py:88
    x = conv(P, "decoder.conv_in", x, 1)
    x = resblock(P, "decoder.mid.block_1", x)
    x = attnblock(P, "decoder.mid.attn_1", x)
    x = resblock(P, "decoder.mid.block_2", x)
    nres = len(cfg["ch_mult"])
    curr = cfg["resolution"] // 2 ** (nres - 1)
    for lvl in reversed(range(nres)):                                                   # vq_modules.py:399-406
        for ib in range(cfg["num_res_blocks"] + 1):
            x = resblock(P, f"decoder.up.{lvl}.block.{ib}", x)
            if curr in cfg["attn_resolutions"]:
                x = attnblock(P, f"decoder.up.{lvl}.attn.{ib}", x)
        if lvl != 0:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = conv(P, f"decoder.up.{lvl}.upsample.conv", x, 1)
            curr *= 2
    return conv(P, "decoder.conv_out", swish(gnorm(P, "decoder.norm_out", x)), 1)


def to_rgb(x: torch.Tensor, seed: int = 0) -> torch.Tensor:                          # tools/decode_map.py:25-30
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(3, x.shape[1], 1, 1, generator=g)
    y = F.conv2d(x, w)
    return 2.0 * (y - y.min()) / (y.max() - y.min()) - 1.0
