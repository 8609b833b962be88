"""Pack a reference-keyed state_dict into the device layout of include/umgen.h.

Matrices go to fp16 (the reference runs them under fp16 autocast, UMGen.py:1605); LayerNorm weights,
biases and embedding tables stay fp32 (autocast leaves them fp32); bf16 sinusoid tables are widened
to fp32."""
from __future__ import annotations

from typing import Dict, Mapping

import numpy as np
import torch

from .config import ModelConfig, TASK_ID

BOX_RANGES = np.array([(-64, 64), (-64, 64), (-5, 5), (0, 15), (0, 4), (0, 5), (-3.14, 3.14),
                       (-20, 20), (-15, 15), (-0.3, 0.3)], dtype=np.float64)


def box_value_lut() -> np.ndarray:
    """[1028, 10] float64 attribute-token -> value table: bin midpoints on linspace(0,1,1024) with
    out-of-range tokens clipped (reference tokenizer.py:332-354 via :679-687), then min-max
    un-normalisation (normalize.py:147-160 with config:126-137)."""
    bins = np.linspace(0.0, 1.0, 1024)
    tok = np.arange(1028)
    mid = (bins[np.clip(tok - 1, 0, 1023)] + bins[np.clip(tok, 0, 1023)]) / 2
    return mid[:, None] * (BOX_RANGES[:, 1] - BOX_RANGES[:, 0])[None, :] + BOX_RANGES[:, 0][None, :]


def pose_value_lut() -> np.ndarray:
    """[1024, 3] float32 pose-token -> (dx, dy, dheading): bin midpoints on linspace(-1,1,1024)
    (tokenizer.py:332-354) divided by float32(1/std), std = (10, 4, 1) (normalize.py:65-76)."""
    bins = np.linspace(-1.0, 1.0, 1024)
    tok = np.arange(1024)
    mid = (bins[np.clip(tok - 1, 0, 1023)] + bins[np.clip(tok, 0, 1023)]) / 2
    inv_std = 1.0 / np.array([10.0, 4.0, 1.0], dtype=np.float32)
    return (mid[:, None] / inv_std[None, :]).astype(np.float32)


def _h(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float16).contiguous()


def _f(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def token_table(sd: Mapping[str, torch.Tensor], which: str, dev) -> torch.Tensor:
    """[8192, 768] fp32 embedding of every map / image token: {map,img}_mlp_pre(codebook.weight)
    (reference UMGen.py:449-450, 465-466, 1067-1068, 1135-1136; GMLP = c_fc -> erf-GELU -> c_proj, module.py:723-728).
    The codebook is frozen and the GMLP is weights-only, so the composition is folded once at load time."""
    cb = _f(sd[f"{which}_codebook.weight"], dev)
    fc = _f(sd[f"{which}_mlp_pre.c_fc.weight"], dev)
    proj = _f(sd[f"{which}_mlp_pre.c_proj.weight"], dev)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out = torch.nn.functional.linear(torch.nn.functional.gelu(torch.nn.functional.linear(cb, fc)), proj)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return out.contiguous()


def pack_oar(sd: Mapping[str, torch.Tensor], cfg: ModelConfig, dev) -> Dict[str, torch.Tensor]:
    """Device tensors for UmgenDecodeArgs (weights part)."""
    t = "transformer."
    hs, fs = [], []
    for i in range(cfg.n_oar_layer):
        p = f"{t}OAR.{i}."
        hs.append(torch.cat([_h(sd[p + "temporal_attn.c_attn.weight"], dev).view(-1),
                             _h(sd[p + "temporal_attn.c_proj.weight"], dev).view(-1),
                             _h(sd[p + "mlp.c_fc.weight"], dev).view(-1),
                             _h(sd[p + "mlp.c_proj.weight"], dev).view(-1)]))
        fs.append(torch.cat([_f(sd[p + "ln_1.weight"], dev), _f(sd[p + "temporal_attn.c_attn.bias"], dev),
                             _f(sd[p + "temporal_attn.c_proj.bias"], dev), _f(sd[p + "ln_2.weight"], dev)]))
    out = {
        "oar_h": torch.stack(hs).contiguous(), "oar_f": torch.stack(fs).contiguous(),
        "ln_oar_f": _f(sd[t + "ln_oar.weight"], dev),
        "head_map_h": _h(sd[t + "head_ar_map.weight"], dev), "head_bbox_h": _h(sd[t + "head_ar_bbox3d.weight"], dev),
        "head_img_h": _h(sd[t + "head_ar_img.weight"], dev),
        "head_tar_bbox_h": _h(sd[t + "head_tar_bbox3d.weight"], dev),
        "map_table_f": token_table(sd, "map", dev), "img_table_f": token_table(sd, "img", dev),
        "be_f": _f(sd[t + "be.weight"], dev), "axe_f": _f(sd[t + "axe.weight"], dev),
        "tske_f": _f(sd[t + "tske.weight"][TASK_ID], dev), "fpe_f": _f(sd["fouier_pe"], dev),
        "box_lut_d": torch.from_numpy(box_value_lut()).to(dev).contiguous(),
    }
    return out
