"""Drop-in factories of the VQ pixel decoders (reference projects/tokenizer/vq_model.py:178-202): same names,
same ``{"state_dict": ...}`` checkpoint format (vq_model.py:65-78), objects exposing ``eval()``,
``decode_code(idx)``, ``indices_to_quant(idx)`` and ``decode(quant)`` (vq_model.py:87-101) backed by the B200 kernels."""
from __future__ import annotations

import torch

from umgen_b200.vq import VQDecoder


class _NormVQDecodeOnly:
    def __init__(self, kind: str, ckpt, device):
        sd = torch.load(ckpt, map_location="cpu")["state_dict"]
        dev = device if str(device) != "cuda" else f"cuda:{torch.cuda.current_device()}"
        self._dec = VQDecoder(sd, kind, dev)

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def decode_code(self, code_b):
        return self._dec.decode_code(code_b)

    def indices_to_quant(self, indices):
        """vq_model.py:98-101: codebook rows of `indices` [B, h, w] as a [B, 16, h, w] tensor."""
        idx = torch.as_tensor(indices).to(self._dec.dev).long()
        return self._dec.raw_codebook[idx].permute(0, 3, 1, 2).contiguous()

    def decode(self, quant):
        """vq_model.py:87-90 for a `quant` made of codebook rows (what indices_to_quant returns): the B200 decoder gathers from the codebook
        itself, so the rows are mapped back to their indices (exact match required; an arbitrary latent is not supported)."""
        q = torch.as_tensor(quant).to(self._dec.dev).float().permute(0, 2, 3, 1)              # [B, h, w, 16]
        cb = self._dec.raw_codebook
        idx = (2.0 * q.reshape(-1, 16) @ cb.t() - (cb * cb).sum(dim=1)[None]).argmax(dim=1)       # nearest row
        if not torch.equal(cb[idx], q.reshape(-1, 16)):
            raise ValueError("decode(quant): quant is not made of codebook rows; use decode_code(indices)")
        return self._dec.decode_code(idx.view(q.shape[:3]))


def get_normvq_dim16_res512_f16(device: str = "cuda", ckpt=None):
    return _NormVQDecodeOnly("image", ckpt or "data/image_decoder.pt", device)


def get_map_normvq_dim16_res256_f8(device: str = "cuda", ckpt=None):
    return _NormVQDecodeOnly("map", ckpt or "data/weights/map_vq", device)
