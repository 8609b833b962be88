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

    def indices_to_quant(self, indices):            # the decoder consumes indices directly; keep them as the "quant" handle
        return indices

    def decode(self, quant):
        return self._dec.decode_code(quant)


def get_normvq_dim16_res512_f16(device: str = "cuda", ckpt=None):
    return _NormVQDecodeOnly("image", ckpt or "data/image_decoder.pt", device)


def get_map_normvq_dim16_res256_f8(device: str = "cuda", ckpt=None):
    return _NormVQDecodeOnly("map", ckpt or "data/weights/map_vq", device)
