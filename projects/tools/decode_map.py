"""Drop-in ``Mapdecoder`` / ``Imagedecoder`` (reference projects/tools/decode_map.py:110-183): same constructor
``(ckpt, device)`` and ``decode_maps`` / ``decode_images`` methods, backed by the B200 VQ decoders."""
from __future__ import annotations

import torch

from umgen_b200 import vq as _vq
from umgen_b200.visualize import frame_label as add_frame_number, to_uint8 as postprocess_image, write_video_single  # noqa: F401  (decode_map.py:39-107)


def _load(ckpt):
    return torch.load(ckpt, map_location="cpu")["state_dict"]


class Mapdecoder(_vq.Mapdecoder):
    def __init__(self, ckpt, device="cuda"):
        super().__init__(_load(ckpt), device if str(device) != "cuda" else f"cuda:{torch.cuda.current_device()}")


class Imagedecoder(_vq.Imagedecoder):
    def __init__(self, ckpt, device="cuda"):
        super().__init__(_load(ckpt), device if str(device) != "cuda" else f"cuda:{torch.cuda.current_device()}")
