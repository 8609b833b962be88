"""Host-side post-processing of a finished rollout: the step right after the hot path.

Vectorised numpy restatement of what the reference's Lightning module does with ``UMGen.inference``'s output before it
visualises a scene (reference ``tools/model_pl.py:235-241`` ``save_tokens``, ``:243-302`` the bbox3d / pose part of
``decode_tokens``; ``plugin/data/transforms/tokenizer.py:689-774`` ``BBox3DTokenizer.decode(keep_order=True, no_special=True)``,
``:332-354`` ``DigitalBinsTokenizer.decode``; ``plugin/data/transforms/normalize.py:65-76, 189-229``).  The reference loops
over frames, attributes and slots in Python; here one table lookup per array.  ``decode_tokens`` is the whole of
``UMGen_PL.decode_tokens`` (values here, pixels through ``umgen_b200/vq.py`` on the GPU).  Pinned against the reference's own output by
``tests/test_postprocess.py`` (golden made by ``oracle/make_golden.py``).  This is CPU code like the reference's -- the pixel
decoders (``umgen_b200/vq.py``) are the GPU part of ``decode_tokens``.
"""
from __future__ import annotations

import os
import pickle
from typing import Dict, List, Tuple

import numpy as np

from .weights import box_value_lut

N_SLOTS, SLOT_LEN, N_ATTR = 60, 11, 10
PAD_TOKEN = 1027
CATEGORY_START = 1024
CATEGORIES = ("vehicle", "bicycle", "pedestrian")      # projects/configs/category.txt; every other token decodes to "none"


def pose_value_lut64() -> np.ndarray:
    """[1024, 3] float64 pose-token -> (dx, dy, dheading) as ``decode_tokens`` computes it: bin midpoints on linspace(-1, 1, 1024)
    (token 0 -> bins[0]) divided by the float32 inverse std (0.1, 0.25, 1) in float64 (normalize.py:65-76)."""
    bins = np.linspace(-1.0, 1.0, 1024)
    tok = np.arange(1024)
    mid = (bins[np.clip(tok - 1, 0, 1023)] + bins[np.clip(tok, 0, 1023)]) / 2
    inv_std = (1.0 / np.array([10.0, 4.0, 1.0])).astype(np.float32)
    return mid[:, None] / inv_std[None, :].astype(np.float64)


def decode_pose(pose_tokens: np.ndarray) -> np.ndarray:
    """[..., 3] int -> [..., 3] float64 metres / radians (model_pl.py:288-291)."""
    t = np.clip(np.asarray(pose_tokens, dtype=np.int64), 0, 1023)
    lut = pose_value_lut64()
    return np.stack([lut[t[..., c], c] for c in range(3)], axis=-1)


def decode_bbox3d(bbox_tokens: np.ndarray) -> Tuple[List[np.ndarray], List[List[str]]]:
    """[T, 660] int (or [1, T, 660]) -> (T arrays [60, 10] float64: x, y, z, l, w, h, yaw, vx, vy, vz of every slot in slot order;
    T lists of 60 category names).  ``<pad>`` slots decode to the upper end of every range, like in the reference, and carry the
    category "none" (model_pl.py:262-276: non-pad tokens are clipped to [0, 1026] first)."""
    tok = np.asarray(bbox_tokens, dtype=np.int64)
    if tok.ndim == 3:
        tok = tok[0]
    tok = tok.copy()
    nonpad = tok != PAD_TOKEN
    tok[nonpad] = np.clip(tok[nonpad], 0, PAD_TOKEN - 1)
    slots = tok.reshape(tok.shape[0], N_SLOTS, SLOT_LEN)
    lut = box_value_lut()                                   # [1028, 10]
    attr = slots[:, :, :N_ATTR]
    vals = lut[attr, np.arange(N_ATTR)[None, None, :]]      # [T, 60, 10]
    cat = slots[:, :, N_ATTR] - CATEGORY_START
    names = [[CATEGORIES[c] if 0 <= c < len(CATEGORIES) else "none" for c in row] for row in cat.tolist()]
    return [v for v in vals], names


def decode_scene(pred_tokens: Dict[str, np.ndarray]) -> Dict[str, object]:
    """The value part of ``decode_tokens`` for one scene: pred_tokens as returned by ``UMGenEngine.inference`` ([1, T, S_mod] int64)."""
    boxes, classes = decode_bbox3d(pred_tokens["bbox3d"][0])
    return {"bboxes": boxes, "bbox_classes": classes, "pose_values": decode_pose(pred_tokens["pose"][0])}


def decode_annotation_bbox3d(gt_bbox_tokens: np.ndarray) -> Tuple[List[np.ndarray], List[List[str]]]:
    """The ground-truth side of ``decode_tokens`` (model_pl.py:278-286: ``bbox3d_tokenizer.decode(gt, no_special=True)`` without
    ``keep_order``, then ``unnormalize_bbox3d``): per frame only the slots that hold no ``<pad>`` token at all survive
    (tokenizer.py:804-806), then those whose category id is in range (``del_unreason_tokens``, tokenizer.py:661-670); the survivors
    decode like the predicted ones.  Returns T arrays [n_t, 10] float64 (n_t may be 0) and T lists of n_t category names."""
    tok = np.asarray(gt_bbox_tokens, dtype=np.int64)
    if tok.ndim == 3:
        tok = tok[0]
    lut = box_value_lut()
    boxes, names = [], []
    for frame in tok:
        slots = frame.reshape(-1, SLOT_LEN)
        slots = slots[~np.any(slots == PAD_TOKEN, axis=1)]
        cat = slots[:, N_ATTR] - CATEGORY_START
        keep = ~((cat < 0) | (cat > len(CATEGORIES)))
        slots, cat = slots[keep], cat[keep]
        if slots.shape[0] == 0:
            boxes.append(np.zeros((0, N_ATTR)))
            names.append([])
            continue
        boxes.append(lut[slots[:, :N_ATTR], np.arange(N_ATTR)[None, :]])
        names.append([CATEGORIES[c] for c in cat.tolist()])
    return boxes, names


def decode_tokens(pred_tokens: Dict[str, np.ndarray], gt_tokens: Dict[str, np.ndarray] = None, mapdecoder=None, imagedecoder=None,
                  chunk: int = 6):
    """``UMGen_PL.decode_tokens`` (model_pl.py:357-457) for one scene: the same 7-tuple
    ``(bboxes, anno_bboxes, pose_values, real_pose, maps, decoded_image, map_transformed)``.

    pred_tokens: what ``UMGen.inference`` returned ({mod: int64 [1, T, S_mod]}); gt_tokens: the dataset's tokens of the same scene or
    None; mapdecoder / imagedecoder: ``umgen_b200.vq.Mapdecoder`` / ``Imagedecoder`` (or the ``projects.tools.decode_map`` wrappers) --
    the pixel part runs on the GPU in ``chunk``-frame pieces like the reference's (min-max of ``to_rgb`` per piece, model_pl.py:418-442),
    results on the host.  Without a decoder the corresponding entry is None (the value part needs no GPU)."""
    bboxes, _ = decode_bbox3d(pred_tokens["bbox3d"])
    pose_values = decode_pose(np.asarray(pred_tokens["pose"])[0])
    anno_bboxes = real_pose = None
    if gt_tokens is not None:
        anno_bboxes, _ = decode_annotation_bbox3d(np.asarray(gt_tokens["bbox3d"]))
        real_pose = decode_pose(np.asarray(gt_tokens["pose"])[0])
    maps = decoded_image = map_transformed = None
    if mapdecoder is not None and "map" in pred_tokens:
        import torch
        t = np.asarray(pred_tokens["map"])
        maps = torch.cat([mapdecoder.decode_maps(t[:, i:i + chunk]).cpu() for i in range(0, t.shape[1], chunk)], dim=0)
        if "map_transformed" in pred_tokens:
            map_transformed = mapdecoder.decode_maps(np.asarray(pred_tokens["map_transformed"]))
    if imagedecoder is not None and "image" in pred_tokens:
        import torch
        t = np.asarray(pred_tokens["image"])
        decoded_image = torch.cat([imagedecoder.decode_images(t[:, i:i + chunk]).cpu() for i in range(0, t.shape[1], chunk)], dim=0)
    return bboxes, anno_bboxes, pose_values, real_pose, maps, decoded_image, map_transformed


def save_tokens(out_tokens: Dict[str, np.ndarray], token_save_path: str, file_name: str) -> str:
    """``<token_save_path>/<file_name>_tokens.pkl`` holding the dict ``inference`` returned (model_pl.py:235-241)."""
    os.makedirs(token_save_path, exist_ok=True)
    path = os.path.join(token_save_path, file_name + "_tokens.pkl")
    with open(path, "wb") as f:
        pickle.dump(out_tokens, f)
    return path


def load_tokens(path: str) -> Dict[str, np.ndarray]:
    with open(path, "rb") as f:
        return pickle.load(f)
