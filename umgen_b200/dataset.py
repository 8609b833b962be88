"""The step right before the hot path: raw tokenised nuPlan scene pickles -> the token tensors
``UMGen.inference`` is conditioned on.

Restates, vectorised where the arithmetic allows it, what the reference does between ``evaluate.py`` building its dataset and
``UMGen_PL.test_step`` receiving a batch: ``plugin/data/datasets/UMGen_nuplan_dataset.py:139-417`` (``NuPlanTokenDataset``: frame
sampling, ego-motion deltas, category / range filter) followed by the evaluation transform list of
``configs/UMGen_config_evaluation.py:186-255`` (``SplitAttriute`` -> ``Normalize`` -> ``MergeAttribute`` -> ``Normalize_Standard`` ->
``BBox3DTokenizer`` (``plugin/data/transforms/tokenizer.py:442-660, 810-950``: digitise, category ids, track-id slotting into 60 slots) ->
``DigitalBinsTokenizer`` (``tokenizer.py:254-330``) -> ``ToTensor``).  The dtype flow is the reference's (float32 boxes normalised in
float32 and digitised against float64 bin edges; float64 ego deltas scaled by a float32 inverse std), so the tokens are identical, not
close: ``tests/test_dataset.py`` compares with the reference's own class on synthetic raw scenes (``tests/golden/dataset.npz``, made by
``oracle/make_golden.py dataset``).  Pure host code, like the reference's; nothing here touches the GPU.

Raw scene schema: ``{"tokens": {view: {"tokens": [n x int[16,32]], "file_list": [...]}}, "raster_tokens": int[n,32,32],
"ego_pose_all": float[n,16] (column 6 = heading), "meta_info": [n x {"T_lidar2global": float[4,4], "bboxes_3d": float[k,>=10],
"track_ids": int[k], "categories": [k x str]}], "lidar_bboxes": {...}}``.
"""
from __future__ import annotations

import os
import pickle
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

CATEGORIES = ("vehicle", "bicycle", "pedestrian")           # projects/configs/category.txt
N_SLOTS, SLOT_LEN, N_ATTR = 60, 11, 10
PAD_TOKEN, CATEGORY_START = 1027, 1024
# configs/UMGen_config_evaluation.py:128-139 (x, y, z, l, w, h, yaw, vx, vy, vz)
BOX_RANGE = ((-64, 64), (-64, 64), (-5, 5), (0, 15), (0, 4), (0, 5), (-3.14, 3.14), (-20, 20), (-15, 15), (-0.3, 0.3))
POSE_STD = (10.0, 4.0, 1.0)                                 # Normalize_Standard(mean 0, std): dx, dy, dheading


def frame_indices(seq_len: int, block_size: int, sampling_gap: int, start_index: int = 10, inference: bool = True) -> List[int]:
    """``NuPlanTokenDataset.get_frame_indices`` (UMGen_nuplan_dataset.py:139-169): ``block_size`` frames ``sampling_gap`` apart from
    ``start_index`` (4 when training), the start pulled forward -- or the clip shortened -- when the scene is too short."""
    start = start_index if inference else 4
    max_start = seq_len - block_size * sampling_gap - sampling_gap
    if max_start < sampling_gap:
        max_start = sampling_gap
        block_size = (seq_len - sampling_gap - 1) // sampling_gap
    start = min(start, max_start)
    return [start + i * sampling_gap for i in range(block_size)]


def ego_deltas(meta_info: Sequence[dict], ego_pose_all: np.ndarray, frames: Sequence[int], sampling_gap: int) -> np.ndarray:
    """[T, 3] float64 (dx, dy, dheading): the translation of the NEXT sample's lidar origin in the previous sample's lidar frame and the
    heading change wrapped into [-pi, pi) (UMGen_nuplan_dataset.py:252-275, 300-302).  Row i describes the step INTO frame i."""
    out = np.empty((len(frames), 3), dtype=np.float64)
    origin = np.array([0, 0, 0, 1.0])
    for i, f in enumerate(frames):
        index = f - sampling_gap if i == 0 else frames[i - 1]
        assert index >= 0
        tr = np.linalg.inv(meta_info[index]["T_lidar2global"]) @ (meta_info[index + sampling_gap]["T_lidar2global"] @ origin.T)
        h = ego_pose_all[index + sampling_gap, 6] - ego_pose_all[index, 6]
        if h >= np.pi:
            h -= 2 * np.pi
        if h < -np.pi:
            h += 2 * np.pi
        out[i] = (tr[0], tr[1], h)
    return out


def pose_tokens(deltas: np.ndarray) -> np.ndarray:
    """Normalize_Standard (normalize.py:49-62: float32 mean / inverse std) then DigitalBinsTokenizer.encode on linspace(-1, 1, 1024)
    (tokenizer.py:316-330): int64 [T, 3] in [0, 1023]."""
    mean = np.zeros(3, dtype=np.float32)
    inv_std = 1.0 / np.array(POSE_STD, dtype=np.float32)
    x = (deltas - mean) * inv_std
    return np.clip(np.digitize(x, np.linspace(-1.0, 1.0, 1024)), 0, 1023)


def box_attribute_tokens(boxes: np.ndarray) -> np.ndarray:
    """float32 [k, >=10] metric boxes -> int64 [k, 10] attribute tokens: min-max to [0, 1] column by column in float32 (Normalize.normalize,
    normalize.py:124-139: the python-float range is a weak scalar, the quotient stays float32), digitised against the float64 edges
    linspace(0, 1, 1024) and clipped (values outside the range land in the first / last bin)."""
    b = np.asarray(boxes, dtype=np.float32)[:, :N_ATTR]
    lo = np.array([r[0] for r in BOX_RANGE], dtype=np.float32)
    span = np.array([r[1] - r[0] for r in BOX_RANGE], dtype=np.float64).astype(np.float32)
    x = (b - lo) / span
    return np.clip(np.digitize(x, np.linspace(0.0, 1.0, 1024)), 0, 1023)


def slot_boxes(frame_tokens: Sequence[np.ndarray], frame_track_ids: Sequence[np.ndarray]) -> np.ndarray:
    """``BBox3DTokenizer.bbox_slotting`` (tokenizer.py:810-950, no shuffle: ``shift_object_order_pro=0``): every track id of the clip gets
    one of the 60 slots in order of first appearance (ids beyond the 60th are dropped), a frame's boxes go to their tracks' slots, everything
    else is ``<pad>``.  Quirks kept: a frame whose ids are all zero counts as empty (``np.any``), a duplicated id keeps its last box.
    Returns int64 [T, 660]."""
    T = len(frame_tokens)
    ids_all = [np.asarray(t) for t in frame_track_ids]
    live = [t for t in ids_all if np.any(t)]
    order = np.concatenate(live) if live else np.array([])
    if np.any(order):
        _, first = np.unique(order, return_index=True)
        order = order[np.sort(first)]
    order = order[:N_SLOTS]
    slot_of = {tid: i for i, tid in enumerate(order.tolist())}
    out = np.full((T, N_SLOTS, SLOT_LEN), PAD_TOKEN, dtype=np.int64)
    for t in range(T):
        ids = ids_all[t]
        if not np.any(ids):
            continue
        keep = [i for i, tid in enumerate(ids.tolist()) if tid in slot_of]
        if not keep or not np.any(ids[keep]):
            continue
        out[t, [slot_of[ids[i].item()] for i in keep]] = np.asarray(frame_tokens[t])[keep]
    return out.reshape(T, N_SLOTS * SLOT_LEN)


def bbox3d_tokens(boxes: Sequence[np.ndarray], cats: Sequence[Sequence[str]], track_ids: Sequence[np.ndarray],
                  categories: Sequence[str] = CATEGORIES) -> np.ndarray:
    """Per frame: keep the boxes whose category is in the vocabulary and whose centre lies within 64 m in x and y
    (UMGen_nuplan_dataset.py:318-345), tokenise the 10 attributes + the category (1024 + index), slot by track id.  int64 [T, 660]."""
    toks, ids = [], []
    for b, c, t in zip(boxes, cats, track_ids):
        b = np.array(b).astype(np.float32)
        keep = [j for j, name in enumerate(c) if name in categories and not (abs(b[j][0]) > 64 or abs(b[j][1]) > 64)]
        if not keep:
            toks.append(np.zeros((0, SLOT_LEN), dtype=np.int64))
            ids.append(np.array([]))
            continue
        cat = np.array([categories.index(c[j]) for j in keep], dtype=np.int64) + CATEGORY_START
        toks.append(np.concatenate([box_attribute_tokens(b[keep]), cat[:, None]], axis=-1))
        ids.append(np.array(t)[keep])
    return slot_boxes(toks, ids)


class NuPlanTokenScenes:
    """Counterpart of ``NuPlanTokenDataset`` as ``tools/infer_fun.py:189-213`` configures it for ``evaluate.py`` (test mode, evaluation
    transforms, ``return_scene_name=True``): ``scenes[i]`` is the dict the reference's ``__getitem__`` returns --
    ``pose`` int64 [T, 3], ``map`` [T, 1024], ``pose_diff`` float32 [T, 3] (the raw deltas), ``bbox3d`` int64 [T, 660],
    ``image`` [T, 512] (with ``sample_img``), ``file_name`` -- and ``batch(i)`` adds the leading batch axis a ``DataLoader(batch_size=1)``
    would.  With ``control_test`` the pickle holds finished tokens and is returned as it is (UMGen_nuplan_dataset.py:204-209)."""

    def __init__(self, data_root: Sequence[str], block_size: int, sampling_gap: int = 4, start_index: int = 10, inference_flag: bool = True,
                 views: Sequence[str] = ("CAM_F0",), categories: Optional[Sequence[str]] = None, categories_file: Optional[str] = None,
                 sample_img: bool = True, control_test: bool = False, return_scene_name: bool = True):
        if isinstance(data_root, str):
            data_root = [data_root]
        self.files: List[str] = []
        for path in data_root:
            if os.path.isfile(path) and path.endswith(".pkl"):
                self.files.append(path)
                continue
            self.files += [os.path.join(path, f) for f in os.listdir(path) if f.endswith(".pkl")]
        self.files = sorted(self.files)
        if categories is None and categories_file is not None:
            categories = [l.strip() for l in open(categories_file) if l.strip()]
        self.categories = tuple(categories) if categories is not None else CATEGORIES
        self.block_size, self.sampling_gap, self.start_index, self.inference_flag = block_size, sampling_gap, start_index, inference_flag
        self.views, self.sample_img, self.control_test, self.return_scene_name = tuple(views), sample_img, control_test, return_scene_name

    def __len__(self) -> int:
        return len(self.files)

    def __getitem__(self, idx: int):
        path = self.files[idx]
        with open(path, "rb") as f:
            scene = pickle.load(f)
        if self.control_test:
            return scene
        image = np.stack(scene["tokens"][self.views[0]]["tokens"], axis=0)
        frames = frame_indices(image.shape[0], self.block_size, self.sampling_gap, self.start_index, self.inference_flag)
        meta = scene["meta_info"]
        deltas = ego_deltas(meta, scene["ego_pose_all"], frames, self.sampling_gap)
        data: Dict[str, object] = {
            "pose": torch.from_numpy(pose_tokens(deltas)),
            "map": torch.from_numpy(np.asarray(scene["raster_tokens"])[frames].reshape(len(frames), -1)),
            "pose_diff": torch.from_numpy(deltas.astype(np.float32)),       # ToTensor narrows float64 arrays (normalize.py:265-269)
            "bbox3d": torch.from_numpy(bbox3d_tokens([meta[f]["bboxes_3d"] for f in frames], [meta[f]["categories"] for f in frames],
                                                     [meta[f]["track_ids"] for f in frames], self.categories)),
        }
        if self.sample_img:
            data["image"] = torch.from_numpy(image[frames].reshape(len(frames), -1))
        if self.return_scene_name:
            data["file_name"] = str(idx) + "_" + path
        return data

    def batch(self, idx: int):
        """What ``DataLoader(dataset, batch_size=1)`` hands to ``UMGen_PL.test_step`` (evaluate.py:196-203)."""
        from torch.utils.data import default_collate
        return default_collate([self[idx]])
