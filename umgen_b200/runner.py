"""The caller of the hot path: one dataset batch -> ``UMGen.inference`` -> token pickle -> decoded values / pixels.

Counterpart of ``UMGen_PL.world_model_evaluate`` and ``generate_init_tokens`` (reference ``tools/model_pl.py:95-262``) without the Lightning
harness; the scene video is written when a ``umgen_b200.visualize.SceneVideo`` is handed in (``generate_videos``, model_pl.py:283-314): the same two branches (free rollout of a dataset scene; controlled rollout of a ``controlled_scenes`` pickle),
the same keyword arguments to ``inference`` (pinned by ``tests/test_runner.py`` against what the reference's own method passes to a recording
model; golden from ``oracle/make_golden.py runner``), the same ``<name>_tokens.pkl`` and the ``decode_tokens`` 7-tuple.  Host glue; the model
is ``projects.models.UMGen.UMGen`` (or anything with its ``inference``), the decoders are ``umgen_b200.vq.Mapdecoder`` / ``Imagedecoder``.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Dict, List, Optional, Sequence

import torch

from . import postprocess


@dataclasses.dataclass
class RunSettings:
    """What ``UMGen_PL.__init__`` reads from its config (model_pl.py:20-58)."""
    new_frames: int = 30                       # -1: up to the end of the scene's tokens
    cond_frames: int = 20
    input_cond_frames: int = 20
    pred_task: str = "pose_map_bbox3d_image"
    infer_task: str = "video"                  # "control" in the name selects the controlled branch (model_pl.py:43-46)
    infer_from_gt: bool = False
    init_token_mod: Optional[Sequence[str]] = None      # e.g. ("pose", "map"): ground-truth tokens handed in as init tokens (FID / MMD runs)
    token_save_path: Optional[str] = None      # where <name>_tokens.pkl goes; None: nothing is written and nothing is skipped
    generate_video: bool = False               # dataset scenes: a video for every scene instead of every 100th (model_pl.py:260); controlled scenes always get one

    @property
    def control_test(self) -> bool:
        return "control" in self.infer_task


def generate_init_tokens(gt_tokens: Dict[str, torch.Tensor], input_cond_frames: int = 20, init_token_mod: Optional[Sequence[str]] = None,
                         control_tokens: Optional[Dict[str, torch.Tensor]] = None, device=None) -> Optional[Dict[str, torch.Tensor]]:
    """model_pl.py:95-130.  Ground-truth init tokens are the frames after the conditioning ones; control tokens get a batch axis when they come
    without one and go to ``device`` (the reference calls ``.cuda()``); neither -> None."""
    if init_token_mod is not None:
        return {m: gt_tokens[m][:, input_cond_frames:, ...].clone() for m in init_token_mod}
    if control_tokens is not None:
        out = {}
        for m, v in control_tokens.items():
            v = v if v.dim() == 3 else v[None, ...]
            out[m] = v.to(device) if device is not None else v
        return out
    return None


def scene_name_of(batch: dict, control: bool) -> str:
    if control:
        return batch["scene_name"][0]
    name = batch["file_name"][0]
    return name.split("/")[-1][:-4]          # "<idx>_<path>/<scene>.pkl" -> "<scene>" (model_pl.py:203-209)


def inference_kwargs(batch: dict, s: RunSettings, device=None) -> dict:
    """The keyword arguments ``world_model_evaluate`` builds for ``model.inference`` (model_pl.py:139-239), both branches."""
    kw = dict(new_frames=s.new_frames, cond_frames=s.cond_frames, pred_task=s.pred_task, cond_on_par=True, infer_from_gt=s.infer_from_gt)
    if s.control_test:
        gt = batch["dataset_token"]
        init = generate_init_tokens(dict(gt), s.input_cond_frames, s.init_token_mod, batch["control_dict"], device)
        kw["input_cond_tokens"] = gt
        controlled = "no_control" not in scene_name_of(batch, True)
        kw["control_test"] = controlled
        kw["init_tokens"] = init if controlled else None
        kw["input_cond_frames"] = batch["input_cond_frame"] if "input_cond_frame" in batch else s.input_cond_frames
        return kw
    kw["input_cond_tokens"] = batch
    kw["init_tokens"] = generate_init_tokens(batch, s.input_cond_frames, s.init_token_mod, None, device)
    kw["input_cond_frames"] = s.input_cond_frames
    kw["control_test"] = False
    if kw["new_frames"] == -1:
        kw["new_frames"] = batch["bbox3d"].shape[1] - s.input_cond_frames
    return kw


def write_scene_video(video, decoded, name: str) -> str:
    """``UMGen_PL.generate_videos`` (model_pl.py:283-314): predicted boxes, poses, annotated poses, decoded map and camera frames go to the
    compositor; the annotation boxes of the 7-tuple are not drawn (the reference does not pass them on either)."""
    import numpy as np
    bboxes, _anno, pose_values, real_pose, maps, decoded_image, _map_tr = decoded
    return video.visulize(box=None if bboxes is None else np.array(bboxes, dtype=object), scene_name=name, pose=pose_values, real_pose=real_pose,
                          maps=None if maps is None else {"map": maps}, decoded_image=decoded_image, collision=None, anno_collision=None)


def run_scene(model, batch: dict, settings: RunSettings, mapdecoder=None, imagedecoder=None, device=None, video=None, batch_idx: int = 0) -> Optional[dict]:
    """One scene through the path.  Returns ``{"name", "tokens", "token_path", "decoded", "video_path"}`` (``decoded`` = the ``decode_tokens``
    7-tuple), or None for a dataset scene whose token pickle already exists (model_pl.py:214-215 skips it).  video: a
    ``umgen_b200.visualize.SceneVideo`` (or the drop-in ``Visulizer``); controlled scenes always get a video, dataset scenes when
    ``settings.generate_video`` is set or for every 100th batch (model_pl.py:188-198, 260-274)."""
    control = settings.control_test
    name = scene_name_of(batch, control)
    if control and video is not None:          # model_pl.py:140-147: the caption names the controlled object; the text grows scene by scene like there
        try:
            video.spe_text = video.spe_text + str(batch["control_object"].item())
        except Exception:
            video.spe_text = video.spe_text + "_ego"
    path = None
    if settings.token_save_path is not None:
        path = os.path.join(settings.token_save_path, name + "_tokens.pkl")
        if not control and os.path.exists(path):
            return None
    kw = inference_kwargs(batch, settings, device)
    out = model.inference(**kw)
    if settings.token_save_path is not None:
        path = postprocess.save_tokens(out, settings.token_save_path, name)
    gt = batch["dataset_token"] if control else batch
    gt_np = {m: gt[m].detach().cpu().numpy() for m in ("pose", "bbox3d") if m in gt}
    decoded = postprocess.decode_tokens(dict(out), gt_np, mapdecoder, imagedecoder)
    video_path = None
    if video is not None and (control or settings.generate_video or batch_idx % 100 == 0):
        video_path = write_scene_video(video, decoded, name)
    return {"name": name, "tokens": out, "token_path": path, "decoded": decoded, "video_path": video_path}


def run_dataset(model, scenes, settings: RunSettings, mapdecoder=None, imagedecoder=None, device=None, indices: Optional[Sequence[int]] = None,
                video=None) -> List[dict]:
    """Every scene of a ``umgen_b200.dataset.NuPlanTokenScenes`` (or ``indices`` of it: one rank's share, ``umgen_b200.dp.shard_scenes``),
    batch 1 like the reference's ``DataLoader`` (evaluate.py:196-203)."""
    results = []
    for i in (range(len(scenes)) if indices is None else indices):
        r = run_scene(model, scenes.batch(i), settings, mapdecoder, imagedecoder, device, video, batch_idx=i)
        if r is not None:
            r["index"] = i
            results.append(r)
    return results
