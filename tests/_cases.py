"""Seeded test-case definitions shared by oracle/make_golden.py (which runs the real reference on
them) and the parity tests (which replay the committed answers)."""
from __future__ import annotations

import numpy as np

EGO_BOX = np.array([0, 0, 0, 5.176, 2.297, 1.777, 0, 0, 0, 0], dtype=np.float64)

_RANGES = np.array([(-64, 64), (-64, 64), (-5, 5), (0, 15), (0, 4), (0, 5), (-3.14, 3.14),
                    (-20, 20), (-15, 15), (-0.3, 0.3)], dtype=np.float64)


def _token_box(rs, spread):
    """A box whose attributes sit on the 1024-bin token grid, like the decode loop produces."""
    bins = np.linspace(0.0, 1.0, 1024)
    tok = rs.randint(0, 1028, size=10)
    tok[0] = np.clip(512 + rs.randint(-spread, spread + 1), 0, 1027)
    tok[1] = np.clip(512 + rs.randint(-spread, spread + 1), 0, 1027)
    right = np.clip(tok, 0, 1023)
    left = np.clip(tok - 1, 0, 1023)
    mid = (bins[left] + bins[right]) / 2
    return mid * (_RANGES[:, 1] - _RANGES[:, 0]) + _RANGES[:, 0]


def collision_cases():
    """List of box lists (each box: x,y,z,l,w,h,yaw,vx,vy,vz); the question asked of each list is
    BoxOverlap.check_collision(list, fliter=True) (reference plugin/misc/misc.py:591-630)."""
    cases = []
    e = EGO_BOX

    def box(x, y, l=4.5, w=2.0, yaw=0.0):
        return np.array([x, y, 0, l, w, 1.5, yaw, 0, 0, 0], dtype=np.float64)

    # known-answer cases recorded in SURVEY.md section 8c
    cases += [[e, box(20, 0)], [e, box(2, 0.5)], [e, box(0, 0, 1, .5)], [e, box(0, 0, 12, 3.9)],
              [e, box(4, 0, yaw=.7)], [e], [e, box(70, 0)], [e, box(63, 0)], [e, box(62.99, 0)],
              [box(70, 0), box(70.5, 0)], [e, box(2, 0.5), box(70, 0)], [e, box(30, 30), box(30, 30)]]
    rs = np.random.RandomState(20260117)
    for i in range(700):
        n = rs.randint(1, 12)
        spread = [20, 60, 200, 500][i % 4]
        cases.append([e] + [_token_box(rs, spread) for _ in range(n)])
    for i in range(200):          # continuous boxes, tight cluster, arbitrary yaw
        n = rs.randint(1, 6)
        lst = [e]
        for _ in range(n):
            lst.append(np.array([rs.uniform(-12, 12), rs.uniform(-8, 8), 0, rs.uniform(.3, 9), rs.uniform(.3, 3.5),
                                 1.5, rs.uniform(-3.14, 3.14), 0, 0, 0]))
        cases.append(lst)
    return cases


# tiny-depth rollouts executed by the real reference in oracle/make_golden.py
ROLLOUT_CASES = {
    # free video rollout: 2 conditioning frames, window of 2, 2 new frames (window slides once)
    "video_L1": dict(layers=1, weight_seed=0, scene_seed=1, input_frames=4, input_cond_frames=2,
                     cond_frames=2, new_frames=2),
    # control rollout: forced ego poses + one forced agent slot
    "control_L1": dict(layers=1, weight_seed=3, scene_seed=2, input_frames=3, input_cond_frames=2,
                       cond_frames=3, new_frames=2, control=True),
    # deeper stacks, one frame
    "video_L2": dict(layers=2, weight_seed=5, scene_seed=7, input_frames=3, input_cond_frames=3,
                     cond_frames=3, new_frames=1),
    # ground-truth pose + map given as init tokens (tools/model_pl.py:103-115 with init_token_mod = ["pose", "map"]): they are the forced
    # prefix of the OAR decode (UMGen.py:1184-1201) and replace the generated rows in the output
    "initmap_L1": dict(layers=1, weight_seed=4, scene_seed=3, input_frames=4, input_cond_frames=2,
                       cond_frames=2, new_frames=2, init_mods=("pose", "map")),
    # the headline window: 20 conditioning frames, depth 4 in every stack, 2 new frames (the window slides once, so the
    # second frame runs the look-ahead schedule's last-frame path at T = 20)
    "video_T20_L4": dict(layers=4, weight_seed=21, scene_seed=9, input_frames=20, input_cond_frames=20,
                         cond_frames=20, new_frames=2),
    # control working point: 13 conditioning frames, window growing 13 -> 15 inside a 20-frame limit, depth 2
    "control_T13_L2": dict(layers=2, weight_seed=22, scene_seed=10, input_frames=13, input_cond_frames=13,
                           cond_frames=20, new_frames=3, control=True),
}


# decode-only cases: the reference's infer_oar_net driven with a seeded random conditioning feature
OAR_CASES = {
    "oar_L2": dict(oar_layers=2, weight_seed=11, scene_seed=4, feat_seed=21, control_slot=None),
    "oar_L1_control": dict(oar_layers=1, weight_seed=12, scene_seed=5, feat_seed=22, control_slot=3),
    # AR bbox head biased towards <pad>: exercises the TAR-head resample of UMGen.py:1092-1104
    "oar_L1_padheavy": dict(oar_layers=1, weight_seed=13, scene_seed=6, feat_seed=23, control_slot=None,
                            tweak="padheavy"),
    # AR bbox head biased towards the middle bins: new-born boxes land within ~16 m of the ego box with mid-range sizes, so the
    # rotated-box collision test (not the 30-box rule) decides which slots are wiped (UMGen.py:1336-1377)
    "oar_L1_collide": dict(oar_layers=1, weight_seed=14, scene_seed=8, feat_seed=24, control_slot=None,
                           tweak="collide"),
}


def apply_tweak(sd, name):
    """In-place edits of a synthetic state_dict that steer a rollout into rarely taken branches."""
    if name is None:
        return sd
    if name == "padheavy":
        w = sd["transformer.head_ar_bbox3d.weight"]
        w[:1027] = (w[:1027] / 64).half().float()
        return sd
    if name == "collide":
        w = sd["transformer.head_ar_bbox3d.weight"]
        w[384:640] = (w[384:640] * 8).half().float()
        return sd
    raise ValueError(name)



def rollout_init(spec, scene):
    """init_tokens of a ROLLOUT_CASES entry (None for a free rollout)."""
    from umgen_b200 import synth
    if spec.get("control"):
        return synth.make_control(seed=spec["scene_seed"], n_frames=spec["new_frames"])
    if spec.get("init_mods"):
        n_in = spec["input_cond_frames"]
        return {m: scene[m][:, n_in:n_in + spec["new_frames"]].clone() for m in spec["init_mods"]}
    return None


def oar_inputs(spec):
    """(tar_feat [2207,768] fp32, pose_tok [3], prev_bbox [660]) for an OAR_CASES entry."""
    import torch
    from umgen_b200 import synth
    g = torch.Generator().manual_seed(spec["feat_seed"])
    tar_feat = torch.randn(2207, 768, generator=g)
    scene = synth.make_scene(seed=spec["scene_seed"], n_frames=1)
    pose = torch.randint(0, 1024, (3,), generator=g)
    return tar_feat, pose, scene["bbox3d"][0, 0].clone()


def vq_codes(kind: str):
    """Seeded token grids for the VQ decoder cases: 2 frames of [32,32] (map) or [16,32] (image)."""
    import torch
    h, w = (32, 32) if kind == "map" else (16, 32)
    return torch.randint(0, 8192, (2, h, w), generator=torch.Generator().manual_seed(77 if kind == "map" else 78))


# ---- raw tokenised nuPlan scene pickles for the dataset front-end -------------------------------------------------------
DATASET_CASES = {      # name -> (seed, frames in the scene, block_size, sampling_gap, tracks in the scene)
    "video_50": (3, 240, 50, 4, 90),      # evaluate.py --infer_task video: 20 conditioning + 30 new frames, gap 4
    "short_clip": (4, 70, 50, 4, 90),     # too short for the block: get_frame_indices shortens the clip
    "dense_gap1": (5, 64, 33, 1, 260),    # more than 60 tracks inside the clip: the slot table overflows
}


def raw_scene(seed: int, n: int, n_tracks: int = 90):
    """A raw scene dict with the corner cases of the reference's dataset code: objects that appear and vanish, more than 60 tracks per clip, track id
    0, frames without objects, one-object frames, categories outside the vocabulary, boxes beyond 64 m and beyond the normalisation ranges, 12-column
    boxes, headings that wrap through +-pi."""
    rs = np.random.RandomState(seed)
    cats_pool = ["vehicle", "bicycle", "pedestrian", "traffic_cone", "barrier", "czone_sign"]
    t_cat = [cats_pool[i] for i in rs.choice(len(cats_pool), n_tracks, p=[0.5, 0.12, 0.2, 0.08, 0.05, 0.05])]
    t_id = rs.permutation(np.arange(0, 1000))[:n_tracks]
    t_id[7] = 0                                               # a track whose id is "falsy"
    birth = rs.randint(-20, n, n_tracks)
    life = rs.randint(5, n, n_tracks)
    base = np.stack([rs.uniform(-75, 75, n_tracks), rs.uniform(-70, 70, n_tracks), rs.uniform(-6, 6, n_tracks), rs.uniform(0.3, 17, n_tracks),
                     rs.uniform(0.3, 4.5, n_tracks), rs.uniform(0.5, 5.5, n_tracks), rs.uniform(-3.3, 3.3, n_tracks), rs.uniform(-22, 22, n_tracks),
                     rs.uniform(-16, 16, n_tracks), rs.uniform(-0.4, 0.4, n_tracks), rs.uniform(-1, 1, n_tracks), rs.uniform(-1, 1, n_tracks)], axis=1)
    heading = 3.0 + np.cumsum(rs.uniform(-0.02, 0.06, n))     # crosses pi
    heading = (heading + np.pi) % (2 * np.pi) - np.pi
    xy = np.cumsum(np.stack([rs.uniform(0.2, 1.2, n), rs.uniform(-0.1, 0.1, n)], axis=1), axis=0)
    meta = []
    for i in range(n):
        T = np.eye(4)
        c, s = np.cos(heading[i]), np.sin(heading[i])
        T[:2, :2] = [[c, -s], [s, c]]
        T[:2, 3] = xy[i]
        T[2, 3] = 0.01 * i
        alive = np.nonzero((birth <= i) & (i < birth + life))[0]
        if i % 37 == 11:
            alive = alive[:0]                                 # a frame without objects
        elif i % 41 == 14:
            alive = alive[:1]                                 # a frame with one object
        elif i % 53 == 30:
            alive = np.array([7])                             # a frame whose only object has track id 0
        boxes = base[alive].copy()
        boxes[:, 0] += 0.04 * i * boxes[:, 7] / 5
        boxes[:, 1] += 0.04 * i * boxes[:, 8] / 5
        order = rs.permutation(len(alive))
        meta.append({"T_lidar2global": T, "bboxes_3d": boxes[order].astype(np.float32), "track_ids": t_id[alive][order].copy(),
                     "categories": [t_cat[j] for j in alive[order]]})
    ego = np.zeros((n, 16))
    ego[:, 6] = heading
    return {
        "tokens": {"CAM_F0": {"tokens": [rs.randint(0, 8192, (16, 32)) for _ in range(n)], "file_list": [f"{i:06d}.jpg" for i in range(n)]}},
        "raster_tokens": rs.randint(0, 8192, (n, 32, 32)),
        "ego_pose_all": ego,
        "meta_info": meta,
        "lidar_bboxes": {"CAM_F0": {"bboxes_3d": [m["bboxes_3d"] for m in meta], "categories": [m["categories"] for m in meta],
                                    "track_ids": [m["track_ids"] for m in meta]}},
    }


# ---- batches for the scene runner (umgen_b200/runner.py vs the reference's UMGen_PL.world_model_evaluate) ------------------------------------
RUNNER_CASES = {      # name -> (infer_task, new_frames, init_token_mod, scene name of the control pickle, with input_cond_frame)
    "video_3": ("video", 3, None, None, False),
    "video_to_end": ("video", -1, None, None, False),
    "video_init_pose_map": ("video", 2, ["pose", "map"], None, False),
    "control": ("control", 30, None, "ctrl_scene_0007_obj", True),
    "control_off": ("control", 30, None, "ctrl_scene_0007_no_control", False),
}


def runner_batch(name: str):
    """The batch a DataLoader(batch_size=1) would hand over for a runner case: a dataset scene (free rollout) or a controlled_scenes pickle."""
    import torch
    from torch.utils.data import default_collate
    task, _, _, scene_name, with_icf = RUNNER_CASES[name]
    g = torch.Generator().manual_seed(17)
    T = 26
    toks = {"pose": torch.randint(0, 1024, (T, 3), generator=g), "map": torch.randint(0, 8192, (T, 1024), generator=g),
            "pose_diff": torch.randn(T, 3, generator=g), "bbox3d": torch.randint(0, 1028, (T, 660), generator=g),
            "image": torch.randint(0, 8192, (T, 512), generator=g)}
    if task == "video":
        return default_collate([dict(toks, file_name="0_/data/tokenized_origin_scenes/synthetic_scene_0003_clip_a.pkl")])
    ctrl = {"pose": torch.randint(0, 1024, (30, 3), generator=g), "bbox3d": torch.full((30, 660), -1, dtype=torch.int64)}
    ctrl["bbox3d"][:, 22:33] = torch.randint(0, 1027, (30, 11), generator=g)
    item = {"dataset_token": {m: toks[m][:13] for m in ("pose", "map", "bbox3d", "image")}, "control_dict": ctrl, "scene_name": scene_name,
            "control_object": 2}
    if with_icf:
        item["input_cond_frame"] = 13
    return default_collate([item])


def summarise_inference_kwargs(kw: dict) -> dict:
    """A JSON-able fingerprint of the keyword arguments handed to UMGen.inference: scalars as they are, token dicts as {key: [shape, sha256]}."""
    import hashlib
    import torch

    def fp(v):
        if torch.is_tensor(v):
            return [list(v.shape), hashlib.sha256(v.detach().cpu().contiguous().to(torch.float64).numpy().tobytes()).hexdigest()[:16]]
        if isinstance(v, dict):
            return {k: fp(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return [fp(x) for x in v]
        return v
    return {k: fp(v) for k, v in sorted(kw.items())}


VISUALIZE_CASES = {     # name -> (seed, frames, canvas width, cond_frames, put_text, map side, image (h, w), annotated poses)
    "bev512": (21, 7, 512, 3, True, 256, (256, 512), 5),        # what tools/model_pl.py builds; real_pose shorter than the rollout
    "bev256_notext": (22, 4, 256, 2, False, 64, (64, 128), 0),     # the 256-pixel style set, no captions / ids
    "bev512_crowded": (23, 5, 512, 20, True, 128, (128, 256), 5),   # every slot alive, agents near and over the canvas border, tiny agents
    "bev512_annotated": (24, 4, 512, 2, True, 64, (64, 128), 4),    # annotation boxes underneath, collision highlights on both sides (not on evaluate.py's path)
}


def visualize_inputs(name: str):
    """Seeded inputs of the scene compositor (umgen_b200/visualize.py), shaped like what ``UMGen_PL.decode_tokens`` hands to ``generate_videos``:
    boxes = T arrays [60, 10] decoded from bbox3d tokens (pad slots included), pose / real_pose in metres and radians, map and camera pixels in
    about [-1.2, 1.2] as float tensors."""
    import torch
    from umgen_b200 import postprocess, synth
    seed, T, width, cond, put_text, map_side, (ih, iw), n_real = VISUALIZE_CASES[name]
    scene = synth.make_scene(seed=seed, n_frames=T)
    bt = scene["bbox3d"][0, :T].numpy().astype(np.int64).copy()
    rs = np.random.RandomState(seed)
    if name.endswith("crowded"):
        slots = bt.reshape(T, 60, 11)
        slots[:, :, :10] = rs.randint(0, 1024, size=(T, 60, 10))
        slots[:, :, 10] = 1024 + rs.randint(0, 3, size=(T, 60))
        slots[:, ::7, 0] = rs.randint(1000, 1028, size=slots[:, ::7, 0].shape)          # x at / beyond the filter threshold (63 m) and <pad>
        slots[:, 1::7, 3:5] = rs.randint(0, 60, size=slots[:, 1::7, 3:5].shape)         # thinner than 4 px
        slots[:, 2::9, 0:2] = 512                                                       # on top of the ego
        slots[:, 3::9, 7:9] = 512                                                       # standing still: zero-length arrow
    boxes, _ = postprocess.decode_bbox3d(bt)
    pose = postprocess.decode_pose(scene["pose"][0, :T].numpy())
    real = postprocess.decode_pose(synth.make_scene(seed=seed + 100, n_frames=T)["pose"][0, :n_real].numpy()) if n_real else None
    g = torch.Generator().manual_seed(seed)
    maps = torch.randn(T, 3, map_side, map_side, generator=g) * 0.6
    maps[:, :, ::5, :] = -1.0 + 128 / 255 * 2 + 1e-3             # rows that decode to the background grey (128): they read as "nothing drawn" nowhere, the canvas mask looks at the canvas
    image = torch.randn(T, 3, ih, iw, generator=g) * 0.6
    extra = {}
    if name.endswith("annotated"):
        gt = synth.make_scene(seed=seed + 200, n_frames=T)["bbox3d"][0, :T].numpy().astype(np.int64)
        anno, _ = postprocess.decode_annotation_bbox3d(gt)
        extra = dict(anno_boxes=anno, collision=[sorted(rs.choice(60, size=6, replace=False).tolist()) for _ in range(T)],
                     anno_collision=[sorted(rs.choice(max(len(a), 1), size=min(2, len(a)), replace=False).tolist()) for a in anno])
    return dict(boxes=boxes, pose=pose, real_pose=real, maps=maps, image=image, width=width, cond_frames=cond, put_text=put_text,
                scene_name=f"synthetic_{name}", **extra)


SAMPLER_CASES = [      # (name, method, parameter, vocabulary, rows, logit scale, seed)
    ("topk5_bbox", "topk", 5, 1028, 64, 3.0, 1), ("topk5_map", "topk", 5, 8192, 48, 4.0, 2), ("topk16_image", "topk", 16, 8192, 48, 2.0, 3),
    ("topk5_ego", "topk", 5, 1024, 3, 5.0, 4), ("topk_more_than_vocab", "topk", 2000, 1028, 8, 1.0, 5),
    ("topp0.4_bbox", "topp", 0.4, 1028, 64, 3.0, 6), ("topp0.4_map", "topp", 0.4, 8192, 48, 4.0, 7), ("topp0.9_flat", "topp", 0.9, 8192, 32, 0.5, 8),
    ("topp16_image_quirk", "topp", 16.0, 8192, 32, 2.0, 9),        # UMGen.py:1133 hands topk_image (16) to sample_top_p as p: the whole vocabulary stays
    ("topp_tiny", "topp", 1e-6, 1028, 32, 3.0, 10), ("topp0.4_ego", "topp", 0.4, 1024, 3, 5.0, 11),
]


def sampler_logits(vocab: int, rows: int, scale: float, seed: int):
    """Seeded logits [rows, vocab] with a few exact ties in every row (the truncation thresholds are inclusive / exclusive in different places)."""
    import torch
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.randn(rows, vocab, generator=g) * scale
    top = torch.topk(x, 6, dim=-1).indices
    x[torch.arange(rows), top[:, 4]] = x[torch.arange(rows), top[:, 5]]            # 5th and 6th largest equal
    return x
