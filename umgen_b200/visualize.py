"""Scene video compositor: the last stage of ``evaluate.py``.

Restates what the reference's ``Visulizer`` does on the path ``UMGen_PL.generate_videos`` / ``generate_compare_videos`` takes
(reference ``tools/visulize.py``: ``visulize`` :1635-1715, ``visulize_objects`` :595-684, ``draw_box`` :813-967, ``draw_ego`` :686-782,
``draw_map`` :1150-1193, ``put_text`` :969-1078, ``concatenate_images`` :1202-1259, ``vis_pred_video`` :1607-1633,
``generate_img_and_video`` / ``create_video_from_images`` :61-75, 1080-1120; ``tools/decode_map.py:39-107`` ``add_frame_number`` /
``write_video_single``): per frame a bird's-eye canvas with the predicted agent boxes, their velocity arrows and slot ids, the ego box and its
motion arrow, the decoded map underneath, six text rows, and the decoded camera image stacked below; the frames of a scene go to an mp4.

Host code like the reference's (OpenCV rasterises the primitives; this image has the same ``cv2``) but organised differently: the geometry of
all boxes of a frame is computed in one vectorised pass (the reference loops per box), the frames go straight into the ``VideoWriter`` (the
reference writes a PNG per frame into a cache directory, reads them back and shells out to ``rm -rf``), and nothing is written besides the
video (or, with ``save_video=False``, the frame PNGs).  Frames are bit-identical to the reference's (tests/test_visualize.py: golden hashes
made by ``oracle/make_golden.py visualize`` from the reference's own class, plus a live comparison wherever the reference tree is mounted).
The pixels it composes come from the GPU (``umgen_b200/vq.py``); the values from ``umgen_b200/postprocess.py``.
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Optional, Sequence, Tuple

import cv2
import numpy as np

BEV_RANGE_M = 128.0                     # the canvas spans x, y in (-64, 64) m (visulize.py:495-500, 579-590)
EGO_LENGTH_M, EGO_WIDTH_M = 5.176, 2.297          # nuPlan ego box (visulize.py:558-560)
BACKGROUND = 128
COLOR_AGENT = (0, 255, 0)
COLOR_SMALL = (0, 165, 255)             # agents thinner than 4 px in either direction
COLOR_EGO = (0, 0, 255)
COLOR_ANNOTATION = (255, 0, 0)
COLOR_HIT = (255, 0, 255)               # agents named in a frame's collision list
COLOR_ID = (0, 255, 0)
COLOR_COND = (0, 0, 255)                # text colour of the conditioning frames
COLOR_NEW = (255, 255, 255)             # ... of the generated ones
FONT = cv2.FONT_HERSHEY_SIMPLEX
QUARTER = np.pi / 2                     # the BEV is drawn with the ego heading up: everything turns by 90 degrees (visulize.py:500, 689-691, 829-838)


@dataclasses.dataclass(frozen=True)
class BevStyle:
    """Pixel sizes of the primitives; the reference keeps two sets, one for 256-pixel canvases and one for everything else (visulize.py:459-489)."""
    width: int = 512
    height: int = 512
    font_scale: float = 0.5
    font_thickness: int = 1
    line_thickness: int = 2
    arrow_length: float = 1.5           # pixels per (m / frame) of agent velocity; the ego arrow is 4 x as long
    arrow_thickness: int = 2
    text_rows: Tuple[Tuple[int, int], ...] = ((10, 30), (10, 60), (10, 90), (10, 120), (10, 150), (10, 180))

    @classmethod
    def for_canvas(cls, width: int, height: int, arrow_length_scale: float = 1) -> "BevStyle":
        if width == 256:
            return cls(width, height, 0.3, 1, 2, 1 * arrow_length_scale, 1, ((5, 20), (5, 30), (5, 40), (5, 50), (6, 60), (7, 70)))
        return cls(width, height)

    @property
    def px_per_m(self) -> float:
        return self.width / BEV_RANGE_M


def to_uint8(image, renormalize: bool = True) -> np.ndarray:
    """Decoder output in [-1, 1] (or [0, 1] with renormalize=False) -> uint8, truncating (visulize.py:1122-1130)."""
    import torch
    t = image if isinstance(image, torch.Tensor) else torch.as_tensor(np.asarray(image))
    t = t.detach().float()
    t = torch.clamp((t + 1.0) / 2.0 if renormalize else t, min=0.0, max=1.0)
    return (t.cpu().numpy() * 255).astype(np.uint8)


def blank_canvas(style: BevStyle) -> np.ndarray:
    return np.full((style.width, style.height, 3), BACKGROUND, dtype=np.uint8)          # rows = width like the reference (square in practice)


def turn_quarter(xy: np.ndarray) -> np.ndarray:
    """[n, 2] points turned by +90 degrees with the float64 rotation matrix of pi / 2 (cos = 6.1e-17, not 0: the term survives truncation to
    pixels in exact cases, so it is kept) -- transform_box (visulize.py:784-811) without the homogeneous column."""
    c, s = np.cos(QUARTER), np.sin(QUARTER)
    pts = np.concatenate([xy, np.ones((xy.shape[0], 1))], axis=1)
    rot = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    return (rot @ pts.T).T[:, :2]


def live_slots(boxes: np.ndarray) -> np.ndarray:
    """Indices of the slots that hold an agent: <pad> slots decode to the upper end of every range (x = 64, l = 15), which is what the
    reference filters on (fliter_and_map_object, visulize.py:46-58)."""
    b = np.asarray(boxes, dtype=np.float64)
    if b.size == 0:
        return np.zeros(0, dtype=np.int64)
    gone = (b[:, 0] >= 63) | (b[:, 1] > 63)
    return np.nonzero(~gone & (b[:, 3] != 15))[0]


@dataclasses.dataclass
class BoxSprites:
    """Everything OpenCV needs to draw the agents of one frame, in draw order."""
    corners: np.ndarray        # [n, 4, 2] int pixels
    centre: np.ndarray         # [n, 2] int
    arrow_end: np.ndarray      # [n, 2] int
    id_anchor: np.ndarray      # [n, 2] int
    small: np.ndarray          # [n] bool
    ids: np.ndarray            # [n] slot index


def box_sprites(boxes: np.ndarray, style: BevStyle, keep_all: bool = False) -> BoxSprites:
    """Pixel geometry of every live agent of a frame in one pass.  boxes: [60, >= 10] float (x, y, z, l, w, h, yaw, vx, vy, vz) in metres, ego
    frame, x forward.  Canvas: ego heading up, y axis down (draw_box, visulize.py:829-905, 931-966).  keep_all: no <pad> / range filter (the
    annotation side arrives filtered, visulize.py:635-650)."""
    b = np.asarray(boxes, dtype=np.float64)
    if b.ndim != 2:                                              # a frame without annotation boxes may arrive as an empty 1-D array
        b = b.reshape(0, 10)
    ids = np.arange(b.shape[0]) if keep_all else live_slots(b)
    b = b[ids]
    n = b.shape[0]
    s = style.px_per_m
    pos = turn_quarter(b[:, 0:2])
    vel = turn_quarter(b[:, 7:9])
    heading = -(b[:, 6] + QUARTER)                              # canvas y points down: angles flip sign
    cx = pos[:, 0] * s + style.width / 2
    cy = -pos[:, 1] * s + style.height / 2
    half_l, half_w = b[:, 3] * s / 2, b[:, 4] * s / 2
    local = np.stack([np.stack([-half_l, -half_w], -1), np.stack([half_l, -half_w], -1), np.stack([half_l, half_w], -1),
                      np.stack([-half_l, half_w], -1)], axis=1)                                    # [n, 4, 2]
    cos, sin = np.cos(heading), np.sin(heading)
    rx = local[:, :, 0] * cos[:, None] + local[:, :, 1] * (-sin)[:, None]
    ry = local[:, :, 0] * sin[:, None] + local[:, :, 1] * cos[:, None]
    corners = np.stack([rx + cx[:, None], ry + cy[:, None]], axis=-1).astype(int) if n else np.zeros((0, 4, 2), dtype=int)
    centre = np.stack([cx, cy], -1).astype(int)
    arrow_end = np.stack([cx + vel[:, 0] * style.arrow_length, cy + (-vel[:, 1]) * style.arrow_length], -1).astype(int)
    id_anchor = np.stack([(cx - half_l).astype(int), (cy - half_w).astype(int) - 10], -1)
    small = (half_l * 2 < 4) | (half_w * 2 < 4)
    return BoxSprites(corners, centre, arrow_end, id_anchor, small, ids)


def draw_agents(canvas: np.ndarray, boxes: np.ndarray, style: BevStyle, with_ids: bool = True, colour=COLOR_AGENT, highlight=None,
                keep_all: bool = False) -> int:
    """Draws the live agents of one frame onto `canvas` in slot order (later slots paint over earlier ones); returns how many there were.
    highlight: ids (slot indices; row indices with keep_all) drawn in COLOR_HIT -- when a list is given, even an empty one, the thin-agent
    colour is not used (the reference's if / elif chain, visulize.py:891-905)."""
    sp = box_sprites(boxes, style, keep_all)
    base = colour
    if highlight is not None:
        highlight = set(np.asarray(highlight).reshape(-1).tolist())
    for k in range(len(sp.ids)):
        if highlight is not None:
            colour = COLOR_HIT if int(sp.ids[k]) in highlight else base
        else:
            colour = COLOR_SMALL if sp.small[k] else base
        quad = [tuple(int(v) for v in p) for p in sp.corners[k]]
        for e in range(4):
            cv2.line(canvas, quad[e], quad[(e + 1) % 4], colour, style.line_thickness)
        cv2.arrowedLine(canvas, tuple(int(v) for v in sp.centre[k]), tuple(int(v) for v in sp.arrow_end[k]), colour, style.arrow_thickness, cv2.LINE_AA)
        if with_ids:
            cv2.putText(canvas, f"{int(sp.ids[k])}", tuple(int(v) for v in sp.id_anchor[k]), FONT, style.font_scale, COLOR_ID, style.font_thickness)
    return len(sp.ids)


def ego_quad(style: BevStyle) -> np.ndarray:
    """The ego box in pixels: half sizes truncated first, turned by the float64 quarter turn, truncated again (draw_ego, visulize.py:709-753) --
    with cos(pi / 2) = 6e-17 a corner at 4 - 6e-16 truncates to 3: the quadrilateral comes out slightly skewed, and the frames are pinned to that."""
    s = style.px_per_m
    hl, hw = EGO_LENGTH_M * s / 2, EGO_WIDTH_M * s / 2
    local = np.array([[-hl, -hw], [hl, -hw], [hl, hw], [-hl, hw]]).astype(int)
    c, sn = np.cos(QUARTER), np.sin(QUARTER)
    quad = np.dot(local, np.array([[c, -sn], [sn, c]]).T).astype(int)
    quad[:, 0] += int(style.width / 2)
    quad[:, 1] += int(style.height / 2)
    return quad


def draw_ego(canvas: np.ndarray, turned_motion: np.ndarray, style: BevStyle) -> None:
    """Ego box at the canvas centre, heading up, and the frame's ego motion (dx, dy), already turned to the canvas (turn_quarter), as an
    arrow 4 x the agents' scale."""
    quad = [tuple(int(v) for v in p) for p in ego_quad(style)]
    for e in range(4):
        cv2.line(canvas, quad[e], quad[(e + 1) % 4], COLOR_EGO, style.line_thickness)
    cx, cy = style.width / 2, style.height / 2
    v = turned_motion
    length = style.arrow_length * 4
    cv2.arrowedLine(canvas, (int(cx), int(cy)), (int(cx + v[0] * length), int(cy + (-v[1]) * length)), COLOR_EGO, style.arrow_thickness, cv2.LINE_AA)


def underlay_map(canvas: np.ndarray, map_rgb: np.ndarray, style: BevStyle) -> np.ndarray:
    """The decoded map (uint8 [3, h, w]) scaled to the canvas by nearest neighbour, shown wherever nothing was drawn: a canvas pixel counts as
    drawn when any channel differs from the background grey (draw_map, visulize.py:1150-1193)."""
    m = cv2.resize(np.ascontiguousarray(map_rgb.transpose(1, 2, 0)), (style.width, style.height), interpolation=cv2.INTER_NEAREST)
    if m.shape != canvas.shape:            # non-square canvas: the map is centred on black
        pad = np.zeros_like(canvas)
        y0, x0 = (pad.shape[0] - m.shape[0]) // 2, (pad.shape[1] - m.shape[1]) // 2
        pad[y0:y0 + m.shape[0], x0:x0 + m.shape[1]] = m
        m = pad
    drawn = (canvas[..., 0] != BACKGROUND) | (canvas[..., 1] != BACKGROUND) | (canvas[..., 2] != BACKGROUND)
    return cv2.copyTo(canvas, drawn.view(np.uint8), m)


def caption(canvas: np.ndarray, rows: Sequence[Optional[str]], style: BevStyle, colour) -> np.ndarray:
    out = canvas.copy()
    for text, at in zip(rows, style.text_rows):
        if text is not None:
            cv2.putText(out, text, at, FONT, style.font_scale, colour, style.font_thickness)
    return out


def stack_rows(layers: Sequence[Sequence[np.ndarray]]) -> List[np.ndarray]:
    """Frame i = the i-th image of every layer, top to bottom on black, left-aligned; a layer that is shorter than the first one repeats its last
    image (concatenate_images(mode="vertical"), visulize.py:1202-1259)."""
    n = len(layers[0])
    width = max(img.shape[1] for layer in layers for img in layer)
    height = sum(layer[0].shape[0] for layer in layers)
    frames = []
    for i in range(n):
        f = np.zeros((height, width, 3), dtype=np.uint8)
        y = 0
        for layer in layers:
            img = layer[i] if i < len(layer) else layer[-1]
            f[y:y + img.shape[0], :img.shape[1]] = img
            y += img.shape[0]
        frames.append(f)
    return frames


def write_mp4(frames: Sequence[np.ndarray], path: str, fps: int = 5) -> str:
    """mp4v at 5 frames/s (create_video_from_images, visulize.py:61-75), frames written as they are (the reference never swaps channels either)."""
    if len(frames) == 0:
        raise ValueError("no frames to write")
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    h, w = frames[0].shape[:2]
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
    if not vw.isOpened():
        raise RuntimeError(f"cannot open a video writer for {path}")
    try:
        for f in frames:
            vw.write(np.ascontiguousarray(f))
    finally:
        vw.release()
    return path


class SceneVideo:
    """Composes and writes the videos of a scene.  Constructor arguments and attribute names follow the reference's ``Visulizer`` where
    ``model_pl.py`` touches them (model_pl.py:61-73, 142-147: ``spe_text`` is reassigned per scene)."""

    def __init__(self, video_save_path="output/videos/", video_pretext="test", width=256, height=256, project_name="test", spe_text="p=0.5",
                 save_video=True, addtion_ego=False, bbox3d_arrow_length_scale=1, cond_frames=20, put_text=True, frame_dir: Optional[str] = None):
        self.video_save_base_path = video_save_path
        self.video_pretext = video_pretext
        self.project_name = project_name
        self.spe_text = spe_text
        self.save_video = save_video
        self.addtion_ego = addtion_ego
        self.cond_frames = cond_frames
        self.put_text_on_img = put_text
        self.style = BevStyle.for_canvas(width, height, bbox3d_arrow_length_scale)
        self.frame_dir = frame_dir or os.path.join("output/tmp_cache", project_name)          # only used with save_video=False

    # ---- frames ------------------------------------------------------------------------------------------------------------------------
    def compose(self, boxes=None, pose=None, real_pose=None, map_images=None, decoded_image=None, scene_name: str = "0", anno_boxes=None,
                collision=None, anno_collision=None) -> List[np.ndarray]:
        """The frames ``visulize`` would write (visulize.py:1635-1715; generate_videos passes the first five, model_pl.py:283-314):
        boxes: T arrays [60, 10] (postprocess.decode_bbox3d); pose / real_pose: [T, 3] / [T', 3] metres and radians; map_images: [T, 3, h, w]
        in [-1, 1]; decoded_image: [T, 3, H, W] in [-1, 1]; anno_boxes: T arrays [n_t, 10] (postprocess.decode_annotation_bbox3d), drawn first,
        unfiltered, without ids; collision / anno_collision: per frame the slot ids / annotation rows to highlight."""
        st = self.style
        n = len(boxes) if boxes is not None else (len(anno_boxes) if anno_boxes is not None else (len(pose) if pose is not None else 0))
        if n == 0:
            raise ValueError("nothing to draw: neither boxes nor poses")
        if (boxes is not None or anno_boxes is not None) and not self.addtion_ego:
            raise NotImplementedError("addtion_ego=False (slot 0 drawn as the ego) is not a path evaluate.py takes")
        if anno_boxes is not None and len(anno_boxes) < n:
            raise ValueError(f"{len(anno_boxes)} frames of annotation boxes for {n} frames")
        canvases = [blank_canvas(st) for _ in range(n)]
        counts, anno_counts = [0] * n, [0] * n
        for i in range(n):
            if anno_boxes is not None:
                anno_counts[i] = draw_agents(canvases[i], np.array(anno_boxes[i], dtype=np.float64), st, with_ids=False, colour=COLOR_ANNOTATION,
                                             highlight=None if anno_collision is None else anno_collision[i], keep_all=True)
            if boxes is not None:
                counts[i] = draw_agents(canvases[i], np.array(boxes[i], dtype=np.float64), st, with_ids=self.put_text_on_img,
                                        highlight=None if collision is None else collision[i])
        if pose is not None and self.addtion_ego:
            turned = turn_quarter(np.asarray(pose, dtype=np.float64)[:, 0:2])
            for i in range(n):
                draw_ego(canvases[i], turned[i], st)
        if map_images is not None:
            maps8 = to_uint8(map_images)
            canvases = [underlay_map(canvases[i], maps8[i], st) for i in range(len(maps8))]
        if self.put_text_on_img:
            canvases = [self._caption(c, i, counts[i], anno_counts[i], scene_name, pose, real_pose) for i, c in enumerate(canvases)]
        layers = [canvases]
        if decoded_image is not None:
            layers.append(list(to_uint8(decoded_image).transpose(0, 2, 3, 1)))
        return stack_rows(layers)

    def _caption(self, canvas, i, n_pred, n_anno, scene_name, pose, real_pose):
        rows = [f"Frame {i}: pbox={n_pred}, abox={n_anno}", f"Project: {self.project_name}",
                f"{self.spe_text}" if self.spe_text is not None else None,
                f"Scene: {scene_name}" if scene_name is not None else None, None, None]
        if pose is not None:
            p = np.round(np.asarray(pose), 2)
            rows[4] = f"Pose: ({p[i][0]:.2f}, {p[i][1]:.2f}, {p[i][2]:.2f})"
        if real_pose is not None:
            r = np.round(np.asarray(real_pose), 2)
            rows[5] = "GTPose: out of annotaion" if i >= len(r) else f"GTPose: ({r[i][0]:.2f}, {r[i][1]:.2f}, {r[i][2]:.2f})"
        return caption(canvas, rows, self.style, COLOR_COND if i < self.cond_frames else COLOR_NEW)

    # ---- files -------------------------------------------------------------------------------------------------------------------------
    def _emit(self, frames, scene_name, base=None) -> str:
        if self.save_video:
            return write_mp4(frames, os.path.join(base or self.video_save_base_path, f"{self.video_pretext}_{scene_name}.mp4"))
        d = os.path.join(self.frame_dir, scene_name)
        os.makedirs(d, exist_ok=True)
        for i, f in enumerate(frames):
            cv2.imwrite(os.path.join(d, f"{i}.png"), f)
        return d

    def visulize(self, box=None, anno_box=None, collision=None, anno_collision=None, test_object=0, set_index=None, scene_name="0", maps=None,
                 pose=None, real_pose=None, decoded_image=None, view_mask=None) -> str:
        """Same call as ``Visulizer.visulize``.  Of ``maps`` the "map" entry is drawn (generate_videos never produces another: "map_trans"
        needs tokens ``inference`` does not return, "map_tokens" is a debugging view); ``test_object`` / ``view_mask`` are unused there too."""
        extra = set(maps or {}) - {"map"}
        if extra:
            raise NotImplementedError(f"map layers {sorted(extra)} are not on evaluate.py's path")
        if set_index is not None:
            raise NotImplementedError("set_index (all frames written to one file name) is not on evaluate.py's path")
        frames = self.compose(box, pose, real_pose, None if not maps else maps.get("map"), decoded_image, scene_name, anno_box, collision, anno_collision)
        return self._emit(frames, scene_name)

    def vis_pred_video(self, decoded_image, scene_name, video_type="pred", renormalize=True) -> str:
        """The decoded camera frames alone, into ``<video dir>_<video_type>/`` (visulize.py:1607-1633)."""
        frames = list(to_uint8(decoded_image, renormalize).transpose(0, 2, 3, 1))
        base = self.video_save_base_path
        name = os.path.basename(base) or os.path.basename(base[:-1])
        return self._emit(frames, scene_name, base.replace(name, f"{name}_{video_type}"))


# ---- tools/decode_map.py's small video helper --------------------------------------------------------------------------------------------
def frame_label(image: np.ndarray, frame_idx: int, pose_value=None, font_scale: float = 0.2, cond_num: int = 20) -> np.ndarray:
    """add_frame_number (decode_map.py:39-77): frame index (and the pose truncated to two decimals) in the top-left corner, green for the
    conditioning frames, red afterwards."""
    colour = (0, 255, 0) if frame_idx < cond_num else (0, 0, 255)
    if pose_value is not None:
        pv = np.trunc(np.asarray(pose_value) * 10 ** 2) / (10 ** 2)
        text = f"F: {frame_idx}   [dx, dy, dh]: {pv}"
    else:
        text = f"Frame: {frame_idx}"
    (_, th), _ = cv2.getTextSize(text, FONT, font_scale, 1)
    return cv2.putText(image.copy(), text, (10, 10 + th), FONT, font_scale, colour, 1)


def write_video_single(images, pose_values=None, save_path=None, cond_num=20, font_scale=0.5, fps=10, h=256, w=256) -> None:
    """write_video_single (decode_map.py:80-107): [T, 3, h, w] decoder output -> labelled mp4."""
    frames = to_uint8(images).transpose(0, 2, 3, 1)
    out = cv2.VideoWriter(save_path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
    for i, f in enumerate(frames):
        out.write(frame_label(f, i, None if pose_values is None else pose_values[i], font_scale, cond_num))
    out.release()
