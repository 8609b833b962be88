"""Scene compositor (umgen_b200/visualize.py) against the reference's Visulizer: the golden hashes were made by the
reference's own class (oracle/make_golden.py visualize); where the reference tree is mounted the frames are also compared pixel by pixel, live."""
import hashlib
import json
import os
import sys

import cv2
import numpy as np
import pytest
import torch

from tests._cases import VISUALIZE_CASES, visualize_inputs
from umgen_b200 import visualize as V

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "visualize.json")))
same_cv2 = pytest.mark.skipif(cv2.__version__ != GOLD["cv2"], reason=f"golden hashes were rasterised by OpenCV {GOLD['cv2']}")


def make(d, tmp, **kw):
    return V.SceneVideo(video_save_path=os.path.join(str(tmp), "clips/"), video_pretext="UMGen", width=d["width"], height=d["width"],
                        project_name="UMGen_infer", spe_text="synthetic_video", save_video=True, addtion_ego=True, bbox3d_arrow_length_scale=1,
                        cond_frames=d["cond_frames"], put_text=d["put_text"], **kw)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@same_cv2
@pytest.mark.parametrize("name", list(VISUALIZE_CASES))
def test_frames_and_videos_match_the_reference_golden(name, tmp_path):
    d, g = visualize_inputs(name), GOLD["cases"][name]
    vis = make(d, tmp_path)
    frames = vis.compose(d["boxes"], d["pose"], d["real_pose"], d["maps"], d["image"], d["scene_name"], d.get("anno_boxes"), d.get("collision"),
                         d.get("anno_collision"))
    assert list(frames[0].shape) == g["shape"] and len(frames) == len(g["frames"])
    assert [sha(f) for f in frames] == g["frames"]
    # the same through the two calls UMGen_PL makes (model_pl.py:305-331), down to the bytes of the mp4 files
    path = vis.visulize(box=np.array(d["boxes"], dtype=object), anno_box=d.get("anno_boxes"), scene_name=d["scene_name"], pose=d["pose"],
                        real_pose=d["real_pose"], maps={"map": d["maps"]}, decoded_image=d["image"], collision=d.get("collision"),
                        anno_collision=d.get("anno_collision"))
    assert path == os.path.join(str(tmp_path), "clips/", f"UMGen_{d['scene_name']}.mp4")
    pred = vis.vis_pred_video(d["image"], d["scene_name"], video_type="pred")
    assert pred == os.path.join(str(tmp_path), "clips_pred/", f"UMGen_{d['scene_name']}.mp4")
    assert hashlib.sha256(open(path, "rb").read()).hexdigest() == g["mp4_sha256"]
    assert hashlib.sha256(open(pred, "rb").read()).hexdigest() == g["pred_mp4_sha256"]
    cap = cv2.VideoCapture(path)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == len(frames) and cap.get(cv2.CAP_PROP_FPS) == 5
    ok, first = cap.read()
    assert ok and first.shape == frames[0].shape
    assert np.abs(first.astype(int) - frames[0].astype(int)).mean() < 40          # lossy codec on noise pixels: the picture, not the bits (unrelated frames differ by ~85)
    pf = list(V.to_uint8(d["image"]).transpose(0, 2, 3, 1))
    assert [sha(f) for f in pf] == g["pred_frames"] and list(pf[0].shape) == g["pred_shape"]


def test_inputs_are_left_alone_and_frames_do_not_depend_on_call_order(tmp_path):
    """The reference turns the box arrays in place (a second call on the same arrays would draw them turned twice); here inputs are read-only."""
    d = visualize_inputs("bev512_crowded")
    keep = [b.copy() for b in d["boxes"]], d["pose"].copy(), d["maps"].clone()
    vis = make(d, tmp_path)
    a = vis.compose(d["boxes"], d["pose"], d["real_pose"], d["maps"], d["image"], d["scene_name"])
    b = vis.compose(d["boxes"], d["pose"], d["real_pose"], d["maps"], d["image"], d["scene_name"])
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert all(np.array_equal(x, y) for x, y in zip(keep[0], d["boxes"])) and np.array_equal(keep[1], d["pose"]) and torch.equal(keep[2], d["maps"])


def test_layers():
    st = V.BevStyle.for_canvas(512, 512)
    assert st.px_per_m == 4 and V.BevStyle.for_canvas(256, 256, 2).arrow_length == 2 and V.BevStyle.for_canvas(256, 256).font_scale == 0.3
    # <pad> slots decode to the upper end of every range and are skipped; so is anything beyond 63 m
    b = np.zeros((5, 10))
    b[:, 3] = 4.0
    b[1, 0], b[2, 1], b[3, 3], b[4, 1] = 63.0, 63.5, 15.0, 63.0
    assert V.live_slots(b).tolist() == [0, 4] and V.live_slots(np.zeros((0, 10))).size == 0
    # ego box: both truncations of the reference (half sizes, then the corners turned with cos(pi / 2) = 6e-17: 4 - 6e-16 truncates to 3) -- a
    # slightly skewed quadrilateral, which the reference's frames show and the golden frames pin
    q = V.ego_quad(st)
    assert q.tolist() == [[259, 246], [260, 266], [253, 266], [252, 246]]
    # nothing drawn -> the map shows everywhere; a drawn pixel wins even where the map is background grey
    canvas = V.blank_canvas(st)
    canvas[10, 10] = (0, 255, 0)
    m = np.full((3, 256, 256), 7, dtype=np.uint8)
    out = V.underlay_map(canvas, m, st)
    assert out[10, 10].tolist() == [0, 255, 0] and out[11, 10].tolist() == [7, 7, 7] and out.shape == (512, 512, 3)
    # vertical stack: left-aligned on black, a shorter layer repeats its last image
    top = [np.full((4, 6, 3), 1, np.uint8)] * 3
    low = [np.full((2, 3, 3), 2, np.uint8), np.full((2, 3, 3), 3, np.uint8)]
    fr = V.stack_rows([top, low])
    assert len(fr) == 3 and fr[0].shape == (6, 6, 3) and fr[2][4, 0, 0] == 3 and fr[2][4, 4, 0] == 0 and fr[1][0, 5, 0] == 1
    assert V.to_uint8(torch.tensor([-2.0, -1.0, 0.0, 0.999, 1.0, 3.0])).tolist() == [0, 0, 127, 254, 255, 255]
    assert V.to_uint8(np.array([0.5, 2.0]), renormalize=False).tolist() == [127, 255]
    with pytest.raises(ValueError):
        V.write_mp4([], "x.mp4")


def test_frames_without_video_and_without_boxes(tmp_path):
    d = visualize_inputs("bev256_notext")
    vis = V.SceneVideo(width=256, height=256, project_name="P", save_video=False, addtion_ego=True, put_text=True, cond_frames=1,
                       frame_dir=os.path.join(str(tmp_path), "frames"))
    out = vis.visulize(box=None, scene_name="s", pose=d["pose"], maps=None, decoded_image=None)
    assert sorted(os.listdir(out)) == [f"{i}.png" for i in range(len(d["pose"]))]
    img = cv2.imread(os.path.join(out, "0.png"))
    assert img.shape == (256, 256, 3) and (img != 128).any()                   # ego box, arrow and captions on the grey canvas
    with pytest.raises(NotImplementedError):
        vis.visulize(box=d["boxes"], scene_name="s", pose=d["pose"], maps={"map": d["maps"], "map_tokens": d["maps"]})
    with pytest.raises(ValueError):
        vis.visulize(box=d["boxes"], anno_box=d["boxes"][:1], scene_name="s", pose=d["pose"])
    with pytest.raises(ValueError):
        vis.visulize(box=None, pose=None)


def test_dropin_visulizer_takes_what_model_pl_passes(tmp_path):
    """tools/model_pl.py:61-73 builds the visualiser with these keywords and later reassigns ``spe_text`` (:142-147)."""
    sys.path.insert(0, ROOT)
    for k in [k for k in sys.modules if k == "projects" or k.startswith("projects.")]:
        del sys.modules[k]
    from projects.tools.visulize import Visulizer, add_frame_number, write_video_single
    vis = Visulizer(video_save_path=os.path.join(str(tmp_path), "v/"), video_pretext="UMGen", width=512, height=512, project_name="UMGen_infer",
                    spe_text="x_video", save_video=True, addtion_ego=True, bbox3d_arrow_length_scale=1, cond_frames=20, put_text=True)
    vis.spe_text = vis.spe_text + "_ego"
    d = visualize_inputs("bev512")
    p = vis.visulize(box=np.array(d["boxes"], dtype=object), scene_name="s", pose=d["pose"], real_pose=d["real_pose"], maps={"map": d["maps"]},
                     decoded_image=d["image"], collision=None, anno_collision=None)
    assert os.path.getsize(p) > 10000
    with pytest.raises(NotImplementedError):
        Visulizer(dataset="waymo")
    lab = add_frame_number(np.zeros((64, 256, 3), np.uint8), 3, pose_value=np.array([0.1234, -0.5678, 0.0]), font_scale=0.4, cond_num=2)
    assert lab.shape == (64, 256, 3) and (lab[..., 2] > 0).any() and not (lab[..., 1] > 0).any()          # past the conditioning frames: red text
    out = os.path.join(str(tmp_path), "single.mp4")
    write_video_single(d["image"][:, :, :64, :64], pose_values=d["pose"], save_path=out, cond_num=2, h=64, w=64)
    cap = cv2.VideoCapture(out)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == len(d["pose"]) and cap.get(cv2.CAP_PROP_FPS) == 10


@pytest.mark.skipif(not os.path.isfile("/root/reference/projects/tools/visulize.py"), reason="reference tree not mounted")
def test_live_against_the_reference_class(tmp_path):
    """Run in a subprocess: the reference's `projects` package and this repository's share a name."""
    import subprocess
    code = r'''
import os, sys, tempfile
import numpy as np
sys.path.insert(0, %r)
import tests._cases as C
from oracle import ref_import as R
R.load()
sys.path.insert(0, os.path.join(%r, "tests", "shims"))
with R.reference_cwd():
    import projects.tools.visulize as ref_vis
from umgen_b200 import visualize as V
import cv2
cv2.destroyAllWindows = lambda: None
os.chdir(%r)
bad = tot = 0
for seed in (301, 302, 303):
    C.VISUALIZE_CASES["live"] = (seed, 12, 512, 6, True, 64, (32, 64), 5)
    C.VISUALIZE_CASES["live_crowded"] = (seed, 12, 512, 6, True, 64, (32, 64), 5)
    for name in ("live", "live_crowded"):
        d = C.visualize_inputs(name)
        rv = ref_vis.Visulizer(video_save_path="rv/", video_pretext="U", width=512, height=512, project_name="P", spe_text="s", save_video=False,
                               addtion_ego=True, cond_frames=6, put_text=True)
        got = {}
        rv.generate_img_and_video = lambda imgs, *a, **k: got.update(frames=[np.array(f).copy() for f in imgs])
        rv.visulize(box=np.array([b.copy() for b in d["boxes"]], dtype=object), scene_name="s", pose=d["pose"].copy(), real_pose=d["real_pose"].copy(),
                    maps={"map": d["maps"].clone()}, decoded_image=d["image"].clone())
        fr = V.SceneVideo(width=512, height=512, project_name="P", spe_text="s", addtion_ego=True, cond_frames=6, put_text=True).compose(
            d["boxes"], d["pose"], d["real_pose"], d["maps"], d["image"], "s")
        assert len(fr) == len(got["frames"])
        for a, b in zip(fr, got["frames"]):
            tot += 1
            bad += int(a.shape != b.shape or (a != b).any())
print("LIVE", tot, bad)
''' % (ROOT, ROOT, str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    tot, bad = [int(x) for x in [l for l in r.stdout.splitlines() if l.startswith("LIVE")][-1].split()[1:]]
    assert tot == 72 and bad == 0, (tot, bad)
