"""BASELINE configs[0]: the reference's UNMODIFIED `projects/tools/evaluate.py` driving this repository's drop-in `projects` package
(`--infer_task video --set_num_new_frames 1`, one synthetic tokenised scene, greedy-irrelevant plumbing run).

The reference tree is needed for its driver files (evaluate.py, infer_fun.py, model_pl.py, configs/, plugin/): the test builds a
working directory of symlinks -- `projects/{__init__,registry,models,tokenizer/vq_model,tools/decode_map,tools/visulize,plugin/data/datasets}` from THIS
repository, everything else from the reference -- and runs evaluate.py there with stand-ins for the packages this image lacks (tests/shims).  Skipped where the reference is
not mounted (the GPU boxes).  Without CUDA the engine and pixel decoders are shape-correct fakes (tests/shims/cpu_stubs.py): the run then proves
the plumbing -- config -> dataset -> transforms -> registry -> UMGen(config) -> Lightning harness -> UMGen.inference signature -> token pickle ->
value decode -> decoder classes -> visualiser; with CUDA the same command runs the real engine and VQ decoders."""
import json
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("UMGEN_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "projects", "tools", "evaluate.py")), reason="reference tree not mounted")


def _link(src, dst):
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    os.symlink(src, dst)


def build_workdir(tmp):
    """cwd for evaluate.py: ours where this repository re-implements the reference, the reference's own files elsewhere."""
    ours, ref = os.path.join(ROOT, "projects"), os.path.join(REF, "projects")
    for rel in ("__init__.py", "registry.py", "models", "tokenizer/__init__.py", "tokenizer/vq_model.py", "tools/__init__.py", "tools/decode_map.py",
                "tools/visulize.py", "plugin/data/datasets"):
        _link(os.path.join(ours, rel), os.path.join(tmp, "projects", rel))
    for rel in ("configs", "plugin/misc", "plugin/data/transforms", "tokenizer/weights", "tools/evaluate.py", "tools/infer_fun.py", "tools/model_pl.py"):
        _link(os.path.join(ref, rel), os.path.join(tmp, "projects", rel))
    _link(os.path.join(ROOT, "umgen_b200"), os.path.join(tmp, "lib", "umgen_b200"))       # the engine package without the repo's own `projects/`
    _link(os.path.join(ROOT, "include"), os.path.join(tmp, "lib", "include"))
    return tmp


def make_raw_scene(path, n=110, seed=0):
    """A tokenised nuPlan scene pickle with the schema NuPlanTokenDataset reads (UMGen_nuplan_dataset.py:211-306)."""
    rs = np.random.RandomState(seed)
    cats = [c.strip() for c in open(os.path.join(REF, "projects", "configs", "category.txt")) if c.strip()]
    k = 12
    track_ids = np.arange(100, 100 + k)
    base = np.stack([rs.uniform(-40, 40, k), rs.uniform(-40, 40, k), np.zeros(k), rs.uniform(3, 6, k), rs.uniform(1.5, 2.5, k), rs.uniform(1.4, 2, k),
                     rs.uniform(-3, 3, k), rs.uniform(-3, 3, k), rs.uniform(-1, 1, k), np.zeros(k)], axis=1)
    names = [cats[i % len(cats)] for i in range(k)]
    meta, heading = [], np.cumsum(rs.uniform(-0.01, 0.01, n))
    for i in range(n):
        T = np.eye(4)
        T[:2, :2] = [[np.cos(heading[i]), -np.sin(heading[i])], [np.sin(heading[i]), np.cos(heading[i])]]
        T[0, 3], T[1, 3] = 0.6 * i, 0.02 * i
        boxes = base.copy()
        boxes[:, 0] += 0.05 * i * base[:, 7]
        boxes[:, 1] += 0.05 * i * base[:, 8]
        meta.append({"T_lidar2global": T, "bboxes_3d": boxes.astype(np.float32), "track_ids": track_ids.copy(), "categories": list(names)})
    ego = np.zeros((n, 16))
    ego[:, 6] = heading
    scene = {
        "tokens": {"CAM_F0": {"tokens": [rs.randint(0, 8192, (16, 32)) for _ in range(n)], "file_list": [f"{i:06d}.jpg" for i in range(n)]}},
        "raster_tokens": rs.randint(0, 8192, (n, 32, 32)),
        "ego_pose_all": ego,
        "meta_info": meta,
        "lidar_bboxes": {"CAM_F0": {"bboxes_3d": [m["bboxes_3d"] for m in meta], "categories": [m["categories"] for m in meta],
                                    "track_ids": [m["track_ids"] for m in meta]}},
    }
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump(scene, f)


def test_evaluate_py_drives_the_dropin_package_unchanged(tmp_path):
    wd = build_workdir(str(tmp_path))
    make_raw_scene(os.path.join(wd, "data", "tokenized_origin_scenes", "synthetic_scene_0001_clip_a.pkl"))
    real = torch.cuda.is_available()
    from umgen_b200 import synth
    for kind, name in (("map", "map_vae.ckpt"), ("image", "image_vae.tar")):
        os.makedirs(os.path.join(wd, "data", "weights"), exist_ok=True)
        torch.save({"state_dict": synth.make_vq_state_dict(kind, seed=1) if real else {}}, os.path.join(wd, "data", "weights", name))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests", "shims"), os.path.join(wd, "lib")]), UMGEN_SKIP_INIT="1")
    cmd = [sys.executable, os.path.join(ROOT, "tests", "shims", "run_evaluate.py"), os.path.join(wd, "projects", "tools", "evaluate.py"),
           "--infer_task", "video", "--set_num_new_frames", "1", "--debug", "1", "--model_scale", "debug", "--output_path", "output/UMGen/",
           "--map_decoder_weights_path", "data/weights/map_vae.ckpt", "--image_decoder_weights_path", "data/weights/image_vae.tar"]
    r = subprocess.run(cmd, cwd=wd, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    assert "Sucess" in r.stdout                                         # evaluate.py's own last line
    pk = [f for f in os.listdir(os.path.join(wd, "output", "UMGen", "saved_token")) if f.endswith("_tokens.pkl")]
    assert len(pk) == 1, pk
    out = pickle.load(open(os.path.join(wd, "output", "UMGen", "saved_token", pk[0]), "rb"))
    for m, width in (("pose", 3), ("map", 1024), ("bbox3d", 660), ("image", 512)):
        assert out[m].dtype == np.int64 and out[m].shape == (1, 21, width), (m, out[m].shape)
    assert out["map"].max() < 8192 and out["bbox3d"].max() <= 1027 and out["pose"].max() < 1024
    # the scene video, composed by this repository's visualiser (projects/tools/visulize.py -> umgen_b200/visualize.py): 21 frames of BEV canvas + camera image
    import cv2
    vids = [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(wd, "output")) for f in fs if f.endswith(".mp4")]
    assert len(vids) == 1, vids
    cap = cv2.VideoCapture(vids[0])
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 21 and int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)) == 512
    if not real:
        calls = json.loads([l for l in r.stdout.splitlines() if l.startswith("STUB_CALLS ")][-1][len("STUB_CALLS "):])
        inf = [c for c in calls if c[0] == "inference"]
        assert inf == [["inference", 1, 20, 20, "pose_map_bbox3d_image", ["cond_on_par", "infer_from_gt"]]], inf     # model_pl.py:30-36,237-239
        # projects/tools/decode_map.py's Mapdecoder / Imagedecoder (subclasses of the decoders the stubs replaced) were built from the ckpt paths
        assert {c[1] for c in calls if c[0] == "decoder"} == {"Mapdecoder", "Imagedecoder"}


def test_evaluate_py_control_task_drives_the_dropin_package(tmp_path):
    """BASELINE configs[3] through the unmodified evaluate.py: `--infer_task control` reads data/controlled_scenes/*.pkl (the dataset hands the
    pickle over as it is, UMGen_nuplan_dataset.py:202-206), conditions on 13 frames, forces 30 frames of ego poses and one agent slot
    (infer_fun.py:64-67; model_pl.py:135-198) and always writes the scene video."""
    if torch.cuda.is_available():
        pytest.skip("plumbing run of the control task: the real engine's control path is covered by tests/test_engine_gpu.py")
    from umgen_b200 import synth
    wd = build_workdir(str(tmp_path))
    scene = synth.make_scene(seed=4, n_frames=13)
    ctrl = synth.make_control(seed=4, n_frames=30, slot=2)
    item = {"dataset_token": {m: scene[m][0, :13] for m in ("pose", "map", "bbox3d", "image")}, "control_dict": {m: v[0] for m, v in ctrl.items()},
            "scene_name": "synthetic_scene_0004_shift_left", "control_object": 2, "input_cond_frame": 13}
    os.makedirs(os.path.join(wd, "data", "controlled_scenes"))
    with open(os.path.join(wd, "data", "controlled_scenes", "synthetic_scene_0004_shift_left.pkl"), "wb") as f:
        pickle.dump(item, f)
    os.makedirs(os.path.join(wd, "data", "weights"), exist_ok=True)
    for name in ("map_vae.ckpt", "image_vae.tar"):
        torch.save({"state_dict": {}}, os.path.join(wd, "data", "weights", name))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests", "shims"), os.path.join(wd, "lib")]), UMGEN_SKIP_INIT="1")
    cmd = [sys.executable, os.path.join(ROOT, "tests", "shims", "run_evaluate.py"), os.path.join(wd, "projects", "tools", "evaluate.py"),
           "--infer_task", "control", "--debug", "1", "--model_scale", "debug", "--output_path", "output/UMGen/",
           "--map_decoder_weights_path", "data/weights/map_vae.ckpt", "--image_decoder_weights_path", "data/weights/image_vae.tar"]
    r = subprocess.run(cmd, cwd=wd, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    assert "Sucess" in r.stdout
    calls = json.loads([l for l in r.stdout.splitlines() if l.startswith("STUB_CALLS ")][-1][len("STUB_CALLS "):])
    assert [c for c in calls if c[0] == "inference"] == [["inference", 30, 20, 13, "pose_map_bbox3d_image", ["cond_on_par", "infer_from_gt"]]]
    assert [c for c in calls if c[0] == "inference_control"] == [["inference_control", True, {"bbox3d": [1, 30, 660], "pose": [1, 30, 3]}]]
    out = pickle.load(open(os.path.join(wd, "output", "UMGen", "saved_token", "synthetic_scene_0004_shift_left_tokens.pkl"), "rb"))
    assert out["bbox3d"].shape == (1, 43, 660) and out["pose"].dtype == np.int64
    import cv2
    vids = [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(wd, "output")) for f in fs if f.endswith(".mp4")]
    assert len(vids) == 1 and os.path.basename(vids[0]) == "UMGen_synthetic_scene_0004_shift_left.mp4", vids
    assert int(cv2.VideoCapture(vids[0]).get(cv2.CAP_PROP_FRAME_COUNT)) == 43
