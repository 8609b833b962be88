"""The scene runner (umgen_b200/runner.py) against the reference's own UMGen_PL.world_model_evaluate / generate_init_tokens
(tools/model_pl.py:95-262): the keyword arguments handed to `model.inference` and the token pickle written, for the free-rollout and the
controlled branch (tests/golden/runner.json: the reference's method bodies run on a recording model, oracle/make_golden.py runner)."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from tests._cases import RUNNER_CASES, runner_batch, summarise_inference_kwargs
from umgen_b200 import runner as R

WIDTH = {"pose": 3, "map": 1024, "bbox3d": 660, "image": 512}


class Recorder:
    """Stands in for projects.models.UMGen.UMGen: records the call, returns tokens of the right shape."""

    def __init__(self):
        self.calls = []

    def inference(self, **kw):
        self.calls.append(kw)
        n = int(kw["input_cond_frames"]) + int(kw["new_frames"])
        rs = np.random.RandomState(n)
        return {m: rs.randint(0, 1024, size=(1, n, w)).astype(np.int64) for m, w in WIDTH.items()}


@pytest.mark.parametrize("name", list(RUNNER_CASES))
def test_inference_kwargs_match_the_reference_harness(name, golden_dir, tmp_path):
    golden = json.load(open(os.path.join(golden_dir, "runner.json")))[name]
    task, new_frames, init_mod, _, _ = RUNNER_CASES[name]
    s = R.RunSettings(new_frames=new_frames, infer_task=task, init_token_mod=init_mod, token_save_path=str(tmp_path))
    model = Recorder()
    res = R.run_scene(model, runner_batch(name), s)
    assert len(model.calls) == 1
    assert json.loads(json.dumps(summarise_inference_kwargs(model.calls[0]))) == golden["kwargs"]
    assert sorted(os.listdir(tmp_path)) == golden["saved"]
    back = pickle.load(open(res["token_path"], "rb"))
    assert list(back) == ["pose", "map", "bbox3d", "image"] and all(np.array_equal(back[m], res["tokens"][m]) for m in back)
    bboxes, anno, pose, real_pose, maps, image, map_tr = res["decoded"]
    n = res["tokens"]["pose"].shape[1]
    assert len(bboxes) == n and pose.shape == (n, 3) and maps is None and image is None           # no decoders given: values only
    assert anno is not None and real_pose.shape[1] == 3                                          # the annotation side comes from the batch


def test_a_processed_scene_is_skipped_and_a_dataset_is_walked(tmp_path):
    """model_pl.py:214-215: a dataset scene whose <name>_tokens.pkl exists is not generated again; run_dataset walks a NuPlanTokenScenes
    (or one rank's share of it) with batch 1."""
    from tests._cases import DATASET_CASES, raw_scene
    from umgen_b200.dataset import NuPlanTokenScenes
    seed, n, block, gap, n_tracks = DATASET_CASES["short_clip"]
    root = tmp_path / "scenes"
    root.mkdir()
    for k in range(3):
        with open(root / f"synthetic_scene_{k:04d}_clip_a.pkl", "wb") as f:
            pickle.dump(raw_scene(seed + k, n, n_tracks), f)
    scenes = NuPlanTokenScenes([str(root)], block_size=block, sampling_gap=gap)
    s = R.RunSettings(new_frames=2, cond_frames=13, input_cond_frames=13, token_save_path=str(tmp_path / "out"))
    model = Recorder()
    first = R.run_dataset(model, scenes, s, indices=[0, 2])
    assert [r["name"] for r in first] == ["synthetic_scene_0000_clip_a", "synthetic_scene_0002_clip_a"] and len(model.calls) == 2
    kw = model.calls[0]
    assert kw["input_cond_tokens"]["bbox3d"].shape == (1, 16, 660) and kw["input_cond_tokens"]["bbox3d"].dtype == torch.int64
    again = R.run_dataset(model, scenes, s)
    assert [r["name"] for r in again] == ["synthetic_scene_0001_clip_a"] and len(model.calls) == 3      # 0 and 2 were skipped
    assert first[0]["tokens"]["map"].shape == (1, 15, 1024)


def test_generate_init_tokens_branches():
    gt = {"pose": torch.arange(2 * 5 * 3).view(2, 5, 3), "map": torch.zeros(2, 5, 4, dtype=torch.int64)}
    a = R.generate_init_tokens(gt, 3, ["pose"])
    assert list(a) == ["pose"] and torch.equal(a["pose"], gt["pose"][:, 3:]) and a["pose"].data_ptr() != gt["pose"].data_ptr()
    c = R.generate_init_tokens(gt, 3, None, {"pose": torch.ones(30, 3), "bbox3d": torch.ones(1, 30, 660)})
    assert c["pose"].shape == (1, 30, 3) and c["bbox3d"].shape == (1, 30, 660)
    assert R.generate_init_tokens(gt, 3) is None


class FakePixels:
    """Shape-correct stand-in for the GPU pixel decoders (umgen_b200/vq.py needs a CUDA device): tokens [1, t, S] -> floats [t, 3, h, w]."""

    def __init__(self, h, w):
        self.h, self.w = h, w

    def decode_maps(self, tok):
        t = np.asarray(tok)
        return torch.from_numpy((t[0, :, :3].astype(np.float32) / 512 - 1)[:, :, None, None].repeat(self.h, 2).repeat(self.w, 3))

    decode_images = decode_maps


def test_scene_videos_are_written_like_generate_videos_does(tmp_path):
    """model_pl.py:188-198, 260-274: a controlled scene always gets its video, a dataset scene when generate_video_flag is set or batch_idx % 100 == 0;
    the caption of a controlled scene names the controlled object (:140-147)."""
    import cv2
    from umgen_b200.visualize import SceneVideo
    video = SceneVideo(video_save_path=str(tmp_path / "clips") + "/", video_pretext="UMGen", width=512, height=512, project_name="UMGen_infer",
                       spe_text="x_video", addtion_ego=True, cond_frames=20, put_text=True)
    name = next(k for k, v in RUNNER_CASES.items() if "control" not in v[0])
    s = R.RunSettings(new_frames=RUNNER_CASES[name][1], infer_task=RUNNER_CASES[name][0], init_token_mod=RUNNER_CASES[name][2])
    kw = dict(mapdecoder=FakePixels(64, 64), imagedecoder=FakePixels(32, 64), video=video)
    assert R.run_scene(Recorder(), runner_batch(name), s, batch_idx=7, **kw)["video_path"] is None
    r = R.run_scene(Recorder(), runner_batch(name), s, batch_idx=100, **kw)
    n = r["tokens"]["pose"].shape[1]
    cap = cv2.VideoCapture(r["video_path"])
    assert os.path.basename(r["video_path"]) == f"UMGen_{r['name']}.mp4" and int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == n
    assert (int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))) == (512, 512 + 32)
    s2 = R.RunSettings(new_frames=RUNNER_CASES[name][1], infer_task=RUNNER_CASES[name][0], init_token_mod=RUNNER_CASES[name][2], generate_video=True)
    assert R.run_scene(Recorder(), runner_batch(name), s2, batch_idx=7, **kw)["video_path"] is not None
    cname = next(k for k, v in RUNNER_CASES.items() if "control" in v[0])
    sc = R.RunSettings(new_frames=RUNNER_CASES[cname][1], infer_task=RUNNER_CASES[cname][0], init_token_mod=RUNNER_CASES[cname][2])
    rc = R.run_scene(Recorder(), runner_batch(cname), sc, batch_idx=7, **kw)
    assert rc["video_path"] is not None and video.spe_text.startswith("x_video") and video.spe_text != "x_video"
