"""The whole of SURVEY.md section 8 in one run on the GPU: raw tokenised scene pickle -> dataset front-end (8f rank 3) -> scene runner ->
`UMGen.inference` of the drop-in module on the engine (8a) -> token pickle and value decode (8f rank 2) -> VQ pixel decoders (a11) -> scene
video (8f rank 4).  Depth-1 model, two new frames.  (Named to sort last: a failure here hides no other GPU test under `-x`.)"""
import os
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_raw_scene_to_tokens_to_pixels_to_video(tmp_path):
    import cv2
    from projects.models.UMGen import UMGen
    from tests._cases import DATASET_CASES, raw_scene
    from tests.test_dropin_surface import eval_namespace
    from umgen_b200 import runner as R, synth
    from umgen_b200.config import ModelConfig
    from umgen_b200.dataset import NuPlanTokenScenes
    from umgen_b200.visualize import SceneVideo
    from umgen_b200.vq import Imagedecoder, Mapdecoder

    seed, n, block, gap, n_tracks = DATASET_CASES["short_clip"]
    root = tmp_path / "scenes"
    root.mkdir()
    with open(root / "synthetic_scene_0000_clip_a.pkl", "wb") as f:
        pickle.dump(raw_scene(seed, n, n_tracks), f)
    scenes = NuPlanTokenScenes([str(root)], block_size=block, sampling_gap=gap)
    T = int(scenes.batch(0)["pose"].shape[1])
    cond = min(T, 13)

    model = UMGen(eval_namespace(layers=1, cond_frame=20, top_k=1, top_k_map=1)).eval()
    model.load_state_dict(synth.make_state_dict(ModelConfig.tiny(1), seed=3), strict=False)
    model.sample_param_map = model.topk_image = 1
    model.cuda()
    md, idec = Mapdecoder(synth.make_vq_state_dict("map", seed=1)), Imagedecoder(synth.make_vq_state_dict("image", seed=1))
    video = SceneVideo(video_save_path=str(tmp_path / "clips") + "/", video_pretext="UMGen", width=512, height=512, project_name="UMGen_infer",
                       spe_text="synthetic_video", addtion_ego=True, cond_frames=cond, put_text=True)
    s = R.RunSettings(new_frames=2, cond_frames=cond, input_cond_frames=cond, token_save_path=str(tmp_path / "tokens"), generate_video=True)
    (res,) = R.run_dataset(model, scenes, s, md, idec, video=video)

    tok = res["tokens"]
    batch = scenes.batch(0)
    for m, w in (("pose", 3), ("map", 1024), ("bbox3d", 660), ("image", 512)):
        assert tok[m].shape == (1, cond + 2, w) and tok[m].dtype == np.int64
        assert np.array_equal(tok[m][0, :cond], batch[m][0, :cond].numpy())            # the conditioning frames come back unchanged
    assert tok["map"].max() < 8192 and tok["image"].max() < 8192 and tok["bbox3d"].max() <= 1027 and tok["pose"].max() < 1024
    assert pickle.load(open(res["token_path"], "rb"))["map"].shape == (1, cond + 2, 1024)
    bboxes, anno, pose, real_pose, maps, image, _ = res["decoded"]
    assert len(bboxes) == cond + 2 and pose.shape == (cond + 2, 3) and real_pose.shape == (T, 3) and len(anno) == T
    assert maps.shape == (cond + 2, 3, 256, 256) and image.shape == (cond + 2, 3, 256, 512)
    assert torch.isfinite(maps).all() and torch.isfinite(image).all() and float(image.abs().max()) > 0.05
    cap = cv2.VideoCapture(res["video_path"])
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == cond + 2
    assert (int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))) == (512, 512 + 256)
    ok, frame = cap.read()
    assert ok and frame[:512].std() > 5 and frame[512:].std() > 5                      # the BEV canvas shows the map, the lower part the camera image
    # the same scene again: its token pickle exists, so it is skipped like in the reference (model_pl.py:214-215)
    assert R.run_dataset(model, scenes, s, md, idec, video=video) == []
