"""The dataset front-end (umgen_b200/dataset.py) against the reference's own NuPlanTokenDataset + evaluation transforms
(tests/golden/dataset.npz from oracle/make_golden.py: the reference class run on the synthetic raw scenes of tests/_cases.py).  Token ids are
compared for equality, the raw ego deltas bit for bit."""
import os
import pickle

import numpy as np
import pytest
import torch

from tests._cases import DATASET_CASES, raw_scene
from umgen_b200 import dataset as D


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "dataset.npz"))


@pytest.mark.parametrize("name", list(DATASET_CASES))
def test_scene_tokens_match_the_reference_dataset(name, golden, tmp_path):
    seed, n, block, gap, n_tracks = DATASET_CASES[name]
    path = tmp_path / f"synthetic_scene_{seed:04d}_clip_a.pkl"
    with open(path, "wb") as f:
        pickle.dump(raw_scene(seed, n, n_tracks), f)
    ds = D.NuPlanTokenScenes([str(tmp_path)], block_size=block, sampling_gap=gap)
    assert len(ds) == 1
    d = ds[0]
    assert list(d) == ["pose", "map", "pose_diff", "bbox3d", "image", "file_name"]
    for k in ("pose", "map", "bbox3d", "image"):
        g = golden[f"{name}.{k}"]
        assert d[k].dtype == torch.int64 and tuple(d[k].shape) == g.shape, (k, d[k].shape, g.shape)
        np.testing.assert_array_equal(d[k].numpy(), g, err_msg=k)
    assert d["pose_diff"].dtype == torch.float32
    np.testing.assert_array_equal(d["pose_diff"].numpy(), golden[f"{name}.pose_diff"])          # float64 arithmetic in the reference's order, narrowed by ToTensor
    assert d["file_name"] == "0_" + str(path)
    b = ds.batch(0)
    assert b["bbox3d"].shape == (1,) + tuple(d["bbox3d"].shape) and b["file_name"] == [d["file_name"]]


def test_the_cases_reach_the_corner_cases(golden):
    """The synthetic scenes are only worth something if they exercise the branches: overflowing slot table, empty frames, out-of-vocabulary and
    out-of-range objects, heading wrap, shortened clip."""
    seed, n, block, gap, n_tracks = DATASET_CASES["dense_gap1"]
    sc = raw_scene(seed, n, n_tracks)
    frames = D.frame_indices(n, block, gap)
    valid = set()
    dropped_cat = dropped_range = 0
    for f in frames:
        m = sc["meta_info"][f]
        for b, c, t in zip(m["bboxes_3d"], m["categories"], m["track_ids"]):
            if c not in D.CATEGORIES:
                dropped_cat += 1
            elif abs(b[0]) > 64 or abs(b[1]) > 64:
                dropped_range += 1
            else:
                valid.add(int(t))
    assert len(valid) > 60 and dropped_cat > 0 and dropped_range > 0
    bb = golden["dense_gap1.bbox3d"].reshape(len(frames), 60, 11)
    assert (bb[:, :, 0] != D.PAD_TOKEN).any(axis=0).all()                  # every slot is used at some point: the table is full
    assert ((bb != D.PAD_TOKEN).sum(axis=(1, 2)) == 0).any()               # and some frame is empty
    assert len(golden["short_clip.pose"]) == (70 - 4 - 1) // 4 < 50        # the clip was shortened (UMGen_nuplan_dataset.py:152-161)
    sc = raw_scene(DATASET_CASES["video_50"][0], DATASET_CASES["video_50"][1], DATASET_CASES["video_50"][4])
    h = sc["ego_pose_all"][:, 6]
    assert (np.abs(np.diff(h)) > 6).any()                                  # the heading wraps through +-pi inside the scene
    assert np.abs(golden["video_50.pose_diff"][:, 2]).max() < 1.0          # ... and the deltas do not


def test_building_blocks():
    assert D.frame_indices(240, 50, 4) == [10 + 4 * i for i in range(50)]
    assert D.frame_indices(240, 50, 4, inference=False)[0] == 4
    assert D.frame_indices(208, 50, 4)[0] == 4 and len(D.frame_indices(208, 50, 4)) == 50        # start pulled forward
    # values outside the normalisation range land in the first / last bin; the category column is 1024 + index
    box = np.array([[-100, 100, 0, 5, 2, 1.5, 0, 0, 0, 0]], dtype=np.float32)
    t = D.box_attribute_tokens(box)[0]
    assert t[0] == 0 and t[1] == 1023 and 0 < t[2] < 1023
    tok = D.bbox3d_tokens([box * 0 + np.array([1, 1, 0, 5, 2, 1.5, 0, 0, 0, 0], np.float32)], [["pedestrian"]], [np.array([42])])
    assert tok.shape == (1, 660) and tok[0, 10] == 1026 and (tok[0, 11:] == D.PAD_TOKEN).all()
    # a frame whose only track id is 0 is treated as empty by the reference (np.any) -- kept
    tok0 = D.bbox3d_tokens([box * 0 + 1], [["vehicle"]], [np.array([0])])
    assert (tok0 == D.PAD_TOKEN).all()
    # a track that shows up after 60 others never gets a slot
    boxes = [np.tile(np.array([[1, 1, 0, 5, 2, 1.5, 0, 0, 0, 0]], np.float32), (61, 1))]
    tokn = D.bbox3d_tokens(boxes, [["vehicle"] * 61], [np.arange(1, 62)])
    assert (tokn.reshape(60, 11)[:, 10] == 1024).all()


def test_control_scene_is_passed_through(tmp_path):
    payload = {"pose": np.zeros((1, 43, 3), np.int64), "control_dict": {"pose": np.ones((1, 30, 3), np.int64)}}
    with open(tmp_path / "ctrl_0001.pkl", "wb") as f:
        pickle.dump(payload, f)
    ds = D.NuPlanTokenScenes([str(tmp_path)], block_size=43, control_test=True)
    got = ds[0]
    assert set(got) == {"pose", "control_dict"} and np.array_equal(got["control_dict"]["pose"], payload["control_dict"]["pose"])
