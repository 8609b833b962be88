"""Host-side post-processing (umgen_b200/postprocess.py) against the reference's own tokenizers / normalisers
(tests/golden/postprocess.npz from oracle/make_golden.py: tools/model_pl.py:262-291 run on a seeded scene)."""
import os
import time

import numpy as np

from umgen_b200 import postprocess as P


def test_bbox3d_and_pose_values_match_the_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "postprocess.npz"))
    boxes, classes = P.decode_bbox3d(g["bbox_tokens"])
    assert len(boxes) == g["bboxes"].shape[0]
    np.testing.assert_array_equal(np.stack(boxes), g["bboxes"])            # float64, bit-exact
    names = ["none", "vehicle", "bicycle", "pedestrian"]
    got = np.array([[names.index(c) for c in row] for row in classes], dtype=np.int8)
    np.testing.assert_array_equal(got, g["classes"])
    np.testing.assert_array_equal(P.decode_pose(g["pose_tokens"]), g["pose_values"])
    # the batched [1, T, 660] form decode_tokens is called with
    b2, _ = P.decode_bbox3d(g["bbox_tokens"][None])
    np.testing.assert_array_equal(np.stack(b2), g["bboxes"])


def test_pad_slots_and_edges():
    tok = np.full((1, 660), P.PAD_TOKEN, dtype=np.int64)
    tok[0, :11] = [0, 1023, 512, 1024, 1026, 5, 7, 9, 11, 13, 1025]
    boxes, classes = P.decode_bbox3d(tok)
    assert boxes[0].shape == (60, 10) and classes[0][0] == "bicycle" and set(classes[0][1:]) == {"none"}
    assert boxes[0][1, 0] == 64.0 and boxes[0][1, 6] == 3.14            # <pad> decodes to the upper end of every range
    assert boxes[0][0, 0] == -64.0                                       # token 0 -> bins[0]
    assert boxes[0][0, 3] == boxes[0][0, 4] / 4 * 15                     # category ids in attribute positions clip to the last bin
    p = P.decode_pose(np.array([[0, 1023, 2000]]))[0]
    assert abs(p[0] + 10.0) < 1e-6                                       # token 0 -> bins[0] = -1, / float32(0.1)
    assert 3.99 < p[1] < 4.0 and p[2] == P.decode_pose(np.array([[0, 0, 1023]]))[0, 2]      # last bin's midpoint; ids beyond 1023 clip


def test_token_pickle_round_trip(tmp_path):
    out = {m: np.arange(2 * n, dtype=np.int64).reshape(1, 2, n) for m, n in (("pose", 3), ("map", 1024), ("bbox3d", 660), ("image", 512))}
    path = P.save_tokens(out, str(tmp_path / "tokens"), "scene_0001")
    assert path.endswith("scene_0001_tokens.pkl")
    back = P.load_tokens(path)
    assert list(back) == ["pose", "map", "bbox3d", "image"] and all(np.array_equal(back[m], out[m]) for m in out)
    vals = P.decode_scene(out)
    assert len(vals["bboxes"]) == 2 and vals["pose_values"].shape == (2, 3)


def test_vectorised_decode_is_fast():
    rs = np.random.RandomState(0)
    tok = rs.randint(0, 1028, size=(50, 660))
    t0 = time.time()
    for _ in range(20):
        P.decode_bbox3d(tok)
    assert (time.time() - t0) / 20 < 0.05        # a 50-frame scene in well under 50 ms


def test_annotation_side_matches_the_reference(golden_dir):
    """model_pl.py:278-286: the ground-truth boxes go through decode() WITHOUT keep_order -- slots holding any <pad> and slots whose category id is
    out of range are dropped, frame by frame."""
    g = np.load(os.path.join(golden_dir, "postprocess.npz"))
    boxes, classes = P.decode_annotation_bbox3d(g["gt_bbox_tokens"])
    assert [len(b) for b in boxes] == g["anno_n"].tolist()
    assert 0 < g["anno_n"].sum() < 60 * len(boxes)                       # the filter removed some slots and kept some
    np.testing.assert_array_equal(np.concatenate(boxes, axis=0), g["anno_boxes"])
    names = ["none", "vehicle", "bicycle", "pedestrian"]
    np.testing.assert_array_equal(np.array([names.index(c) for row in classes for c in row], dtype=np.int8), g["anno_classes"])
    # a frame of nothing but <pad> yields an empty array, not an error
    b, c = P.decode_annotation_bbox3d(np.full((2, 660), P.PAD_TOKEN))
    assert [x.shape for x in b] == [(0, 10), (0, 10)] and c == [[], []]


def test_decode_tokens_value_part_without_a_gpu(golden_dir):
    g = np.load(os.path.join(golden_dir, "postprocess.npz"))
    T = g["bbox_tokens"].shape[0]
    pred = {"pose": g["pose_tokens"][None], "bbox3d": g["bbox_tokens"][None], "map": np.zeros((1, T, 1024), np.int64), "image": np.zeros((1, T, 512), np.int64)}
    gt = {"pose": g["pose_tokens"][None], "bbox3d": g["gt_bbox_tokens"][None]}
    bboxes, anno, pose, real_pose, maps, image, map_tr = P.decode_tokens(pred, gt)
    np.testing.assert_array_equal(np.stack(bboxes), g["bboxes"])
    np.testing.assert_array_equal(pose, g["pose_values"])
    np.testing.assert_array_equal(real_pose, g["pose_values"])
    assert [len(a) for a in anno] == g["anno_n"].tolist()
    assert maps is None and image is None and map_tr is None             # no decoder given: the pixel part is the GPU's
