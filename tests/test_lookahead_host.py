"""Host logic of the look-ahead schedule (umgen_b200/engine.py): which frames of the next window are known in advance, and the check that an
incoming window really continues the previous frame (no GPU needed)."""
import torch

from umgen_b200.config import CONTENT_LEN, MODS
from umgen_b200.engine import next_window_start, window_continues


def _window(T, seed):
    g = torch.Generator().manual_seed(seed)
    return {m: torch.randint(0, 1000, (T, CONTENT_LEN[m]), generator=g) for m in MODS}


def test_next_window_start_slides_once_the_window_is_full():
    assert next_window_start(13, 20) == 0          # control mode: 13 conditioning frames grow to 20
    assert next_window_start(19, 20) == 0
    assert next_window_start(20, 20) == 1
    assert next_window_start(2, 2) == 1


def _assumed(cur, pose_new, window):
    s = next_window_start(cur["pose"].shape[0], window)
    return {"T": cur["pose"].shape[0] - s + 1, "host": {m: cur[m][s:].clone().long() for m in MODS}, "pose_new": pose_new}


def test_window_continues_accepts_only_the_expected_window():
    cur = _window(20, 1)
    pose_new = torch.tensor([5, 6, 7], dtype=torch.int32)
    new = {m: torch.randint(0, 1000, (CONTENT_LEN[m],)) for m in MODS}
    new["pose"] = pose_new.long()
    la = _assumed(cur, pose_new, 20)
    nxt = {m: torch.cat([cur[m][1:], new[m][None]]) for m in MODS}
    assert window_continues(la, nxt)
    assert window_continues(la, {m: v.to(torch.int32) for m, v in nxt.items()})        # dtype of the caller's tokens does not matter
    assert not window_continues(None, nxt)
    bad = {m: v.clone() for m, v in nxt.items()}
    bad["bbox3d"][3, 17] += 1                                                           # an earlier frame was edited
    assert not window_continues(la, bad)
    bad = {m: v.clone() for m, v in nxt.items()}
    bad["pose"][-1, 0] += 1                                                             # the new frame carries another ego action
    assert not window_continues(la, bad)
    ok = {m: v.clone() for m, v in nxt.items()}
    ok["image"][-1, 0] += 1                                                             # the new frame's own content is free
    assert window_continues(la, ok)
    assert not window_continues(la, {m: v[1:] for m, v in nxt.items()})                 # another length


def test_growing_window():
    cur = _window(13, 2)
    pose_new = torch.tensor([1, 2, 3])
    la = _assumed(cur, pose_new, 20)
    assert la["T"] == 14
    new = {m: torch.zeros(CONTENT_LEN[m], dtype=torch.long) for m in MODS}
    new["pose"] = pose_new
    assert window_continues(la, {m: torch.cat([cur[m], new[m][None]]) for m in MODS})
