"""The oracle restatement replays the rollouts the real reference produced (tests/golden/rollout_*.npz,
made by oracle/make_golden.py): identical greedy token ids, matching conditioning features and logits."""
import os

import numpy as np
import pytest
import torch

from oracle import umgen_oracle as O
from tests._cases import ROLLOUT_CASES
from umgen_b200 import synth
from umgen_b200.config import ModelConfig


def run_oracle_case(spec):
    cfg = ModelConfig.tiny(spec["layers"])
    P = synth.make_state_dict(cfg, seed=spec["weight_seed"])
    ocfg = O.ModelCfg.tiny(spec["layers"])
    ocfg.cond_frame = spec["cond_frames"]
    orc = O.UMGenOracle(P, ocfg, O.SampleCfg.greedy())
    orc.keep_trace = True
    scene = synth.make_scene(seed=spec["scene_seed"], n_frames=spec["input_frames"])
    init = synth.make_control(seed=spec["scene_seed"], n_frames=spec["new_frames"]) if spec.get("control") else None
    with torch.no_grad():
        out = orc.inference(spec["new_frames"], spec["cond_frames"], spec["input_cond_frames"], scene, init,
                            control_test=bool(spec.get("control")))
    return orc, out


@pytest.mark.parametrize("name", ["video_L1", "control_L1"])
def test_oracle_replays_reference_rollout(name, golden_dir):
    path = os.path.join(golden_dir, f"rollout_{name}.npz")
    g = np.load(path)
    orc, out = run_oracle_case(ROLLOUT_CASES[name])
    for f, tr in enumerate(orc.trace):
        got = tr.tar_feat[::13].numpy()
        np.testing.assert_allclose(got, g["tar_feat"][f], rtol=0, atol=2e-4)
        if tr.ego_logits is not None:
            np.testing.assert_allclose(tr.ego_logits.numpy(), g["ego_logits"][f], rtol=0, atol=2e-4)
        pos = sorted(p for p in tr.logits if p > 0)
        assert len(pos) == 2196
        top = torch.stack([torch.topk(tr.logits[p], 8).values for p in pos]).numpy()
        np.testing.assert_allclose(top, g["ar_top_vals"][f], rtol=0, atol=3e-4)
    for m in O.MODS:
        assert out[m].shape == g[f"out_{m}"].shape
        assert np.array_equal(out[m], g[f"out_{m}"]), m
