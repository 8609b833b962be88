"""The oracle restatement replays the rollouts the real reference produced (tests/golden/rollout_*.npz,
made by oracle/make_golden.py): identical greedy token ids, matching conditioning features and logits."""
import os

import numpy as np
import pytest
import torch

from oracle import umgen_oracle as O
from tests._cases import OAR_CASES, ROLLOUT_CASES, apply_tweak, oar_inputs, rollout_init
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
    init = rollout_init(spec, scene)
    with torch.no_grad():
        out = orc.inference(spec["new_frames"], spec["cond_frames"], spec["input_cond_frames"], scene, init,
                            control_test=bool(spec.get("control")))
    return orc, out


SLOW = pytest.mark.skipif(os.environ.get("UMGEN_SLOW_TESTS") != "1", reason="minutes of CPU time; set UMGEN_SLOW_TESTS=1 (run once per oracle change, result in DESIGN.md)")


@pytest.mark.parametrize("name", ["video_L1", "control_L1", "initmap_L1", pytest.param("control_T13_L2", marks=SLOW), pytest.param("video_T20_L4", marks=SLOW)])
def test_oracle_replays_reference_rollout(name, golden_dir):
    path = os.path.join(golden_dir, f"rollout_{name}.npz")
    g = np.load(path)
    orc, out = run_oracle_case(ROLLOUT_CASES[name])
    n_sampled = g["input_stream"].shape[1]
    for f, tr in enumerate(orc.trace):
        got = tr.tar_feat[::13].numpy()
        np.testing.assert_allclose(got, g["tar_feat"][f], rtol=0, atol=2e-4)
        if tr.ego_logits is not None:
            np.testing.assert_allclose(tr.ego_logits.numpy(), g["ego_logits"][f], rtol=0, atol=2e-4)
        pos = sorted(p for p in tr.logits if p > 0)
        assert len(pos) == n_sampled                      # 2196, or 1172 when the map block is given as init tokens
        top = torch.stack([torch.topk(tr.logits[p], 8).values for p in pos]).numpy()
        np.testing.assert_allclose(top, g["ar_top_vals"][f], rtol=0, atol=3e-4)
    for m in O.MODS:
        assert out[m].shape == g[f"out_{m}"].shape
        assert np.array_equal(out[m], g[f"out_{m}"]), m


@pytest.mark.parametrize("name", ["oar_L1_collide", "oar_L1_padheavy"])
def test_oracle_replays_reference_decode_through_the_bbox_block(name, golden_dir):
    """infer_oar_net goldens whose bbox3d block takes the rare branches: collision-decided wipes (UMGen.py:1336-1377) and the TAR-head resample
    (UMGen.py:1092-1104).  The oracle decodes positions 1..1693 (pose, map, bbox3d) and must reproduce ids, the decode stream and the wipe count."""
    import dataclasses
    spec = OAR_CASES[name]
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    cfg = dataclasses.replace(ModelConfig.tiny(1), n_oar_layer=spec["oar_layers"])
    sd = apply_tweak(synth.make_state_dict(cfg, seed=spec["weight_seed"]), spec.get("tweak"))
    orc = O.UMGenOracle(sd, O.ModelCfg(n_oar_layer=spec["oar_layers"]), O.SampleCfg.greedy())
    tar_feat, pose, prev = oar_inputs(spec)
    tr = O.FrameTrace()
    with torch.no_grad():
        ids = orc.oar_frame(tar_feat, pose, prev, trace=tr, max_pos=1693)
    assert np.array_equal(ids[6:1030].numpy(), g["map"]) and np.array_equal(ids[1032:1692].numpy(), g["bbox3d"])
    stream = np.array([tr.stream[p] for p in sorted(tr.stream)])
    assert np.array_equal(stream, g["input_stream"][:len(stream)])
    assert len(tr.cleaned_slots) == int(g["n_wipes"]) and len(tr.tar_resampled) == int(g["n_tar_head_calls"])
