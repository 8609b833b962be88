"""End-to-end GPU parity of the engine (TAR encoders + ego head + OAR decode) against rollouts produced by the
UNMODIFIED reference (tests/golden/rollout_*.npz from oracle/make_golden.py).

The engine runs its matrices in fp16 on tensor cores (the reference's own GPU dtype), the goldens come from the
reference's fp32 CPU path, so comparisons are:
  * conditioning feature, ego logits, AR logits: within a stated absolute tolerance
  * greedy token ids: identical wherever the reference's top1-top2 logit gap exceeds MARGIN_TOL; a free-running
    rollout is compared up to its first (necessarily low-margin) divergence, a teacher-forced frame everywhere."""
import os

import numpy as np
import pytest
import torch

from tests._cases import ROLLOUT_CASES, rollout_init
from umgen_b200 import synth
from umgen_b200.config import MODS, ModelConfig, SampleConfig

pytestmark = pytest.mark.gpu

FEAT_ATOL = 3e-2       # conditioning feature (values O(1)), fp16 GEMM chain vs fp32
LOGIT_ATOL = 3e-2      # AR / ego logits
MARGIN_TOL = 6e-2      # ids must agree where the reference's top-2 gap is larger than this


def sampled_positions():
    return list(range(7, 1031)) + list(range(1033, 1693)) + list(range(1695, 2207))


def build(spec):
    from umgen_b200.engine import UMGenEngine
    cfg = ModelConfig.tiny(spec["layers"], cond_frame=max(spec["cond_frames"], spec["input_cond_frames"]))
    sd = synth.make_state_dict(cfg, seed=spec["weight_seed"])
    eng = UMGenEngine(sd, cfg, SampleConfig.greedy())
    eng.keep_trace = True
    eng.want_logits = True
    return eng


def check_frame(tr, g, f, pos, upto=2208):
    feat = tr.tar_feat.cpu().numpy()[::13]
    np.testing.assert_allclose(feat, g["tar_feat"][f], rtol=0, atol=FEAT_ATOL)
    if tr.ego_logits is not None and "ego_logits" in g:
        np.testing.assert_allclose(tr.ego_logits.cpu().numpy(), g["ego_logits"][f], rtol=0, atol=LOGIT_ATOL)
    logits = tr.logits.cpu()
    worst = 0.0
    for i, p in enumerate(pos):
        if p > upto:
            break
        V = 1028 if 1033 <= p <= 1692 else 8192
        top = torch.topk(logits[p - 1, :V], 8).values.numpy()
        worst = max(worst, float(np.abs(top - g["ar_top_vals"][f][i]).max()))
    assert worst < LOGIT_ATOL, f"frame {f}: AR logit error {worst}"
    return worst


@pytest.mark.parametrize("name", ["video_L1", "video_L2", "control_L1", "video_T20_L4", "control_T13_L2", "initmap_L1"])
def test_free_running_rollout_matches_reference(name, golden_dir):
    """Includes the headline window (20 conditioning frames, depth 4 in every stack, the second frame on the look-ahead schedule's
    last-frame path at T = 20), the control working point (13 frames growing inside a 20-frame limit) and a rollout with given pose + map."""
    spec = ROLLOUT_CASES[name]
    g = np.load(os.path.join(golden_dir, f"rollout_{name}.npz"))
    eng = build(spec)
    scene = synth.make_scene(seed=spec["scene_seed"], n_frames=spec["input_frames"])
    init = rollout_init(spec, scene)
    out = eng.inference(spec["new_frames"], spec["cond_frames"], spec["input_cond_frames"], input_cond_tokens=scene,
                        init_tokens=init, control_test=bool(spec.get("control")))
    pos = sampled_positions()
    given_map = "map" in (spec.get("init_mods") or ())
    if given_map:
        pos = [p for p in pos if p > 1031]          # the map block is given: nothing is sampled there
    n_in = spec["input_cond_frames"]
    for f, tr in enumerate(eng.trace):
        gold = np.concatenate([g[f"out_{m}"][0, n_in + f] for m in ("map", "bbox3d", "image")])
        mine = np.concatenate([out[m][0, n_in + f] for m in ("map", "bbox3d", "image")])
        stream = g["input_stream"][f]
        picks = tr.picks.cpu().numpy()[[p - 1 for p in pos]]
        margins = g["ar_top_vals"][f][:, 0] - g["ar_top_vals"][f][:, 1]
        # compare the raw decode streams (what was fed forward), which also covers later-wiped slots
        bad = np.nonzero(picks != stream)[0]
        wiped = np.nonzero(tr.tokens.cpu().numpy()[[p - 1 for p in pos]] != picks)[0]
        first_bad = pos[int(bad[0])] if bad.size else 2208
        worst = check_frame(tr, g, f, pos, upto=first_bad)
        print(f"{name} frame {f}: stream identical through position {first_bad - 1}; worst AR logit err {worst:.2e}; "
              f"{len(wiped)} ids rewritten by wipes; status {tr.status[:4]}")
        if bad.size:
            i = int(bad[0])
            assert margins[i] < MARGIN_TOL, f"frame {f}: diverged at position {pos[i]} where the reference margin is {margins[i]:.3f}"
            break
        assert np.array_equal(mine, gold), f"frame {f}: output ids differ although the decode stream matches"
        assert np.array_equal(out["pose"][0, n_in + f], g["out_pose"][0, n_in + f])
    if name == "video_T20_L4":
        assert eng._la is not None and eng._la["T"] == 20, "the second frame must have run on the look-ahead schedule at T = 20"


def test_teacher_forced_frame_matches_reference_everywhere(golden_dir):
    name = "video_L1"
    spec = ROLLOUT_CASES[name]
    g = np.load(os.path.join(golden_dir, f"rollout_{name}.npz"))
    eng = build(spec)
    scene = synth.make_scene(seed=spec["scene_seed"], n_frames=spec["input_frames"])
    n_in = spec["input_cond_frames"]
    cond = {m: scene[m][0, :n_in].clone() for m in MODS}
    pos = sampled_positions()
    teacher = torch.zeros(2207, dtype=torch.int64)
    teacher[[p - 1 for p in pos]] = torch.from_numpy(g["input_stream"][0].astype(np.int64))
    eng.frame(cond, None, False, teacher=teacher)
    tr = eng.trace[0]
    check_frame(tr, g, 0, pos)
    picks = eng.dec.picks.cpu().numpy()[[p - 1 for p in pos]]
    margins = g["ar_top_vals"][0][:, 0] - g["ar_top_vals"][0][:, 1]
    confident = margins > MARGIN_TOL
    assert confident.sum() > 500
    diff = np.nonzero((picks != g["input_stream"][0]) & confident)[0]
    assert diff.size == 0, f"{diff.size} confident positions differ, first at {pos[int(diff[0])]}"
    print(f"teacher-forced: {int(confident.sum())} confident positions identical; "
          f"{int((picks != g['input_stream'][0]).sum())} low-margin differences among {len(pos)}")


def test_dropin_module_reproduces_reference_rollout(golden_dir):
    """projects.models.UMGen.UMGen (the drop-in module) driven like tools/model_pl.py drives the reference."""
    from projects.models.UMGen import UMGen
    from tests.test_dropin_surface import eval_namespace
    name = "video_L1"
    spec = ROLLOUT_CASES[name]
    g = np.load(os.path.join(golden_dir, f"rollout_{name}.npz"))
    model = UMGen(eval_namespace(layers=spec["layers"], cond_frame=spec["cond_frames"], top_k=1, top_k_map=1)).eval()
    model.load_state_dict(synth.make_state_dict(ModelConfig.tiny(spec["layers"]), seed=spec["weight_seed"]), strict=False)
    model.sample_param_map = 1
    model.topk_image = 1            # greedy recipe of SURVEY.md section 3.4
    model.cuda()
    scene = synth.make_scene(seed=spec["scene_seed"], n_frames=spec["input_frames"])
    out = model.inference(new_frames=spec["new_frames"], cond_frames=spec["cond_frames"], input_cond_frames=spec["input_cond_frames"],
                          pred_task="pose_map_bbox3d_image", input_cond_tokens=scene, init_tokens=None, cond_on_par=True, infer_from_gt=False)
    for m in MODS:
        assert out[m].dtype == np.int64 and out[m].shape == g[f"out_{m}"].shape
    # identical to the reference rollout up to the first position where the reference itself is undecided (top-2 gap below the fp16 tolerance:
    # frame 1 of this golden has gaps of 7e-5 and 1e-4 near its end); the frames before that position must match exactly
    n_in = spec["input_cond_frames"]
    pos = sampled_positions()
    for m in MODS:
        assert np.array_equal(out[m][:, :n_in], g[f"out_{m}"][:, :n_in])
    for f in range(spec["new_frames"]):
        mine = np.concatenate([out[m][0, n_in + f] for m in ("map", "bbox3d", "image")])
        gold = np.concatenate([g[f"out_{m}"][0, n_in + f] for m in ("map", "bbox3d", "image")])
        assert np.array_equal(out["pose"][0, n_in + f], g["out_pose"][0, n_in + f])
        bad = np.nonzero(mine != gold)[0]
        if bad.size:
            margins = g["ar_top_vals"][f][:, 0] - g["ar_top_vals"][f][:, 1]
            i = int(bad[0])
            assert margins[i] < MARGIN_TOL, f"frame {f}: differs from the reference at position {pos[i]} where its margin is {margins[i]:.4f}"
            print(f"frame {f}: first difference at position {pos[i]} (reference margin {margins[i]:.2e}); {bad.size} ids differ in this frame")
            break


def test_box_pass_beside_the_decode_kernel_changes_nothing(golden_dir):
    """engine.overlap runs the box_tar pass on a second stream while the decode kernel already works on the map block (tar_ready_i32 in
    include/umgen.h): same arithmetic, so features, logits and ids must be bit-identical to the sequential schedule."""
    spec = ROLLOUT_CASES["video_L2"]
    scene = synth.make_scene(seed=spec["scene_seed"], n_frames=spec["input_frames"])
    runs = []
    for overlap in (False, True):
        eng = build(spec)
        if eng.dec.kernel_name != "decode_cluster_kernel":
            pytest.skip("the overlapped schedule needs the 8-cluster decode kernel")
        eng.overlap = overlap
        eng.inference(1, spec["cond_frames"], spec["input_cond_frames"], input_cond_tokens=scene)      # an engine's first frame is always sequential
        eng.trace.clear()
        eng.frame_counter = 0
        out = eng.inference(spec["new_frames"], spec["cond_frames"], spec["input_cond_frames"], input_cond_tokens=scene)
        tr = eng.trace[0]
        runs.append((out, tr.tar_feat.cpu(), tr.logits.cpu(), tr.picks.cpu(), tr.status[:4]))
    (o0, f0, l0, p0, s0), (o1, f1, l1, p1, s1) = runs
    assert s0[0] == 0 and s1[0] == 0
    assert torch.equal(f0, f1), "conditioning feature differs between the two schedules"
    assert torch.equal(p0, p1) and torch.equal(l0, l1)
    for m in MODS:
        assert np.array_equal(o0[m], o1[m]), m


def test_lookahead_schedule_changes_nothing(golden_dir):
    """engine.lookahead computes frames 0..T-2 of the next window beside the decode kernel and only the last frame afterwards: per-frame
    spatial attention and causal temporal attention make that the same arithmetic, so every feature, logit and id must be bit-identical."""
    spec = ROLLOUT_CASES["video_L2"]
    scene = synth.make_scene(seed=spec["scene_seed"], n_frames=spec["input_frames"])
    runs = []
    for la in (False, True):
        eng = build(spec)
        eng.lookahead = la
        eng.overlap = False
        out = eng.inference(4, 3, 3, input_cond_tokens=scene)          # 4 new frames, window of 3: slides from the second frame on
        assert len(eng.trace) == 4
        runs.append((out, [(tr.tar_feat.cpu(), tr.logits.cpu(), tr.picks.cpu(), tr.ego_logits.cpu()) for tr in eng.trace], eng))
    (o0, t0, _), (o1, t1, e1) = runs
    assert e1._la is not None and e1._la["T"] == 3
    for f, (a, b) in enumerate(zip(t0, t1)):
        for k, name in enumerate(("conditioning feature", "AR logits", "decode stream", "ego logits")):
            assert torch.equal(a[k], b[k]), f"frame {f}: {name} differs between the schedules"
    for m in MODS:
        assert np.array_equal(o0[m], o1[m]), m


@pytest.mark.parametrize("case", ["slide_T20", "grow_13_to_16_control"])
def test_lookahead_schedule_changes_nothing_at_the_headline_window(case):
    """The same bit-identity at the sizes the benchmark runs: a full 20-frame window that slides (T = 20 last-frame path, CUDA-graph replay,
    three-stream suffix), and the control working point whose window grows 13 -> 16 inside a 20-frame limit (UMGen.py:1600-1603, infer_fun.py:64-71)."""
    from umgen_b200.engine import UMGenEngine
    cfg = ModelConfig.tiny(2, cond_frame=20)
    sd = synth.make_state_dict(cfg, seed=31)
    control = case != "slide_T20"
    n_in, new = (13, 3) if control else (20, 3)
    scene = synth.make_scene(seed=12, n_frames=n_in)
    init = synth.make_control(seed=12, n_frames=new) if control else None
    runs = []
    for la in (False, True):
        eng = UMGenEngine(sd, cfg, SampleConfig.greedy())
        eng.keep_trace, eng.want_logits = True, True
        eng.lookahead, eng.overlap = la, False
        out = eng.inference(new, 20, n_in, input_cond_tokens=scene, init_tokens=init, control_test=control)
        assert len(eng.trace) == new
        runs.append((out, [(tr.tar_feat.cpu(), tr.logits.cpu(), tr.picks.cpu()) for tr in eng.trace], eng))
    (o0, t0, _), (o1, t1, e1) = runs
    assert e1._la is not None and e1._la["T"] == (min(n_in + new, 20))
    for f, (a, b) in enumerate(zip(t0, t1)):
        for k, name in enumerate(("conditioning feature", "AR logits", "decode stream")):
            assert torch.equal(a[k], b[k]), f"{case} frame {f}: {name} differs between the schedules"
    for m in MODS:
        assert np.array_equal(o0[m], o1[m]), m


def test_lookahead_falls_back_when_the_window_does_not_continue(golden_dir):
    """A window that is not the previous one + the returned frame must be recomputed in full (frame() compares on the host)."""
    spec = ROLLOUT_CASES["video_L1"]
    scene = synth.make_scene(seed=spec["scene_seed"], n_frames=4)
    eng = build(spec)
    eng.window = 2
    cond = {m: scene[m][0, :2].clone() for m in MODS}
    eng.frame(cond)
    other = {m: scene[m][0, 2:4].clone() for m in MODS}                # unrelated window of the same length
    assert not eng._continues(other)
    eng.frame(other)
    ref = build(spec)
    ref.lookahead = False
    ref.frame({m: scene[m][0, 2:4].clone() for m in MODS})
    assert torch.equal(eng.trace[1].tar_feat.cpu(), ref.trace[0].tar_feat.cpu())
    assert torch.equal(eng.trace[1].picks.cpu(), ref.trace[0].picks.cpu())


# ---- several scenes per GPU (SceneBatchEngine) -----------------------------------------------------------------------
def _two_scene_tokens(spec):
    a = synth.make_scene(seed=spec["scene_seed"], n_frames=spec["input_frames"])
    b = synth.make_scene(seed=spec["scene_seed"] + 50, n_frames=spec["input_frames"])
    return a, b, {m: torch.cat([a[m], b[m]], dim=0) for m in MODS}


@pytest.mark.parametrize("name", ["video_L2", "video_T20_L4", "control_L1"])
def test_two_scenes_per_gpu_reproduce_the_single_scene_rollouts(name, golden_dir):
    """Scene 0 is the golden scene (so the batched rollout is pinned to the reference too), scene 1 another one; both must come out exactly
    as from one-scene engines -- through the sliding window, the look-ahead schedule (frames after the first) and the control overwrite."""
    from umgen_b200.engine import SceneBatchEngine, UMGenEngine
    spec = ROLLOUT_CASES[name]
    g = np.load(os.path.join(golden_dir, f"rollout_{name}.npz"))
    cfg = ModelConfig.tiny(spec["layers"], cond_frame=max(spec["cond_frames"], spec["input_cond_frames"]))
    sd = synth.make_state_dict(cfg, seed=spec["weight_seed"])
    sa, sb, both = _two_scene_tokens(spec)
    init_a, init_b = rollout_init(spec, sa), rollout_init(spec, sb)
    init_both = None if init_a is None else {m: torch.cat([init_a[m], init_b[m]], dim=0) for m in init_a}
    new = spec["new_frames"] + (1 if name == "video_L2" else 0)
    kw = dict(control_test=bool(spec.get("control")))
    single = []
    for sc, ini in ((sa, init_a), (sb, init_b)):
        eng = UMGenEngine(sd, cfg, SampleConfig.greedy())
        if eng.dec.kernel_name != "decode_cluster_kernel":
            pytest.skip("needs the 8-cluster decode kernel")
        single.append(eng.inference(new, spec["cond_frames"], spec["input_cond_frames"], input_cond_tokens=sc, init_tokens=ini, **kw))
        del eng
    beng = SceneBatchEngine(sd, cfg, SampleConfig.greedy(), scenes=2)
    out = beng.inference(new, spec["cond_frames"], spec["input_cond_frames"], input_cond_tokens=both, init_tokens=init_both, **kw)
    for k in range(2):
        for m in MODS:
            assert out[m].shape[0] == 2 and np.array_equal(out[m][k], single[k][m][0]), f"scene {k} {m}: batched rollout differs from the one-scene rollout"
    n_in = spec["input_cond_frames"]
    # first generated frame of scene 0 against the reference itself: identical up to the first position the reference leaves undecided
    # (positions given by init tokens / wiped slots aside, which the one-scene tests cover)
    if not spec.get("control") and not spec.get("init_mods"):
        mine = np.concatenate([out[m][0, n_in] for m in ("map", "bbox3d", "image")])
        gold = np.concatenate([g[f"out_{m}"][0, n_in] for m in ("map", "bbox3d", "image")])
        bad = np.nonzero(mine != gold)[0]
        if bad.size:
            margins = g["ar_top_vals"][0][:, 0] - g["ar_top_vals"][0][:, 1]
            assert margins[int(bad[0])] < MARGIN_TOL, f"scene 0 differs from the reference at a position with margin {margins[int(bad[0])]:.4f}"
    if new > 1:
        assert all(e._la is not None for e in beng.engines), "frames after the first must have run on the look-ahead schedule"


def test_two_scenes_per_gpu_sample_from_their_own_streams():
    from umgen_b200.engine import SceneBatchEngine, UMGenEngine
    spec = ROLLOUT_CASES["video_L1"]
    cfg = ModelConfig.tiny(spec["layers"], cond_frame=max(spec["cond_frames"], spec["input_cond_frames"]))
    sd = synth.make_state_dict(cfg, seed=spec["weight_seed"])
    sa, _, _ = _two_scene_tokens(spec)
    same = {m: torch.cat([sa[m], sa[m]], dim=0) for m in MODS}          # the SAME scene twice: only the random streams differ
    sample = SampleConfig(top_k=5, top_k_map=5, top_k_image=16, seed=21)
    beng = SceneBatchEngine(sd, cfg, sample, scenes=2)
    out = beng.inference(2, spec["cond_frames"], spec["input_cond_frames"], input_cond_tokens=same)
    n_in = spec["input_cond_frames"]
    assert not np.array_equal(out["map"][0, n_in], out["map"][1, n_in]), "two scenes of a launch must not share their random stream"
    for k in range(2):
        import dataclasses
        eng = UMGenEngine(sd, cfg, dataclasses.replace(sample, seed=21 + k))
        ref = eng.inference(2, spec["cond_frames"], spec["input_cond_frames"], input_cond_tokens=sa)
        for m in MODS:
            assert np.array_equal(out[m][k], ref[m][0]), f"scene {k} {m}: differs from a one-scene engine with seed {21 + k}"
        del eng
