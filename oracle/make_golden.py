"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(imported from /root/reference through oracle/ref_import.py) on seeded synthetic assets.
TEST INFRASTRUCTURE ONLY; run in the build container:  python -m oracle.make_golden

Files written (all small):
  tables.npz      pose/box value LUTs from the reference tokenizers+normalisers, sha256 of the fixed
                  sinusoid tables of a reference model instance, d_token_pos / pos_mod maps
  collision.npz   BoxOverlap.check_collision answers for seeded random box lists (inputs are
                  regenerated from the seed by tests/_cases.py)
  rollout_*.npz   UMGen.inference outputs for tiny-depth models: token ids, the OAR conditioning
                  feature (sub-sampled), every AR-head logit row reduced to (top-8 values, ids),
                  ego logits
"""
from __future__ import annotations

import hashlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import as R                      # noqa: E402
from umgen_b200 import synth                            # noqa: E402
from umgen_b200.config import ModelConfig               # noqa: E402
from tests._cases import collision_cases, ROLLOUT_CASES, OAR_CASES, oar_inputs, apply_tweak, vq_codes, rollout_init, DATASET_CASES, raw_scene, RUNNER_CASES, runner_batch, summarise_inference_kwargs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().view(torch.uint8).numpy().tobytes()).hexdigest()


def tables():
    ns = R.load()
    cfg = R.reference_config(layers=1)
    toks = np.arange(1024)
    pose_lut = np.stack([
        cfg.ego_norm.unnormalize_ego(cfg.ego_tokenlizer.decode(np.stack([toks] * 3, axis=1).copy()))[:, c]
        for c in range(3)], axis=1).astype(np.float32)
    box_lut = np.zeros((1028, 10))
    for t in range(1028):
        tok = np.full((1, 1, 11), t, dtype=np.int64)
        b, _ = cfg.box3d_tokenlizer.decode_single_objects(tok)
        box_lut[t] = cfg.agent_norm.unnormalize_bbox3d(b[None, ...])[0][0]
    model = R.build_reference_model(cfg)
    sd = model.state_dict()
    fixed = {k: sha(sd[k]) for k in ("fouier_pe", "bbox3d_spatial_posi", "grid_center_posi_embedding")}
    mods = cfg.task["pose_map_bbox3d_image"]
    dpos = model.d_token_pos(mods)
    posmod = np.array([mods.index(model.pos_mod(p, mods)) for p in range(1, 2208)], dtype=np.int8)
    np.savez_compressed(os.path.join(OUT, "tables.npz"), pose_lut=pose_lut, box_lut=box_lut,
                        fixed_keys=np.array(list(fixed.keys())), fixed_sha=np.array(list(fixed.values())),
                        dpos_keys=np.array(list(dpos.keys())), dpos_vals=np.array(list(dpos.values())),
                        pos_mod=posmod, pad_token=cfg.box3d_tokenlizer.pad_token,
                        bbox_vocab=len(cfg.box3d_tokenlizer), n_params_1layer=sum(p.numel() for p in model.parameters()))
    print("tables.npz done")


def collision():
    ns = R.load()
    ov = ns.misc.BoxOverlap()
    cases = collision_cases()
    ans = np.array([bool(ov.check_collision([b for b in boxes], fliter=True)) for boxes in cases])
    np.savez_compressed(os.path.join(OUT, "collision.npz"), answers=ans)
    print("collision.npz done:", len(cases), "cases,", int(ans.sum()), "collide")


def record_input_stream(model, sink):
    """Wrap UMGen.sample_next_token (UMGen.py:1029-1139): after every sampled (non-forced) position record the token
    that was appended to res_tokens at that moment, i.e. the id whose embedding feeds the next decode step.  Later
    collision wipes rewrite res_tokens but not these inputs (the KV cache stays stale, UMGen.py:1356-1377)."""
    orig = model.sample_next_token

    def wrapped(curr_seq_len, curr_emb, mod, out_tokens, res_tokens, d_token_pos, *a, **k):
        r = orig(curr_seq_len, curr_emb, mod, out_tokens, res_tokens, d_token_pos, *a, **k)
        if curr_seq_len not in d_token_pos:
            sink.append(int(r[2][mod][-1].reshape(-1)[0]))
        return r

    model.sample_next_token = wrapped


def rollout(name: str, spec: dict):
    cfg = ModelConfig.tiny(spec["layers"])
    ref_cfg = R.reference_config(layers=spec["layers"], cond_frame=spec["cond_frames"])
    sd = synth.make_state_dict(cfg, seed=spec["weight_seed"])
    model = R.build_reference_model(ref_cfg, sd, greedy=True)
    scene = synth.make_scene(seed=spec["scene_seed"], n_frames=spec["input_frames"])
    init = rollout_init(spec, scene)

    cap = {"tar_feat": [], "ego": [], "ar": [], "tar_bbox": []}
    stream = []
    record_input_stream(model, stream)
    orig_oar = model.infer_oar_net

    def wrapped(tar_emb, *a, **k):
        mods = model.task[k.get("pred_task", "pose_map_bbox3d_image")]
        cap["tar_feat"].append(torch.cat([tar_emb[m] for m in mods], dim=-2)[0, -1].float().clone())
        cap["ar"].append([])
        cap["tar_bbox"].append([])
        return orig_oar(tar_emb, *a, **k)

    model.infer_oar_net = wrapped
    tr = model.transformer
    tr.head_ego.register_forward_hook(lambda m, i, o: cap["ego"].append(o[0, -1].float().clone()))
    for hname in ("head_ar_map", "head_ar_bbox3d", "head_ar_img"):
        getattr(tr, hname).register_forward_hook(
            lambda m, i, o: cap["ar"][-1].append(o[0, -1, -1].float().clone()))
    tr.head_tar_bbox3d.register_forward_hook(
        lambda m, i, o: cap["tar_bbox"][-1].append(o[0, -1, -1].float().clone()))

    t0 = time.time()
    out = model.inference(new_frames=spec["new_frames"], cond_frames=spec["cond_frames"],
                          input_cond_frames=spec["input_cond_frames"], pred_task="pose_map_bbox3d_image",
                          input_cond_tokens={k: v.clone() for k, v in scene.items()},
                          init_tokens=init, control_test=bool(spec.get("control")),
                          cond_on_par=True, infer_from_gt=False)
    print(f"{name}: reference ran in {time.time() - t0:.1f}s")
    save = {f"out_{m}": out[m] for m in out}
    nf = spec["new_frames"]
    assert len(cap["tar_feat"]) == nf
    save["tar_feat_rows"] = np.arange(0, 2207, 13)
    save["tar_feat"] = np.stack([f[::13].numpy() for f in cap["tar_feat"]]).astype(np.float32)
    save["tar_feat_absmean"] = np.array([float(f.abs().mean()) for f in cap["tar_feat"]], dtype=np.float32)
    if cap["ego"]:
        save["ego_logits"] = np.stack([e.numpy() for e in cap["ego"]]).astype(np.float32)
    topv, topi = [], []
    for fr in cap["ar"]:
        assert len(fr) == 1024 + 660 + 512 - (1024 if "map" in (spec.get("init_mods") or ()) else 0), len(fr)
        v = [torch.topk(l, 8) for l in fr]
        topv.append(np.stack([x.values.numpy() for x in v]))
        topi.append(np.stack([x.indices.numpy() for x in v]))
    save["ar_top_vals"] = np.stack(topv).astype(np.float32)          # [frames, 2196, 8]
    save["ar_top_ids"] = np.stack(topi).astype(np.int32)
    save["n_tar_bbox_calls"] = np.array([len(x) for x in cap["tar_bbox"]])
    n_sampled = 2196 - (1024 if "map" in (spec.get("init_mods") or ()) else 0)
    assert len(stream) == nf * n_sampled, len(stream)
    save["input_stream"] = np.array(stream, dtype=np.int32).reshape(nf, n_sampled)
    np.savez_compressed(os.path.join(OUT, f"rollout_{name}.npz"), **save)


def oar_case(name: str, spec: dict):
    """UMGen.infer_oar_net (UMGen.py:1151-1273) alone, on a seeded random conditioning feature."""
    import dataclasses
    cfg = dataclasses.replace(ModelConfig.tiny(1), n_oar_layer=spec["oar_layers"])
    ref_cfg = R.reference_config(layers=1, n_oar_layer=spec["oar_layers"])
    sd = apply_tweak(synth.make_state_dict(cfg, seed=spec["weight_seed"]), spec.get("tweak"))
    model = R.build_reference_model(ref_cfg, sd, greedy=True)
    tar_feat, pose, prev_bbox = oar_inputs(spec)
    off = {"pose": 0, "map": 5, "bbox3d": 1031, "image": 1693}
    ln = {"pose": 5, "map": 1026, "bbox3d": 662, "image": 514}
    tar_emb = {m: tar_feat[off[m]:off[m] + ln[m]][None, None].clone() for m in off}
    cap = []
    for hname in ("head_ar_map", "head_ar_bbox3d", "head_ar_img"):
        getattr(model.transformer, hname).register_forward_hook(lambda m, i, o: cap.append(o[0, -1, -1].float().clone()))
    ntar = []
    model.transformer.head_tar_bbox3d.register_forward_hook(lambda m, i, o: ntar.append(1))
    stream = []
    record_input_stream(model, stream)
    wipes = []                        # slots rewritten to <pad> by rule_based_constraint (UMGen.py:1356-1377)
    orig_rule = model.rule_based_constraint

    def counted_rule(current_token, *a, **k):
        before = int(torch.as_tensor(current_token).reshape(-1)[0])
        r = orig_rule(current_token, *a, **k)
        if int(torch.as_tensor(r).reshape(-1)[0]) == 1027 and before != 1027:
            wipes.append(len(model.decoded_bbox))
        return r

    model.rule_based_constraint = counted_rule
    control = None if spec["control_slot"] is None else (np.array([spec["control_slot"]]),)
    t0 = time.time()
    with torch.no_grad():
        res = model.infer_oar_net(tar_emb, None, pred_task="pose_map_bbox3d_image",
                                  init_tokens={"pose": pose.view(1, 1, 3).clone()},
                                  previous_frame_tokens={"bbox3d": prev_bbox.view(1, 1, 660)},
                                  control_objects=control, max_objects=100)
    print(f"{name}: reference infer_oar_net ran in {time.time() - t0:.1f}s; tar-head calls {len(ntar)}")
    v = [torch.topk(l, 8) for l in cap]
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"),
                        map=res["map"].view(-1).numpy(), bbox3d=res["bbox3d"].view(-1).numpy(),
                        image=res["image"].view(-1).numpy(), pose=res["pose"].view(-1).numpy(),
                        top_vals=np.stack([x.values.numpy() for x in v]).astype(np.float32),
                        top_ids=np.stack([x.indices.numpy() for x in v]).astype(np.int32),
                        n_tar_head_calls=len(ntar), input_stream=np.array(stream, dtype=np.int32),
                        n_wipes=len(wipes), boxes_at_wipe=np.array(wipes, dtype=np.int32))


def vq_case(kind: str):
    """NormVQModel.decode_code (tokenizer/vq_model.py:92-96) and, for the map, tools/decode_map.py:to_rgb, run by the
    unmodified reference modules on the seeded synthetic VQ checkpoint; outputs stored sub-sampled (every 4th pixel)."""
    import logging
    logging.disable(logging.WARNING)
    R.load()
    torch.save({"state_dict": {}}, "/tmp/umgen_empty_vq.ckpt")
    with R.reference_cwd():
        from projects.tokenizer import vq_model as V
    fac = V.get_map_normvq_dim16_res256_f8 if kind == "map" else V.get_normvq_dim16_res512_f16
    sd = synth.make_vq_state_dict(kind, seed=1)
    m = fac(device="cpu", ckpt="/tmp/umgen_empty_vq.ckpt").eval()
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and not [k for k in res.missing_keys if k.startswith(("decoder", "post_quant"))]
    h, w = (32, 32) if kind == "map" else (16, 32)
    code = vq_codes(kind)
    with torch.no_grad():
        out = m.decode_code(code)
    save = {"out": out[:, :, ::4, ::4].numpy().astype(np.float32), "absmean": float(out.abs().mean())}
    if kind == "map":
        # decode_map.to_rgb without importing decode_map (it needs cv2 video writers): same three lines, reference semantics
        state = torch.random.get_rng_state()
        torch.manual_seed(0)
        wgt = torch.randn(3, out.shape[1], 1, 1)
        torch.random.set_rng_state(state)
        rgb = torch.nn.functional.conv2d(out, wgt)
        rgb = 2.0 * (rgb - rgb.min()) / (rgb.max() - rgb.min()) - 1.0
        save["rgb"] = rgb[:, :, ::4, ::4].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, f"vq_{kind}.npz"), **save)
    print(f"vq_{kind}.npz done", out.shape)


def postprocess():
    """The bbox3d / pose value decode of tools/model_pl.py:decode_tokens (:262-291) run by the reference's own tokenizers and normalisers on a
    seeded scene whose bbox3d stream also holds out-of-place ids (category ids in attribute positions, attribute ids in category positions)."""
    from umgen_b200 import synth
    R.load()
    cfg = R.reference_config(layers=1)
    scene = synth.make_scene(seed=11, n_frames=6)
    bt = scene["bbox3d"][0].numpy().astype(np.int64).copy()
    rs = np.random.RandomState(5)
    idx = rs.randint(0, bt.size, size=200)
    bt.reshape(-1)[idx] = rs.randint(0, 1028, size=200)
    pt = scene["pose"][0].numpy().astype(np.int64).copy()
    tk = cfg.box3d_tokenlizer
    work = bt.copy()
    pad_mask = work == tk.pad_token
    work[~pad_mask] = np.clip(work[~pad_mask], tk.start, tk.start + tk.vocab_size - 1)
    bboxes, classes = tk.decode(work, keep_order=True, no_special=True)
    bboxes = cfg.agent_norm.unnormalize_bbox3d(bboxes)
    pose = cfg.ego_norm.unnormalize_ego(cfg.ego_tokenlizer.decode(pt.copy()))
    names = ["none", "vehicle", "bicycle", "pedestrian"]
    # the ground-truth side (model_pl.py:278-286): decode without keep_order drops every slot that holds a <pad> and every out-of-range category.
    # Category ids in attribute positions stay (they clip to the last bin); attribute ids in category positions are what the filter removes.
    gt = scene["bbox3d"][0].numpy().astype(np.int64).copy()
    gidx = rs.randint(0, gt.size, size=60)
    gt.reshape(-1)[gidx] = rs.randint(0, 1027, size=60)
    anno, anno_cls = tk.decode(gt.copy(), no_special=True)
    anno = cfg.agent_norm.unnormalize_bbox3d(anno)
    anno_n = np.array([len(a) for a in anno], dtype=np.int64)
    anno_flat = np.concatenate([np.asarray(a, dtype=np.float64).reshape(-1, 10) for a in anno], axis=0)
    anno_cls_flat = np.array([names.index(c) for row in anno_cls for c in row], dtype=np.int8)
    np.savez_compressed(os.path.join(OUT, "postprocess.npz"), bbox_tokens=bt, pose_tokens=pt, bboxes=np.stack([np.asarray(b) for b in bboxes]),
                        classes=np.array([[names.index(c) for c in row] for row in classes], dtype=np.int8), pose_values=np.asarray(pose),
                        gt_bbox_tokens=gt, anno_n=anno_n, anno_boxes=anno_flat, anno_classes=anno_cls_flat)
    print("postprocess.npz done")


def dataset():
    """The reference's NuPlanTokenDataset (plugin/data/datasets/UMGen_nuplan_dataset.py) with the evaluation transform list
    (configs/UMGen_config_evaluation.py:247-255) as tools/infer_fun.py:189-213 configures it, run on the synthetic raw scenes of tests/_cases.py."""
    import pickle
    import tempfile
    R.load()
    cfg = R.reference_config(layers=1)
    with R.reference_cwd():
        from projects.plugin.data.datasets.UMGen_nuplan_dataset import NuPlanTokenDataset
        from projects.plugin.data.transforms.common import MergeAttribute, SplitAttriute
        from projects.plugin.data.transforms.normalize import ToTensor
    data_key = cfg.agent_norm.data_key
    out = {}
    for name, (seed, n, block, gap, n_tracks) in DATASET_CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, f"synthetic_scene_{seed:04d}_clip_a.pkl")
            with open(path, "wb") as f:
                pickle.dump(raw_scene(seed, n, n_tracks), f)
            tf = [SplitAttriute(input_key=["bbox3d"], target_key=[data_key]), cfg.agent_norm,
                  MergeAttribute(input_key=["bbox3d"], target_key=[data_key], merage_name=["bbox3d"]), cfg.ego_norm, cfg.box3d_tokenlizer,
                  cfg.ego_tokenlizer, ToTensor()]
            with R.reference_cwd():
                ds = NuPlanTokenDataset(data_root=[tmp], training=False, block_size=block, views=["CAM_F0"],
                                        categories_file="projects/configs/category.txt", sampling_gap=gap, transforms=tf, inference_flag=True,
                                        start_index=10, sp_list=None, sample_img=True, return_scene_name=True, control_test=False,
                                        long_sceniors=False, img_transform=None, return_ori_image=False)
                d = ds[0]
        assert sorted(d) == ["bbox3d", "file_name", "image", "map", "pose", "pose_diff"], sorted(d)
        for k in ("pose", "map", "bbox3d", "image", "pose_diff"):
            out[f"{name}.{k}"] = d[k].numpy()
        print(name, {k: tuple(d[k].shape) for k in d if k != "file_name"}, "live slots per frame", (d["bbox3d"].numpy().reshape(-1, 60, 11)[:, :, 0] != 1027).sum(1)[:12])
    np.savez_compressed(os.path.join(OUT, "dataset.npz"), **out)
    print("dataset.npz done")


def runner():
    """What the reference's own UMGen_PL.world_model_evaluate / generate_init_tokens (tools/model_pl.py:95-262) hand to model.inference, recorded by a
    stand-in model: the method bodies run unmodified on an instance whose constructor is bypassed (it would load the decoders' checkpoints and build
    the visualiser); decode_tokens / generate_videos are stubbed out, save_tokens is the reference's."""
    import json
    import tempfile
    import types
    R.load()
    sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))          # pytorch_lightning / matplotlib stand-ins (test infrastructure)
    with R.reference_cwd():
        import projects.tools.model_pl as ref_pl
    out = {}
    for name, (task, new_frames, init_mod, _, _) in RUNNER_CASES.items():
        calls = []

        class Recorder:
            def inference(self, **kw):
                calls.append(kw)
                return {m: np.zeros((1, 1, n), dtype=np.int64) for m, n in (("pose", 3), ("map", 1024), ("bbox3d", 660), ("image", 512))}

        with tempfile.TemporaryDirectory() as tmp:
            pl = ref_pl.UMGen_PL.__new__(ref_pl.UMGen_PL)
            pl.__dict__.update(dict(
                inference_setting_dict=dict(new_frames=new_frames, cond_frames=20, pred_task="pose_map_bbox3d_image", cond_on_par=True, infer_from_gt=False),
                control_test="control" in task, init_token_mod=init_mod, input_cond_frames=20, token_save_path=tmp, model=Recorder(),
                visulizer=types.SimpleNamespace(spe_text="x"), generate_video_flag=False))
            pl.decode_tokens = lambda *a, **k: (None,) * 7
            pl.generate_videos = lambda *a, **k: None
            ref_pl.UMGen_PL.global_rank = 0
            pl.world_model_evaluate(runner_batch(name), 0)
            saved = sorted(os.listdir(tmp))
        assert len(calls) == 1
        out[name] = {"kwargs": summarise_inference_kwargs(calls[0]), "saved": saved}
        print(name, {k: v for k, v in out[name]["kwargs"].items() if not isinstance(v, dict)}, saved)
    with open(os.path.join(OUT, "runner.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("runner.json done")


def samplers():
    """The reference's own token samplers -- UMGen.topk (UMGen.py:899-913) and UMGen.sample_top_p (:915-965), the two values `token_sampler` takes
    (:118-126) -- called on seeded logits under torch.manual_seed, plus the parameters the model hands them per modality (:1004, 1063, 1073, 1133)."""
    from tests._cases import SAMPLER_CASES, sampler_logits
    R.load()
    out = {}
    models = {}
    for method in ("topk", "topp"):
        cfg = R.reference_config(layers=1, sample_method=method)
        m = R.build_reference_model(cfg, None)
        models[method] = m
        assert m.token_sampler.__name__ == {"topk": "topk", "topp": "sample_top_p"}[method]
        out[f"params_{method}"] = np.array([float(m.sample_param), float(m.sample_param_map), float(m.topk_image), float(m.sfmx_temp)])
    for name, method, param, vocab, rows, scale, seed in SAMPLER_CASES:
        m = models[method]
        x = sampler_logits(vocab, rows, scale, seed)
        torch.manual_seed(seed)
        pick = m.token_sampler(x[None].clone(), param)
        assert tuple(pick.shape) == (1, 1, rows), pick.shape
        out[name] = pick[0, 0].numpy().astype(np.int64)
        print(name, out[name][:8])
    np.savez_compressed(os.path.join(OUT, "samplers.npz"), **out)
    print("samplers.npz done")


def visualize():
    """The reference's own Visulizer (tools/visulize.py) run the way UMGen_PL.generate_videos / generate_compare_videos run it (model_pl.py:61-73,
    283-331) on the seeded inputs of tests/_cases.py: sha256 of every composed frame (the list handed to generate_img_and_video), of the mp4 it
    writes and of the "pred" video.  One environment patch: this image's OpenCV is headless and the reference calls cv2.destroyAllWindows()
    before it releases the writer."""
    import hashlib
    import json
    import tempfile
    import cv2
    from tests._cases import VISUALIZE_CASES, visualize_inputs
    R.load()
    sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))          # matplotlib stand-in (imported, never called)
    with R.reference_cwd():
        import projects.tools.visulize as ref_vis
    cv2.destroyAllWindows = lambda: None
    out = {"cv2": cv2.__version__, "cases": {}}
    old = os.getcwd()
    for name in VISUALIZE_CASES:
        d = visualize_inputs(name)
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)                                             # the reference keeps its frame cache under ./output/tmp_cache
            try:
                vis = ref_vis.Visulizer(video_save_path=os.path.join(tmp, "videos/"), video_pretext="UMGen", width=d["width"], height=d["width"],
                                        project_name="UMGen_infer", spe_text="synthetic_video", save_video=True, addtion_ego=True,
                                        bbox3d_arrow_length_scale=1, cond_frames=d["cond_frames"], put_text=d["put_text"])
                seen = {}
                inner = vis.generate_img_and_video

                def spy(images_all, *a, _inner=inner, _seen=seen, **k):
                    _seen.setdefault("frames", []).append([np.ascontiguousarray(f).copy() for f in images_all])
                    return _inner(images_all, *a, **k)

                vis.generate_img_and_video = spy
                anno = None
                if d.get("anno_boxes") is not None:
                    anno = np.empty(len(d["anno_boxes"]), dtype=object)
                    for i, a in enumerate(d["anno_boxes"]):
                        anno[i] = np.array(a, dtype=np.float64).copy()
                vis.visulize(box=np.array([b.copy() for b in d["boxes"]], dtype=object), anno_box=anno, scene_name=d["scene_name"], pose=d["pose"].copy(),
                             real_pose=None if d["real_pose"] is None else d["real_pose"].copy(), maps={"map": d["maps"].clone()},
                             decoded_image=d["image"].clone(), collision=d.get("collision"), anno_collision=d.get("anno_collision"))
                vis.vis_pred_video(d["image"].clone(), d["scene_name"], video_type="pred")
                frames, pred = seen["frames"]
                mp4 = open(os.path.join(tmp, "videos", f"UMGen_{d['scene_name']}.mp4"), "rb").read()
                pred_mp4 = open(os.path.join(tmp, "videos_pred", f"UMGen_{d['scene_name']}.mp4"), "rb").read()
            finally:
                os.chdir(old)
        out["cases"][name] = {"shape": list(frames[0].shape), "frames": [hashlib.sha256(f.tobytes()).hexdigest() for f in frames],
                              "pred_shape": list(pred[0].shape), "pred_frames": [hashlib.sha256(np.ascontiguousarray(f).tobytes()).hexdigest() for f in pred],
                              "mp4_sha256": hashlib.sha256(mp4).hexdigest(), "mp4_bytes": len(mp4),
                              "pred_mp4_sha256": hashlib.sha256(pred_mp4).hexdigest()}
        print(name, out["cases"][name]["shape"], len(frames), "frames", len(mp4), "bytes of mp4")
    with open(os.path.join(OUT, "visualize.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("visualize.json done")


def main():
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or (["tables", "collision"] + [f"rollout:{k}" for k in ROLLOUT_CASES]
                             + [f"oar:{k}" for k in OAR_CASES] + ["vq:map", "vq:image", "postprocess", "dataset", "runner", "visualize", "samplers"])
    for w in which:
        if w == "tables":
            tables()
        elif w == "collision":
            collision()
        elif w == "postprocess":
            postprocess()
        elif w == "dataset":
            dataset()
        elif w == "runner":
            runner()
        elif w == "visualize":
            visualize()
        elif w == "samplers":
            samplers()
        elif w.startswith("vq:"):
            vq_case(w.split(":", 1)[1])
        elif w.startswith("oar:"):
            oar_case(w.split(":", 1)[1], OAR_CASES[w.split(":", 1)[1]])
        elif w.startswith("rollout:"):
            rollout(w.split(":", 1)[1], ROLLOUT_CASES[w.split(":", 1)[1]])


if __name__ == "__main__":
    main()
