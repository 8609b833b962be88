"""GPU parity of the persistent OAR decode kernel against the reference's own infer_oar_net
(tests/golden/oar_*.npz, produced by oracle/make_golden.py from the unmodified reference)."""
import dataclasses
import os

import numpy as np
import pytest
import torch

from tests._cases import OAR_CASES, apply_tweak, oar_inputs
from umgen_b200 import synth
from umgen_b200.config import ModelConfig, SampleConfig

pytestmark = pytest.mark.gpu

LOGIT_ATOL = 6e-3      # fp16 KV cache + fp32 accumulation vs the fp32 CPU reference
MARGIN_TOL = 2e-2      # greedy ids must agree wherever the reference's top1-top2 gap exceeds this


def sampled_positions():
    return list(range(7, 1031)) + list(range(1033, 1693)) + list(range(1695, 2207))


def golden_frame(g):
    full = np.zeros(2207, dtype=np.int64)
    full[[0, 4, 5, 1030, 1031, 1692, 1693, 2206]] = [0, 1, 2, 3, 4, 5, 6, 7]
    full[1:4] = g["pose"]
    full[6:1030] = g["map"]
    full[1032:1692] = g["bbox3d"]
    full[1694:2206] = g["image"]
    return full


MODES = [2, 1]      # 2 = 8-cluster kernel (decode_cluster.cu, default), 1 = L2-exchange kernel (decode.cu, the fallback)


def make_decoder(spec, mode=0):
    from umgen_b200.decoder import FrameDecoder
    cfg = dataclasses.replace(ModelConfig.tiny(1), n_oar_layer=spec["oar_layers"])
    sd = apply_tweak(synth.make_state_dict(cfg, seed=spec["weight_seed"]), spec.get("tweak"))
    dec = FrameDecoder(sd, cfg)
    if mode == 2 and dec.cluster_capacity < 8:
        pytest.skip(f"device holds only {dec.cluster_capacity} of the 8 clusters the cluster kernel needs")
    dec.mode = mode
    return dec


def compare_with_golden(res, g, vocab_check=True):
    tokens = res.tokens.cpu().numpy().astype(np.int64)
    gold = golden_frame(g)
    pos = sampled_positions()
    margins = g["top_vals"][:, 0] - g["top_vals"][:, 1]
    mism = np.nonzero(tokens != gold)[0]
    first_bad = int(mism[0]) + 1 if mism.size else 2208           # 1-indexed position
    logits = res.logits.cpu()
    n_cmp, worst = 0, 0.0
    for i, p in enumerate(pos):
        if p > first_bad:
            break
        V = 1028 if 1033 <= p <= 1692 else 8192
        top = torch.topk(logits[p - 1, :V], 8).values.numpy()
        worst = max(worst, float(np.abs(top - g["top_vals"][i]).max()))
        n_cmp += 1
    assert n_cmp > 100
    assert worst < LOGIT_ATOL, f"logit mismatch {worst}"
    if mism.size:
        i = pos.index(first_bad) if first_bad in pos else None
        assert i is not None, f"mismatch at forced/wiped position {first_bad}"
        assert margins[i] < MARGIN_TOL, f"token mismatch at {first_bad} with margin {margins[i]}"
    return first_bad, worst


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", list(OAR_CASES))
def test_decode_frame_matches_reference(name, golden_dir, mode):
    spec = OAR_CASES[name]
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    dec = make_decoder(spec, mode)
    tar_feat, pose, prev = oar_inputs(spec)
    ctrl = None if spec["control_slot"] is None else [spec["control_slot"]]
    res = dec.decode(tar_feat, pose, prev, SampleConfig.greedy(), control_slots=ctrl, want_logits=True)
    first_bad, worst = compare_with_golden(res, g)
    print(f"{name} mode={mode}: identical through position {first_bad - 1}, worst top-8 logit error {worst:.2e}, "
          f"status={res.status.cpu().tolist()}")
    assert first_bad == 2208, f"greedy ids diverge from the reference at position {first_bad}"
    st = res.status.cpu().tolist()
    if spec.get("tweak") == "padheavy":
        assert st[2] == int(g["n_tar_head_calls"]) > 0
    # slots rewritten to <pad> by the rule check (UMGen.py:1336-1377), counted by the reference run itself
    assert st[1] == int(g["n_wipes"]), f"{st[1]} wipes on the device, {int(g['n_wipes'])} in the reference"
    if spec.get("tweak") == "collide":
        # the rotated-box geometry decides here: every wipe happened with <= 30 boxes in the list (the count rule never fired)
        assert int(g["n_wipes"]) > 10 and int(g["boxes_at_wipe"].max()) <= 30
        out = res.tokens.cpu().numpy()[1032:1692].reshape(60, 11)
        born = (prev.view(60, 11)[:, 10].numpy() == 1027) & (out[:, 10] != 1027)
        assert born.sum() > 3, "some new-born boxes must survive the collision test for it to be decisive"


def test_collision_geometry_on_the_device_matches_the_reference(golden_dir):
    """All 912 BoxOverlap.check_collision answers recorded from the reference (tests/golden/collision.npz) reproduced by the device functions
    the decode kernels' rule path runs (umgen_check_collision -> box_corners, last_box_collides, pair_collides in csrc/decode_shared.cuh)."""
    from tests._cases import collision_cases
    from umgen_b200 import capi
    cases = collision_cases()
    ans = np.load(os.path.join(golden_dir, "collision.npz"))["answers"]
    assert len(cases) == len(ans) == 912
    offs = np.zeros(len(cases) + 1, dtype=np.int32)
    offs[1:] = np.cumsum([len(c) for c in cases])
    boxes = torch.from_numpy(np.concatenate([np.stack(c) for c in cases]).astype(np.float64)).cuda()
    offs_d = torch.from_numpy(offs).cuda()
    out = torch.full((len(cases),), -7, dtype=torch.int32, device="cuda")
    capi.check(capi.lib().umgen_check_collision(boxes.data_ptr(), offs_d.data_ptr(), len(cases), out.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream), "umgen_check_collision")
    got = out.cpu().numpy()
    bad = np.nonzero(got != ans.astype(np.int32))[0]
    assert bad.size == 0, f"{bad.size} of 912 collision answers differ from the reference, first case {int(bad[0])}"
    assert 100 < int(ans.sum()) < 800


@pytest.mark.parametrize("mode", MODES)
def test_given_prefix_is_forced_without_sampling(golden_dir, mode):
    """prefix_len (init_tokens of UMGen.py:1184-1201): the map block is given; the kernel must feed it forward unchanged, run no head for it
    and continue exactly like the free-running frame whose map block it is."""
    name = "oar_L2"
    spec = OAR_CASES[name]
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    dec = make_decoder(spec, mode)
    tar_feat, pose, prev = oar_inputs(spec)
    teacher = torch.zeros(2207, dtype=torch.int32)
    teacher[1:4] = pose.to(torch.int32)
    teacher[6:1030] = torch.from_numpy(g["map"].astype(np.int32))
    res = dec.decode(tar_feat, pose, prev, SampleConfig.greedy(), teacher=teacher, prefix_len=1031, want_logits=True)
    toks = res.tokens.cpu().numpy().astype(np.int64)
    assert np.array_equal(toks, golden_frame(g))
    assert float(res.logits[6:1030].abs().max()) == 0.0, "no head may run for the given positions"
    # a different given map changes what follows (the prefix really is the conditioning)
    teacher2 = teacher.clone()
    teacher2[6:1030] = (teacher[6:1030] + 1) % 8192
    l1 = res.logits[1032:1692, :1028].clone()
    res2 = dec.decode(tar_feat, pose, prev, SampleConfig.greedy(), teacher=teacher2, prefix_len=1031, want_logits=True)
    t2 = res2.tokens.cpu().numpy()
    assert np.array_equal(t2[6:1030], teacher2[6:1030].numpy())
    assert float((res2.logits[1032:1692, :1028] - l1).abs().max()) > 1e-3, "the logits after the prefix must depend on it"


@pytest.mark.parametrize("mode", MODES)
def test_decode_is_deterministic_run_to_run(mode):
    spec = OAR_CASES["oar_L2"]
    dec = make_decoder(spec, mode)
    tar_feat, pose, prev = oar_inputs(spec)
    outs = []
    for _ in range(2):
        r = dec.decode(tar_feat, pose, prev, SampleConfig.greedy(), want_logits=True, n_steps=1500)
        outs.append((r.tokens.clone(), r.logits.clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("mode", MODES)
def test_topk_sampling_stays_inside_topk_and_is_seeded(mode):
    spec = OAR_CASES["oar_L2"]
    dec = make_decoder(spec, mode)
    tar_feat, pose, prev = oar_inputs(spec)
    sc = SampleConfig(top_k=5, top_k_map=5, top_k_image=16, seed=7)
    r1 = dec.decode(tar_feat, pose, prev, sc, want_logits=True, n_steps=1100)
    t1, l1 = r1.tokens.cpu().clone(), r1.logits.cpu().clone()
    n_not_argmax = 0
    for p in range(7, 1031):
        top = torch.topk(l1[p - 1], 5).indices.tolist()
        assert int(t1[p - 1]) in top, p
        n_not_argmax += int(t1[p - 1]) != top[0]
    assert n_not_argmax > 100          # genuinely sampling, not arg-max
    r2 = dec.decode(tar_feat, pose, prev, sc, n_steps=1100)
    assert torch.equal(r2.tokens.cpu()[:1100], t1[:1100])            # same seed -> same stream
    sc2 = SampleConfig(top_k=5, top_k_map=5, top_k_image=16, seed=8)
    r3 = dec.decode(tar_feat, pose, prev, sc2, n_steps=1100)
    assert not torch.equal(r3.tokens.cpu()[:1100], t1[:1100])


def _nucleus(logits_row, p):
    """ids kept by the reference's sample_top_p mask (UMGen.py:946-953) for one logits row."""
    probs = torch.softmax(logits_row.double(), -1)
    ps, pi = torch.sort(probs, descending=True)
    keep = (torch.cumsum(ps, -1) - ps) <= p
    return set(pi[keep].tolist())


@pytest.mark.parametrize("mode", MODES)
def test_topp_tiny_p_is_greedy(golden_dir, mode):
    name = "oar_L2"
    spec = OAR_CASES[name]
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    dec = make_decoder(spec, mode)
    tar_feat, pose, prev = oar_inputs(spec)
    sc = SampleConfig(method="topp", p=1e-7, p_map=1e-7, top_k_image=1e-7, seed=3)
    res = dec.decode(tar_feat, pose, prev, sc, n_steps=1400)
    toks = res.tokens.cpu().numpy().astype(np.int64)
    assert np.array_equal(toks[6:1030], g["map"])             # only the arg-max survives a vanishing nucleus
    assert np.array_equal(res.picks.cpu().numpy()[1032:1400], g["input_stream"][1024:1024 + 368])


@pytest.mark.parametrize("mode", MODES)
def test_topp_samples_inside_the_nucleus_and_is_seeded(mode):
    spec = OAR_CASES["oar_L2"]
    dec = make_decoder(spec, mode)
    tar_feat, pose, prev = oar_inputs(spec)
    # peaked logits so the nucleus is small: scale the map head
    dec.w["head_map_h"].mul_(12.0)
    sc = SampleConfig(method="topp", p=0.4, p_map=0.4, seed=11)
    r1 = dec.decode(tar_feat, pose, prev, sc, want_logits=True, n_steps=700)
    t1, l1 = r1.tokens.cpu().clone(), r1.logits.cpu().clone()
    sizes, n_not_top = [], 0
    for p in range(7, 700):
        nuc = _nucleus(l1[p - 1], 0.4)
        assert int(t1[p - 1]) in nuc, (p, int(t1[p - 1]), len(nuc))
        sizes.append(len(nuc))
        n_not_top += int(t1[p - 1]) != int(l1[p - 1].argmax())
    assert n_not_top > 5 and max(sizes) > 1
    r2 = dec.decode(tar_feat, pose, prev, sc, n_steps=700)
    assert torch.equal(r2.tokens.cpu()[:700], t1[:700])
    r3 = dec.decode(tar_feat, pose, prev, SampleConfig(method="topp", p=0.4, p_map=0.4, seed=12), n_steps=700)
    assert not torch.equal(r3.tokens.cpu()[:700], t1[:700])


def _frag_blocks(m):
    """[R, K] (R, K multiples of 16) -> [R/16, K/16, 256]: every 16x16 tile in mma.m16n8k16 A-fragment order (include/umgen.h)."""
    e = torch.arange(256)
    lane, reg, hp = e // 8, (e % 8) // 2, e % 2
    row = lane // 4 + 8 * (reg & 1)
    col = 2 * (lane % 4) + hp + 8 * (reg // 2)
    R, K = m.shape
    t = m.view(R // 16, 16, K // 16, 16).permute(0, 2, 1, 3)          # [mt, kt, row, col]
    return t[:, :, row.to(m.device), col.to(m.device)]


def test_cluster_weight_packing_matches_the_documented_layout():
    """umgen_pack_oar_cluster against the layout stated in include/umgen.h, restated with torch indexing."""
    from umgen_b200 import capi
    lib = capi.lib()
    L = 2
    LH = 2304 * 768 + 768 * 768 + 3072 * 768 + 768 * 3072
    src = (torch.arange(L * LH, device="cuda", dtype=torch.int64) * 7919 % 65521 - 32760).to(torch.int16).view(L, LH)
    dst = torch.empty_like(src)
    capi.check(lib.umgen_pack_oar_cluster(src.data_ptr(), dst.data_ptr(), L, torch.cuda.current_stream().cuda_stream), "pack")
    torch.cuda.synchronize()
    for l in range(L):
        o = 0
        cattn = src[l, o:o + 2304 * 768].view(2304, 768); o += 2304 * 768
        cproj = src[l, o:o + 768 * 768].view(768, 768); o += 768 * 768
        cfc = src[l, o:o + 3072 * 768].view(3072, 768); o += 3072 * 768
        cproj2 = src[l, o:].view(768, 3072)
        for g in (0, 9, 37, 63):
            cluster, rank = divmod(g, 8)
            got = dst[l].view(64, -1)[g]
            # c_attn: 36 local rows, padded to 48 for the tiling; [warp][k-step][tile0 | tile1 | 4-row tile]
            rows = [w * 768 + (2 * cluster + hh) * 48 + 6 * rank + e for hh in range(2) for w in range(3) for e in range(6)]
            m = torch.zeros(48, 768, dtype=torch.int16, device="cuda")
            m[:36] = cattn[rows]
            fb = _frag_blocks(m).view(3, 12, 4, 256)                    # [tile, warp, k-step, 256]
            rem = fb[2].view(12, 4, 32, 4, 2)[:, :, :16][:, :, :, [0, 2]].reshape(12, 4, 64)     # lanes 0..15 x {reg 0, reg 2}
            want = torch.cat([fb[0], fb[1], rem], dim=2).reshape(-1)
            n = want.numel()
            assert n == 36 * 768 and torch.equal(got[:n], want)
            p = n
            # c_proj: [tile 6][k-step 6]
            want = _frag_blocks(cproj[96 * rank:96 * rank + 96, 96 * cluster:96 * cluster + 96].contiguous()).reshape(-1)
            assert torch.equal(got[p:p + want.numel()], want); p += want.numel()
            # c_fc: [warp 12][k-step 4][tile 3]
            want = _frag_blocks(cfc[48 * g:48 * g + 48].contiguous()).view(3, 12, 4, 256).permute(1, 2, 0, 3).reshape(-1)
            assert torch.equal(got[p:p + want.numel()], want); p += want.numel()
            # mlp c_proj: [tile 48][k-step 3]
            want = _frag_blocks(cproj2[:, 48 * g:48 * g + 48].contiguous()).reshape(-1)
            assert torch.equal(got[p:p + want.numel()], want); p += want.numel()
            assert p == got.numel()


# ---- several scenes per launch (umgen_decode_frames) ---------------------------------------------------------------
def _second_scene(spec):
    """Inputs of another scene for the same weights: different conditioning feature, pose and previous boxes."""
    other = dict(spec, feat_seed=spec["feat_seed"] + 100, scene_seed=spec["scene_seed"] + 100)
    return oar_inputs(other)


def _single_and_batched(dec, frames, sample, n_steps=2206, prefix_len=0):
    """Every scene through decode() alone, then all of them through ONE launch; returns (single results, batched results) as host tensors."""
    from umgen_b200.decoder import FrameDecoder
    if dec.cluster_capacity < 8:
        pytest.skip(f"device holds only {dec.cluster_capacity} of the 8 clusters the cluster kernel needs")
    single = []
    for f in frames:
        sc = dataclasses.replace(sample, seed=f["seed"]) if f.get("seed") is not None else sample
        r = dec.decode(f["tar_feat"], f["pose_tok"], f["prev_bbox"], sc, frame_index=f.get("frame_index", 0), control_slots=f.get("control_slots"),
                       teacher=f.get("teacher"), want_logits=True, n_steps=n_steps, prefix_len=prefix_len)
        single.append((r.tokens.cpu().clone(), r.picks.cpu().clone(), r.logits.cpu().clone(), r.status.cpu().tolist()))
    decs = [dec] + [dec.for_scene() for _ in frames[1:]]
    res = FrameDecoder.decode_batch(decs, frames, sample, want_logits=True, n_steps=n_steps, prefix_len=prefix_len)
    batched = [(r.tokens.cpu().clone(), r.picks.cpu().clone(), r.logits.cpu().clone(), r.status.cpu().tolist()) for r in res]
    return single, batched


def _assert_same(single, batched, n):
    for s, (a, b) in enumerate(zip(single, batched)):
        assert torch.equal(a[0][:n], b[0][:n]), f"scene {s}: ids differ from the one-scene launch, first at {int((a[0][:n] != b[0][:n]).nonzero()[0])}"
        assert torch.equal(a[1][:n], b[1][:n]), f"scene {s}: picks differ"
        assert torch.equal(a[2][:n], b[2][:n]), f"scene {s}: logits are not bit-identical (max diff {float((a[2][:n] - b[2][:n]).abs().max())})"
        assert a[3][:4] == b[3][:4], f"scene {s}: status {b[3][:4]} vs {a[3][:4]}"


@pytest.mark.parametrize("name", ["oar_L2", "oar_L1_collide", "oar_L1_control", "oar_L1_padheavy"])
def test_two_scenes_per_launch_are_bit_identical_to_one_scene_launches(name, golden_dir):
    """Scene 0 = the golden case (so the batched launch is also pinned to the reference), scene 1 = other inputs; ids, picks, logits and the
    rule-path counters of both scenes must equal what one-scene launches produce."""
    spec = OAR_CASES[name]
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    dec = make_decoder(spec, 2)
    f0 = dict(zip(("tar_feat", "pose_tok", "prev_bbox"), oar_inputs(spec)))
    f1 = dict(zip(("tar_feat", "pose_tok", "prev_bbox"), _second_scene(spec)))
    if spec["control_slot"] is not None:
        f0["control_slots"] = [spec["control_slot"]]          # scene 1 stays uncontrolled
    single, batched = _single_and_batched(dec, [f0, f1], SampleConfig.greedy())
    _assert_same(single, batched, 2207)
    assert np.array_equal(batched[0][0].numpy().astype(np.int64), golden_frame(g)), "scene 0 of the batched launch must reproduce the reference"
    assert batched[0][3][1] == int(g["n_wipes"])
    assert not torch.equal(batched[0][0], batched[1][0])
    # and with the scenes swapped (column pairs 0/1 and 2/3 of the MMA B operand trade places)
    single_r, batched_r = _single_and_batched(dec, [f1, f0], SampleConfig.greedy(), n_steps=1200)
    _assert_same(single_r, batched_r, 1200)
    assert torch.equal(batched_r[1][0][:1200], batched[0][0][:1200])


@pytest.mark.parametrize("method", ["topk", "topp"])
def test_two_scenes_per_launch_sample_like_one_scene_launches(method):
    spec = OAR_CASES["oar_L2"]
    dec = make_decoder(spec, 2)
    f0 = dict(zip(("tar_feat", "pose_tok", "prev_bbox"), oar_inputs(spec)), seed=7, frame_index=3)
    f1 = dict(zip(("tar_feat", "pose_tok", "prev_bbox"), _second_scene(spec)), seed=8, frame_index=5)
    if method == "topk":
        sc = SampleConfig(top_k=5, top_k_map=5, top_k_image=16)
    else:
        dec.w["head_map_h"].mul_(12.0)
        sc = SampleConfig(method="topp", p=0.4, p_map=0.4)
    n = 1150
    single, batched = _single_and_batched(dec, [f0, f1], sc, n_steps=n)
    _assert_same(single, batched, n)
    arg = single[0][2][6:1030].argmax(-1).to(torch.int32)
    assert int((single[0][0][6:1030] != arg).sum()) > 5, "the test must really sample"


def test_two_scenes_per_launch_with_a_given_prefix(golden_dir):
    spec = OAR_CASES["oar_L2"]
    g = np.load(os.path.join(golden_dir, "oar_L2.npz"))
    dec = make_decoder(spec, 2)
    frames = []
    for k, inp in enumerate((oar_inputs(spec), _second_scene(spec))):
        teacher = torch.zeros(2207, dtype=torch.int32)
        teacher[1:4] = inp[1].to(torch.int32)
        teacher[6:1030] = (torch.from_numpy(g["map"].astype(np.int32)) + 17 * k) % 8192
        frames.append(dict(zip(("tar_feat", "pose_tok", "prev_bbox"), inp), teacher=teacher))
    single, batched = _single_and_batched(dec, frames, SampleConfig.greedy(), n_steps=1500, prefix_len=1031)
    _assert_same(single, batched, 1500)
    n_done = 1032 + (1500 - 1032) // 11 * 11          # the rule check rewrites a slot when its 11th token is sampled: compare whole slots only
    assert np.array_equal(batched[0][0].numpy().astype(np.int64)[:n_done], golden_frame(g)[:n_done])


def test_decode_frames_rejects_mismatched_scenes():
    from umgen_b200 import capi
    from umgen_b200.decoder import FrameDecoder
    spec = OAR_CASES["oar_L2"]
    dec = make_decoder(spec, 2)
    f0 = dict(zip(("tar_feat", "pose_tok", "prev_bbox"), oar_inputs(spec)))
    with pytest.raises(capi.UmgenError, match="share a state"):
        FrameDecoder.decode_batch([dec, dec], [f0, f0], SampleConfig.greedy(), n_steps=10)
    other = make_decoder(spec, 2)
    with pytest.raises(capi.UmgenError, match="share weights"):
        FrameDecoder.decode_batch([dec, other], [f0, f0], SampleConfig.greedy(), n_steps=10)
    with pytest.raises(capi.UmgenError, match="at most"):
        FrameDecoder.decode_batch([dec] + [dec.for_scene() for _ in range(4)], [f0] * 5, SampleConfig.greedy(), n_steps=10)


def test_three_scenes_per_launch_are_bit_identical_to_one_scene_launches(golden_dir):
    """The largest batch of this build (umgen_decode_max_scenes() = 3: column pairs 0/1, 2/3, 4/5 of every MMA's B operand), with the rule path deciding
    wipes in every scene (collision-steered head), a controlled slot in the middle scene only, and greedy / top-k sampling."""
    from umgen_b200 import capi
    if int(capi.lib().umgen_decode_max_scenes()) < 3:
        pytest.skip("this build takes fewer than 3 scenes per launch")
    name = "oar_L1_collide"
    spec = OAR_CASES[name]
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    dec = make_decoder(spec, 2)
    frames = [dict(zip(("tar_feat", "pose_tok", "prev_bbox"), oar_inputs(dict(spec, feat_seed=spec["feat_seed"] + 100 * k, scene_seed=spec["scene_seed"] + 100 * k))),
                   seed=5 + k) for k in range(3)]
    frames[1]["control_slots"] = [2]
    single, batched = _single_and_batched(dec, frames, SampleConfig.greedy())
    _assert_same(single, batched, 2207)
    assert np.array_equal(batched[0][0].numpy().astype(np.int64), golden_frame(g)), "scene 0 of the batched launch must reproduce the reference"
    assert batched[0][3][1] == int(g["n_wipes"]) > 10
    assert len({tuple(b[0].tolist()) for b in batched}) == 3
    sc = SampleConfig(top_k=5, top_k_map=5, top_k_image=16)
    single, batched = _single_and_batched(dec, frames, sc, n_steps=1300)
    _assert_same(single, batched, 1300)
