"""GPU unit parity of the TAR kernels against plain PyTorch fp32 references of the same op
(the op-level semantics are those of the reference's module.py / UMGen.py lines cited in tar.cu)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def rnd(*shape, scale=1.0, dtype=torch.float32, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device=dev()) * scale).to(dtype)


@pytest.mark.parametrize("M,N,K", [(300, 256, 64), (128, 768, 768), (2207, 2304, 768), (1031 * 3, 768, 3072), (3, 1024, 768),
                                   (44140, 3072, 768)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm_matches_torch(M, N, K, epi):
    from umgen_b200 import ops
    if M > 10000 and epi not in (1, 2):
        pytest.skip("large shape checked for two epilogues only")
    a = rnd(M, K, scale=1.0, dtype=torch.float16, seed=1)
    w = rnd(N, K, scale=1.0 / math.sqrt(K), dtype=torch.float16, seed=2)
    bias = rnd(N, scale=0.1, seed=3) if epi != 1 else None
    ref = a.float() @ w.float().t()
    if bias is not None:
        ref = ref + bias
    if epi in (0, 1):
        out = torch.zeros(M, N, dtype=torch.float16, device=dev())
        if epi == 1:
            ref = F.gelu(ref)
        ops.gemm(a, w, bias, out, epi)
        torch.testing.assert_close(out.float(), ref, atol=4e-3, rtol=2e-3)
    else:
        base = rnd(M, N, seed=4)
        out = base.clone()
        ops.gemm(a, w, bias, out, epi)
        want = ref + base if epi == 2 else ref
        torch.testing.assert_close(out, want, atol=2e-4, rtol=1e-4)


def test_layernorm_matches_torch():
    from umgen_b200 import ops
    x = rnd(2207 * 2 + 5, 768, scale=3.0, seed=5) + 0.7
    w = 1 + 0.1 * rnd(768, seed=6)
    ref = F.layer_norm(x, (768,), w, None, 1e-5)
    out32 = torch.empty_like(x)
    ops.layernorm(x, w, out32)
    torch.testing.assert_close(out32, ref, atol=2e-5, rtol=1e-5)
    out16 = torch.empty(x.shape, dtype=torch.float16, device=dev())
    ops.layernorm(x, w, out16)
    torch.testing.assert_close(out16.float(), ref, atol=4e-3, rtol=1e-3)


def _attn_ref(q, k, v, causal):
    # q [B,Tq,16,48] ... fp32 math, bottom-right causal alignment (Tq == Tk here)
    qh, kh, vh = (t.float().transpose(1, 2) for t in (q, k, v))
    att = qh @ kh.transpose(-1, -2) / math.sqrt(48)
    if causal:
        Tq, Tk = att.shape[-2:]
        att = att.masked_fill(~torch.ones(Tq, Tk, dtype=torch.bool, device=att.device).tril(Tk - Tq), float("-inf"))
    return (torch.softmax(att, -1) @ vh).transpose(1, 2)


@pytest.mark.parametrize("kernel", ["tcgen05", "mma"])
@pytest.mark.parametrize("T,S", [(1, 64), (2, 2207), (3, 1031), (2, 1693), (1, 77), (1, 128), (20, 300)])
def test_spatial_attention_matches_torch(T, S, kernel):
    from umgen_b200 import ops
    qkv = rnd(T * S, 2304, scale=1.0, dtype=torch.float16, seed=7)
    y = torch.zeros(T * S, 768, dtype=torch.float16, device=dev())
    (ops.spatial_attention if kernel == "tcgen05" else ops.spatial_attention_mma)(qkv, y, T, S)
    q, k, v = (qkv[:, i * 768:(i + 1) * 768].reshape(T, S, 16, 48) for i in range(3))
    ref = _attn_ref(q, k, v, False).reshape(T * S, 768)
    torch.testing.assert_close(y.float(), ref, atol=3e-3, rtol=2e-3)


def test_spatial_attention_with_peaked_scores_rescales_correctly():
    """Scores that keep growing along the key axis force the lazy rescale of the tcgen05 kernel's TMEM accumulator (a row's reference maximum
    moves only when it is outgrown by 2^8) through many tiles; a plain softmax must still come out."""
    from umgen_b200 import ops
    T, S = 1, 1500
    qkv = rnd(T * S, 2304, scale=0.3, dtype=torch.float16, seed=17)
    ramp = torch.linspace(0, 6, S, device=dev())[:, None]
    qkv[:, 768:1536] = (qkv[:, 768:1536].float() * (1 + ramp)).half()          # keys grow -> later tiles hold the maxima
    qkv[:, :768] = (qkv[:, :768].float() * 4).half()
    y = torch.zeros(T * S, 768, dtype=torch.float16, device=dev())
    ops.spatial_attention(qkv, y, T, S)
    q, k, v = (qkv[:, i * 768:(i + 1) * 768].reshape(T, S, 16, 48) for i in range(3))
    ref = _attn_ref(q, k, v, False).reshape(T * S, 768)
    torch.testing.assert_close(y.float(), ref, atol=4e-3, rtol=4e-3)


@pytest.mark.parametrize("T,S", [(20, 300), (3, 1031), (1, 50), (13, 97)])
def test_temporal_attention_matches_torch(T, S):
    from umgen_b200 import ops
    qkv = rnd(T * S, 2304, scale=1.0, dtype=torch.float16, seed=8)
    y = torch.zeros(T * S, 768, dtype=torch.float16, device=dev())
    ops.small_attention(qkv, y, S, T, 1, S, True)       # group = position s, tokens = frames
    q, k, v = (qkv[:, i * 768:(i + 1) * 768].reshape(T, S, 16, 48).transpose(0, 1) for i in range(3))   # [S,T,16,48]
    ref = _attn_ref(q, k, v, True).transpose(0, 1).reshape(T * S, 768)
    torch.testing.assert_close(y.float(), ref, atol=3e-3, rtol=2e-3)


def test_ego_self_and_cross_attention_match_torch():
    from umgen_b200 import ops
    qkv = rnd(3, 2304, dtype=torch.float16, seed=9)
    y = torch.zeros(3, 768, dtype=torch.float16, device=dev())
    ops.small_attention(qkv, y, 1, 3, 0, 1, False)
    q, k, v = (qkv[:, i * 768:(i + 1) * 768].reshape(1, 3, 16, 48) for i in range(3))
    torch.testing.assert_close(y.float(), _attn_ref(q, k, v, False).reshape(3, 768), atol=3e-3, rtol=2e-3)
    q = rnd(3, 768, dtype=torch.float16, seed=10)
    k = rnd(2207, 768, dtype=torch.float16, seed=11)
    v = rnd(2207, 768, dtype=torch.float16, seed=12)
    y = torch.zeros(3, 768, dtype=torch.float16, device=dev())
    ops.cross_attention(q, k, v, y)
    ref = _attn_ref(q.reshape(1, 3, 16, 48), k.reshape(1, 2207, 16, 48), v.reshape(1, 2207, 16, 48), False).reshape(3, 768)
    torch.testing.assert_close(y.float(), ref, atol=3e-3, rtol=2e-3)


def test_map_warp_matches_grid_sample():
    from umgen_b200 import ops
    from umgen_b200.weights import pose_value_lut
    T = 4
    feat = rnd(T, 1024, 768, seed=13)
    lut = torch.from_numpy(pose_value_lut()).to(dev())
    pose_tok = torch.tensor([[512, 512, 512], [700, 300, 600], [100, 900, 480], [1023, 0, 10]], dtype=torch.int32, device=dev())
    out = torch.empty_like(feat)
    ops.map_warp(feat, pose_tok, lut, out)
    pv = torch.stack([lut[pose_tok[:, c].long(), c] for c in range(3)], 1)
    th, dx, dy = pv[:, 2], 2 * (pv[:, 0] / 4) / 32, 2 * (pv[:, 1] / 4) / 32
    mat = torch.zeros(T, 2, 3, device=dev())
    mat[:, 0, 0], mat[:, 0, 1], mat[:, 0, 2] = torch.cos(-th), -torch.sin(-th), -dy
    mat[:, 1, 0], mat[:, 1, 1], mat[:, 1, 2] = torch.sin(-th), torch.cos(-th), -dx
    img = feat.transpose(1, 2).reshape(T, 768, 32, 32)
    grid = F.affine_grid(mat, (T, 768, 32, 32), align_corners=False)
    ref = F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=False).reshape(T, 768, 1024).transpose(1, 2)
    torch.testing.assert_close(out, ref, atol=2e-4, rtol=1e-4)


def test_sample_rows_greedy_and_topk():
    from umgen_b200 import ops
    logits = rnd(3, 1024, seed=14)
    out = torch.zeros(3, dtype=torch.int32, device=dev())
    ops.sample_rows(logits, 1, 1.0, 0, 0, out)
    assert torch.equal(out.long(), logits.argmax(-1))
    ops.sample_rows(logits, 5, 1.0, 3, 7, out)
    top = torch.topk(logits, 5).indices
    assert all(int(out[i]) in top[i].tolist() for i in range(3))


def _nucleus(row, p):
    """ids kept by the reference's sample_top_p mask (UMGen.py:946-953) for one logits row."""
    probs = torch.softmax(row.double(), -1)
    ps, pi = torch.sort(probs, descending=True)
    return set(pi[(torch.cumsum(ps, -1) - ps) <= p].tolist())


def test_sample_rows_top_p_is_the_reference_nucleus():
    """The ego head under sample_method "topp" (UMGen.py:1001-1004 -> sample_top_p): every pick lies inside the reference's nucleus, a
    vanishing p is the arg-max, draws depend on (seed, frame) only, and different seeds give different draws."""
    from umgen_b200 import ops
    logits = rnd(3, 1024, scale=4.0, seed=15)
    out = torch.zeros(3, dtype=torch.int32, device=dev())
    ops.sample_rows(logits, 1, 1.0, 0, 0, out, top_p=1e-7)
    assert torch.equal(out.long(), logits.argmax(-1))
    seen = [set(), set(), set()]
    for seed in range(40):
        ops.sample_rows(logits, 1, 1.0, seed, 2, out, top_p=0.4)
        for r in range(3):
            assert int(out[r]) in _nucleus(logits[r].cpu(), 0.4), (seed, r)
            seen[r].add(int(out[r]))
    assert max(len(s) for s in seen) > 1, "top-p draws never left the arg-max"
    a, b = out.clone(), torch.zeros_like(out)
    ops.sample_rows(logits, 1, 1.0, 39, 2, b, top_p=0.4)
    assert torch.equal(a, b)
    # temperature: a hot distribution widens the nucleus
    ops.sample_rows(logits, 1, 50.0, 1, 0, out, top_p=0.4)
    assert all(int(out[r]) in _nucleus(logits[r].cpu() / 50.0, 0.4) for r in range(3))


def test_ego_head_follows_the_sample_method():
    """TarEncoders.ego_action dispatches on SampleConfig.method like the reference's token_sampler (UMGen.py:118-126)."""
    from umgen_b200 import synth
    from umgen_b200.config import ModelConfig, SampleConfig
    from umgen_b200.tar import TarEncoders
    cfg = ModelConfig.tiny(1, cond_frame=2)
    tar = TarEncoders(synth.make_state_dict(cfg, seed=2), cfg)
    scene = synth.make_scene(seed=3, n_frames=2)
    tok = TarEncoders.to_device_tokens({m: scene[m][0] for m in scene}, tar.dev)
    picks = set()
    for seed in range(12):
        t = tar.ego_action(tok, SampleConfig(method="topp", p=0.9, seed=seed), 0).cpu()
        for r in range(3):
            assert int(t[r]) in _nucleus(tar.ego_logits[r].cpu(), 0.9)
        picks.add(tuple(t.tolist()))
    assert len(picks) > 1
    greedy = tar.ego_action(tok, SampleConfig(method="topp", p=1e-7), 0).cpu()
    assert torch.equal(greedy.long(), tar.ego_logits.argmax(-1).cpu())
