"""CPU oracle for UMGen's next-scene decode path.  TEST INFRASTRUCTURE ONLY.

This file is a functional fp32 restatement (torch-on-CPU + numpy) of the
reference algorithm for the hot path of SURVEY.md section 8 (rows a1-a10).  It
is imported only by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- never by the
product package ``umgen_b200`` (which fails loudly without its CUDA library).

Parity pinning: the reference ships no tests or golden vectors, so this oracle
is pinned against the *reference implementation itself* imported from
``/root/reference`` in the build container (``oracle/ref_import.py``); the
resulting vectors are committed under ``tests/golden/`` together with the
script that made them (``oracle/make_golden.py``), and
``tests/test_oracle_golden.py`` replays them.

Every function cites the reference lines it follows (paths relative to
``/root/reference/projects``).  Weights are consumed as a flat
``{state_dict key: tensor}`` mapping using the reference's own key names
(SURVEY.md section 8b "Weight-key ABI").
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# Fixed sequence geometry (configs/UMGen_config_evaluation.py:284-290,
# tools/infer_fun.py:99-118)
# ----------------------------------------------------------------------------
MODS = ("pose", "map", "bbox3d", "image")
CONTENT_LEN = {"pose": 3, "map": 1024, "bbox3d": 660, "image": 512}
TOKEN_LEN = {m: CONTENT_LEN[m] + 2 for m in MODS}          # with bos/eos
BOS_EOS = {"pose": (0, 1), "map": (2, 3), "bbox3d": (4, 5), "image": (6, 7)}
SEQ_LEN = sum(TOKEN_LEN.values())                          # 2207
N_SLOTS, N_ATTR = 60, 11                                   # bbox3d: 60 x (10 + cat)
PAD_TOKEN = 1027                                           # BBox3DTokenizer <pad>
BBOX_FIRST_POS = 1032                                      # UMGen.py:1082,1285
TASK_ID = 6                                                # task_name_id (config:158)
TASKS = {
    "pose_map": ("pose", "map"),
    "pose_map_bbox3d": ("pose", "map", "bbox3d"),
    "pose_map_bbox3d_image": MODS,
}
EGO_LWH = (5.176, 2.297, 1.777)                            # UMGen.py:9-12

# normalisation ranges of the 10 box attributes (config:126-137)
BOX_RANGES = np.array(
    [(-64, 64), (-64, 64), (-5, 5), (0, 15), (0, 4), (0, 5), (-3.14, 3.14),
     (-20, 20), (-15, 15), (-0.3, 0.3)], dtype=np.float64)


def mod_offsets(mods: Sequence[str]) -> Dict[str, int]:
    off, out = 0, {}
    for m in mods:
        out[m] = off
        off += TOKEN_LEN[m]
    return out


def forced_positions(mods: Sequence[str] = MODS) -> Dict[int, int]:
    """1-indexed sequence position -> forced bos/eos id (UMGen.py:976-984)."""
    d, cur = {}, 0
    for m in mods:
        cur += 1
        d[cur] = BOS_EOS[m][0]
        cur += TOKEN_LEN[m] - 1
        d[cur] = BOS_EOS[m][1]
    return d


def pos_mod(pos: int, mods: Sequence[str] = MODS) -> str:
    """Modality that owns 1-indexed position ``pos`` (UMGen.py:986-992)."""
    cur = 0
    for m in mods:
        if cur + 1 <= pos <= cur + TOKEN_LEN[m]:
            return m
        cur += TOKEN_LEN[m]
    raise ValueError(pos)


# ----------------------------------------------------------------------------
# Fixed tables
# ----------------------------------------------------------------------------
def sinusoid_table(n_position: int, emb_dim: int, start_index: int = 0) -> torch.Tensor:
    """module.py:746-768 -- row 0 is zero, row p holds sin/cos((p+start)/10000^(2(j//2)/dim)),
    stored as bfloat16."""
    pos = np.arange(n_position, dtype=np.float64)[:, None] + start_index
    j = np.arange(emb_dim)
    ang = pos / np.power(10000.0, 2.0 * (j // 2) / emb_dim)[None, :]
    tab = np.zeros((n_position, emb_dim), dtype=np.float64)
    tab[1:, 0::2] = np.sin(ang[1:, 0::2])
    tab[1:, 1::2] = np.cos(ang[1:, 1::2])
    return torch.from_numpy(tab).to(torch.bfloat16)


def grid_center_tokens() -> Tuple[np.ndarray, np.ndarray]:
    """Tokens of the 32x32 map-cell centres (UMGen.py:140-150, 357-383)."""
    idx = np.arange(32, dtype=np.float32)
    cx = -((idx + 0.5) * 4.0 - 64.0)                        # [32]
    gx, gy = np.meshgrid(cx, cx, indexing="ij")             # [i, j] -> x by i, y by j
    norm_x, norm_y = (gx + 64.0) / 128.0, (gy + 64.0) / 128.0
    bins = np.linspace(0.0, 1.0, 1024)
    return (np.digitize(norm_x, bins).reshape(1024), np.digitize(norm_y, bins).reshape(1024))


def grid_center_embedding(bbox3d_spatial_posi: torch.Tensor) -> torch.Tensor:
    """UMGen.py:151-153: bf16 sum of the x- and y-token sinusoid rows, [1024, C]."""
    tx, ty = grid_center_tokens()
    return bbox3d_spatial_posi[torch.from_numpy(tx)] + bbox3d_spatial_posi[torch.from_numpy(ty)]


def pose_value_lut() -> np.ndarray:
    """[1024, 3] float32: token -> (dx, dy, dheading).

    DigitalBinsTokenizer.decode (tokenizer.py:332-354): midpoint of the bins either side
    of the token on linspace(-1, 1, 1024) (token 0 -> bins[0]); then
    Normalize_Standard.unnormalize_ego (normalize.py:65-76): value / float32(1/std) with
    std = (10, 4, 1) (config:226-234); cast to float32 at UMGen.py:1020-1022."""
    bins = np.linspace(-1.0, 1.0, 1024)
    tok = np.arange(1024)
    right = np.clip(tok, 0, 1023)
    left = np.clip(tok - 1, 0, 1023)
    mid = (bins[left] + bins[right]) / 2
    inv_std = 1.0 / np.array([10.0, 4.0, 1.0], dtype=np.float32)
    return (mid[:, None] / inv_std[None, :] + np.zeros(3, np.float32)).astype(np.float32)


def decode_pose(pose_tokens: torch.Tensor) -> torch.Tensor:
    """UMGen.py:1008-1024: int tokens [..., 3] -> float32 values."""
    lut = torch.from_numpy(pose_value_lut())
    t = pose_tokens.long().clamp(0, 1023)
    return torch.stack([lut[t[..., c], c] for c in range(3)], dim=-1)


def box_value_lut() -> np.ndarray:
    """[1028+, 10] float64: attribute token -> metric value.

    BBox3DTokenizer.decode_single_objects (tokenizer.py:679-687, keep_order=True): bin midpoints
    on linspace(0, 1, 1024), out-of-range tokens (categories, <pad>) clip to the last bin;
    Normalize.unnormalize_bbox3d (normalize.py:189-229): v * (max - min) + min."""
    bins = np.linspace(0.0, 1.0, 1024)
    tok = np.arange(1028)
    right = np.clip(tok, 0, 1023)
    left = np.clip(tok - 1, 0, 1023)
    mid = (bins[left] + bins[right]) / 2                     # [1028]
    lo, hi = BOX_RANGES[:, 0], BOX_RANGES[:, 1]
    return mid[:, None] * (hi - lo)[None, :] + lo[None, :]


# ----------------------------------------------------------------------------
# Rotated-box collision (misc.py:143-311, 475-481, 591-630)
# ----------------------------------------------------------------------------
def bev_corners(boxes: np.ndarray) -> np.ndarray:
    """misc.py:143-177 on columns (x, y, ., l, w, ., yaw): float32 corners [n, 4, 2]."""
    centers, dims, ang = boxes[:, :2], boxes[:, 3:5], boxes[:, 6]
    unit = np.array([[-0.5, -0.5], [-0.5, 0.5], [0.5, 0.5], [0.5, -0.5]], dtype=np.float32)
    c = unit[None] * dims[:, None, :]
    s, co = np.sin(ang), np.cos(ang)
    rot = np.transpose(np.array([[co, -s], [s, co]]), (2, 1, 0))
    c = c @ rot
    c = c + centers[:, None, :]
    return c.astype(np.float32)


def _ccw(p, q, r) -> bool:
    # (r.y - p.y) * (q.x - p.x) > (q.y - p.y) * (r.x - p.x), float32 arithmetic (misc.py:241-255)
    return np.float32(r[1] - p[1]) * np.float32(q[0] - p[0]) > \
        np.float32(q[1] - p[1]) * np.float32(r[0] - p[0])


def _inside_all(outer: np.ndarray, inner: np.ndarray) -> bool:
    """misc.py:267-283 (clockwise=True): every corner of ``inner`` strictly inside ``outer``."""
    for l in range(4):
        for k in range(4):
            vec = -(outer[k] - outer[(k + 1) % 4])
            cross = np.float32(vec[1]) * np.float32(outer[k, 0] - inner[l, 0])
            cross = np.float32(cross - np.float32(vec[0]) * np.float32(outer[k, 1] - inner[l, 1]))
            if cross >= 0:
                return False
    return True


def pair_collides(a: np.ndarray, b: np.ndarray) -> bool:
    """One (box, query-box) cell of box_collision_test (misc.py:203-311); corners float32 [4,2]."""
    a = a.astype(np.float32)
    b = b.astype(np.float32)
    iw = min(a[:, 0].max(), b[:, 0].max()) - max(a[:, 0].min(), b[:, 0].min())
    if not iw > 0:
        return False
    ih = min(a[:, 1].max(), b[:, 1].max()) - max(a[:, 1].min(), b[:, 1].min())
    if not ih > 0:
        return False
    for k in range(4):
        A, B = a[k], a[(k + 1) % 4]
        for l in range(4):
            Cc, Dd = b[l], b[(l + 1) % 4]
            if _ccw(A, Cc, Dd) != _ccw(B, Cc, Dd) and _ccw(A, B, Cc) != _ccw(A, B, Dd):
                return True
    # no edge crossing: containment either way
    if _inside_all(a, b):
        return True
    return _inside_all(b, a)


def check_collision(boxes: Sequence[np.ndarray]) -> bool:
    """BoxOverlap.check_collision(box, fliter=True) (misc.py:591-630): does the LAST kept box
    touch any kept box?  Boxes with x >= 63 are dropped first (misc.py:475-481)."""
    if len(boxes) == 1:
        return False
    arr = np.array(boxes, dtype=np.float64)
    arr = arr[~(arr[:, 0] >= 63)]
    if arr.shape[0] <= 1:
        return False
    flipped = arr[:, :7].copy()
    flipped[:, 6] = -flipped[:, 6]
    corners = bev_corners(flipped)
    q = corners[-1]
    return any(pair_collides(corners[i], q) for i in range(corners.shape[0]))


# ----------------------------------------------------------------------------
# Transformer blocks (module.py)
# ----------------------------------------------------------------------------
def layer_norm(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """module.py:26-37: weight only, eps 1e-5."""
    return F.layer_norm(x, (x.shape[-1],), w, None, 1e-5)


def sdpa(q, k, v, causal: bool, n_head: int) -> torch.Tensor:
    """softmax(q k^T / sqrt(d)) v with flash-attn's bottom-right causal alignment
    (module.py:214-227; flash_attn >= 2.1 semantics).  q [B,Tq,C], k/v [B,Tk,C]."""
    B, Tq, Cc = q.shape
    Tk = k.shape[1]
    d = Cc // n_head
    qh = q.view(B, Tq, n_head, d).transpose(1, 2)
    kh = k.view(B, Tk, n_head, d).transpose(1, 2)
    vh = v.view(B, Tk, n_head, d).transpose(1, 2)
    att = (qh @ kh.transpose(-1, -2)) * (1.0 / math.sqrt(d))
    if causal:
        keep = torch.ones(Tq, Tk, dtype=torch.bool).tril(diagonal=Tk - Tq)
        att = att.masked_fill(~keep, float("-inf"))
    y = torch.softmax(att, dim=-1) @ vh
    return y.transpose(1, 2).reshape(B, Tq, Cc)


def self_attention(P, pre: str, x: torch.Tensor, causal: bool, n_head: int,
                   cache: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
    """CausalFlashAttention.forward (module.py:201-230).  ``cache`` = [k, v] grown in place."""
    qkv = F.linear(x, P[pre + ".c_attn.weight"], P[pre + ".c_attn.bias"])
    q, k, v = qkv.split(x.shape[-1], dim=-1)
    if cache is not None:
        if cache[0] is not None:
            k = torch.cat([cache[0], k], dim=1)
            v = torch.cat([cache[1], v], dim=1)
        cache[0], cache[1] = k, v
    y = sdpa(q, k, v, causal, n_head)
    return F.linear(y, P[pre + ".c_proj.weight"], P[pre + ".c_proj.bias"])


def mlp(P, pre: str, x: torch.Tensor) -> torch.Tensor:
    """MLP.forward (module.py:245-250): c_fc -> erf-GELU -> c_proj, no biases."""
    return F.linear(F.gelu(F.linear(x, P[pre + ".c_fc.weight"])), P[pre + ".c_proj.weight"])


def gmlp(P, pre: str, x: torch.Tensor) -> torch.Tensor:
    """GMLP.forward_func (module.py:723-728)."""
    return mlp(P, pre, x)


def block_tar(P, pre: str, x: torch.Tensor, n_head: int) -> torch.Tensor:
    """BlockTAR.forward_func (module.py:332-359) on x [T, S, C] (batch 1)."""
    x = x + self_attention(P, pre + ".spatial_attn_1", layer_norm(x, P[pre + ".ln_1.weight"]), False, n_head)
    x = x + mlp(P, pre + ".mlp1", layer_norm(x, P[pre + ".ln_2.weight"]))
    xt = x.transpose(0, 1)                                   # [S, T, C]
    xt = xt + self_attention(P, pre + ".temporal_attn", layer_norm(xt, P[pre + ".ln_3.weight"]), True, n_head)
    xt = xt + mlp(P, pre + ".mlp2", layer_norm(xt, P[pre + ".ln_4.weight"]))
    x = xt.transpose(0, 1)
    x = x + self_attention(P, pre + ".spatial_attn_2", layer_norm(x, P[pre + ".ln_5.weight"]), False, n_head)
    x = x + mlp(P, pre + ".mlp3", layer_norm(x, P[pre + ".ln_6.weight"]))
    return x


def block_oar(P, pre: str, x: torch.Tensor, cache: List[torch.Tensor], n_head: int) -> torch.Tensor:
    """BlockOAR.forward_func (module.py:402-416) on x [1, n, C] with KV cache."""
    x = x + self_attention(P, pre + ".temporal_attn", layer_norm(x, P[pre + ".ln_1.weight"]), True, n_head, cache)
    x = x + mlp(P, pre + ".mlp", layer_norm(x, P[pre + ".ln_2.weight"]))
    return x


def ego_decoder(P, pre: str, x: torch.Tensor, p: torch.Tensor, n_head: int) -> torch.Tensor:
    """Decoder.forward_func (module.py:662-683): x [T,3,C] ego queries, p [T,S,C] scene."""
    x = x + self_attention(P, pre + ".self_attn", layer_norm(x, P[pre + ".ln_1.weight"]), False, n_head)
    qn = layer_norm(x, P[pre + ".ln_2.weight"])
    pn = layer_norm(p, P[pre + ".ln_3.weight"])
    ca = pre + ".cross_attn"
    q = F.linear(qn, P[ca + ".q_attn.weight"], P[ca + ".q_attn.bias"])
    k = F.linear(pn, P[ca + ".k_attn.weight"], P[ca + ".k_attn.bias"])
    v = F.linear(pn, P[ca + ".v_attn.weight"], P[ca + ".v_attn.bias"])
    y = sdpa(q, k, v, False, n_head)
    x = x + F.linear(y, P[ca + ".c_proj.weight"], P[ca + ".c_proj.bias"])
    x = x + mlp(P, pre + ".mlp1", layer_norm(x, P[pre + ".ln_4.weight"]))
    return x


# ----------------------------------------------------------------------------
# Embeddings and the map warp (UMGen.py:310-354, 411-515)
# ----------------------------------------------------------------------------
def bbox_spatial_embedding(P, tok: torch.Tensor) -> torch.Tensor:
    """add_spatial_pos_emb (UMGen.py:411-435): bf16(sp[x_tok] + sp[y_tok]) per slot, repeated
    over the slot's 11 attributes.  tok [T, 660] -> [T, 660, C] bf16."""
    T = tok.shape[0]
    slots = tok.view(T, N_SLOTS, N_ATTR)
    sp = P["bbox3d_spatial_posi"]
    e = sp[slots[:, :, 0]] + sp[slots[:, :, 1]]
    return e[:, :, None, :].expand(-1, -1, N_ATTR, -1).reshape(T, N_SLOTS * N_ATTR, -1)


def embed_mod(P, tokens: Dict[str, torch.Tensor], mod: str, *, bbox_pos: bool = False,
              map_grid_pos: bool = False) -> torch.Tensor:
    """get_mod_emb_pre (UMGen.py:438-468) for tokens[mod] [T, S_mod] -> [T, S_mod, C]."""
    t = tokens[mod]
    if mod == "pose":
        return P["fouier_pe"][t]
    if mod == "map":
        f = gmlp(P, "map_mlp_pre", P["map_codebook.weight"][t])
        if map_grid_pos:
            f = f + P["grid_center_posi_embedding"][None]
        return f
    if mod == "image":
        return gmlp(P, "img_mlp_pre", P["img_codebook.weight"][t])
    if mod == "bbox3d":
        f = P["transformer.be.weight"][t]
        if bbox_pos:
            f = f + bbox_spatial_embedding(P, t)
        return f
    raise ValueError(mod)


def with_bos_eos(P, feats: torch.Tensor, mod: str) -> torch.Tensor:
    """add_bos_eos (UMGen.py:470-481)."""
    axe = P["transformer.axe.weight"]
    T = feats.shape[0]
    b = axe[BOS_EOS[mod][0]].expand(T, 1, -1)
    e = axe[BOS_EOS[mod][1]].expand(T, 1, -1)
    return torch.cat([b, feats, e], dim=1)


def add_pos(P, x: torch.Tensor) -> torch.Tensor:
    """add_pos_emb (UMGen.py:483-515) with add_t_pos=True: + spe[s] + tpe[t]."""
    T, S, _ = x.shape
    return x + P["transformer.spe.weight"][:S][None] + P["transformer.tpe.weight"][:T][:, None]


def affine_warp(x: torch.Tensor, pose_diff: torch.Tensor) -> torch.Tensor:
    """affine_transform (UMGen.py:310-354): x [T, 1024, C] map features, pose_diff [T, 3]."""
    T, S, Cc = x.shape
    img = x.transpose(1, 2).reshape(T, Cc, 32, 32)
    th = pose_diff[:, 2]
    dx = 2 * (pose_diff[:, 0] / 4.0) / 32
    dy = 2 * (pose_diff[:, 1] / 4.0) / 32
    mat = torch.zeros(T, 2, 3, dtype=pose_diff.dtype)
    mat[:, 0, 0] = torch.cos(-th)
    mat[:, 0, 1] = -torch.sin(-th)
    mat[:, 0, 2] = -dy
    mat[:, 1, 0] = torch.sin(-th)
    mat[:, 1, 1] = torch.cos(-th)
    mat[:, 1, 2] = -dx
    grid = F.affine_grid(mat, (T, Cc, 32, 32), align_corners=False)
    out = F.grid_sample(img.float(), grid.float(), mode="bilinear", padding_mode="zeros",
                        align_corners=False)
    return out.reshape(T, Cc, S).transpose(1, 2).to(x.dtype)


# ----------------------------------------------------------------------------
# Samplers (UMGen.py:899-974)
# ----------------------------------------------------------------------------
@dataclass
class SampleCfg:
    method: str = "topk"          # "topk" | "topp"
    top_k: int = 5                # pose / bbox3d (config.top_k)
    top_k_map: int = 5            # config.top_k_map
    top_k_image: int = 16         # hard-coded self.topk_image (UMGen.py:103)
    p: float = 0.4
    p_map: float = 0.4
    temp: float = 1.0

    @staticmethod
    def greedy() -> "SampleCfg":
        return SampleCfg(method="topk", top_k=1, top_k_map=1, top_k_image=1)

    def param(self, mod: str):
        if self.method == "topk":
            return {"map": self.top_k_map, "image": self.top_k_image}.get(mod, self.top_k)
        # topp quirk (UMGen.py:1133): the image branch passes topk_image as p
        return {"map": self.p_map, "image": float(self.top_k_image)}.get(mod, self.p)


def sample_rows(logits: torch.Tensor, cfg: SampleCfg, param, gen: Optional[torch.Generator]) -> torch.Tensor:
    """topk (UMGen.py:899-913) / sample_top_p (UMGen.py:915-965) on logits [R, V] -> ids [R]."""
    logits = logits.clone().float()
    if cfg.method == "topk":
        k = min(int(param), logits.shape[-1])
        kth = torch.topk(logits, k).values[..., -1:]
        logits[logits < kth] = float("-inf")
        probs = torch.softmax(logits / cfg.temp, dim=-1)
        if k == 1:      # multinomial over a single non-zero bin (ties: lowest index, engine rule)
            return probs.argmax(dim=-1)
        return torch.multinomial(probs, 1, generator=gen)[:, 0]
    probs = torch.softmax(logits / cfg.temp, dim=-1)
    ps, pi = torch.sort(probs, dim=-1, descending=True)
    cum = torch.cumsum(ps, dim=-1)
    ps[(cum - ps) > float(param)] = 0.0
    ps = ps / ps.sum(dim=-1, keepdim=True)
    pick = torch.multinomial(ps, 1, generator=gen)
    return torch.gather(pi, -1, pick)[:, 0]


# ----------------------------------------------------------------------------
# Model description
# ----------------------------------------------------------------------------
@dataclass
class ModelCfg:
    n_embd: int = 768
    n_head: int = 16
    n_tar_layer: int = 36
    n_oar_layer: int = 36
    n_ego_tar_layer: int = 12
    n_ego_ca_layer: int = 12
    n_map_tar_layer: int = 24
    n_box_tar_layer: int = 24
    cond_frame: int = 20
    rule_constrain: bool = True
    merage_ar_tar: bool = True

    @staticmethod
    def large() -> "ModelCfg":
        return ModelCfg()

    @staticmethod
    def tiny(layers: int = 1) -> "ModelCfg":
        return ModelCfg(n_tar_layer=layers, n_oar_layer=layers, n_ego_tar_layer=layers,
                        n_ego_ca_layer=layers, n_map_tar_layer=layers, n_box_tar_layer=layers)


@dataclass
class FrameTrace:
    """Intermediates of one generated frame, kept for layer-by-layer parity tests."""
    ego_logits: Optional[torch.Tensor] = None        # [3, 1024]
    pose_shifted: Optional[torch.Tensor] = None      # [T, 3] pose stream fed to the TAR passes
    tar_feat: Optional[torch.Tensor] = None          # [2207, C] conditioning feature of the last frame
    tar_bbox_logits: Optional[torch.Tensor] = None   # [2207, 1028] head_tar_bbox3d(tar_feat) (bbox rows used)
    logits: Dict[int, torch.Tensor] = field(default_factory=dict)   # position p -> AR logits [V]
    tokens: Optional[torch.Tensor] = None            # [2207] full frame incl. bos/eos ids (aux ids at forced)
    cleaned_slots: List[int] = field(default_factory=list)
    tar_resampled: List[int] = field(default_factory=list)
    stream: Dict[int, int] = field(default_factory=dict)    # position p -> the id that was fed forward (later wipes rewrite `tokens`, not this)


class UMGenOracle:
    """Functional restatement of UMGen.inference (UMGen.py:1542-1671), batch 1, fp32 on CPU."""

    def __init__(self, params: Dict[str, torch.Tensor], cfg: ModelCfg, sample: Optional[SampleCfg] = None,
                 seed: int = 0):
        self.P = params
        self.cfg = cfg
        self.sample = sample or SampleCfg()
        self.gen = torch.Generator().manual_seed(seed)
        self.box_lut = box_value_lut()
        self.trace: List[FrameTrace] = []
        self.keep_trace = False

    # -- TAR side -------------------------------------------------------------------------
    def _stack(self, name: str, n_layer: int, x: torch.Tensor, ln: str) -> torch.Tensor:
        for i in range(n_layer):
            x = block_tar(self.P, f"transformer.{name}.{i}", x, self.cfg.n_head)
        return layer_norm(x, self.P[f"transformer.{ln}.weight"])

    def tar_inputs(self, tokens: Dict[str, torch.Tensor], mods: Sequence[str], *, map_grid_pos: bool,
                   warp: bool, bbox_pos: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """Input embedding of forward_tar_for_map / _for_box / forward_tar_net / forward_ego_net
        (UMGen.py:691-872, 634-663).  Returns ([T, S', C], map_warped or None)."""
        P = self.P
        parts, warped = [], None
        for m in mods:
            if m == "map":
                f = embed_mod(P, tokens, "map", map_grid_pos=map_grid_pos)
                if warp:
                    warped = affine_warp(f, decode_pose(tokens["pose"]))
                    f = warped + f
            else:
                f = embed_mod(P, tokens, m, bbox_pos=bbox_pos)
            parts.append(with_bos_eos(P, f.float(), m))
        return add_pos(P, torch.cat(parts, dim=1)), warped

    def ego_net(self, tokens: Dict[str, torch.Tensor]) -> torch.Tensor:
        """infer_ego_net / forward_ego_net (UMGen.py:994-1005, 634-687): logits [3, 1024] of the
        last frame's three ego queries."""
        P, cfg = self.P, self.cfg
        x, _ = self.tar_inputs(tokens, MODS, map_grid_pos=False, warp=False)
        scene = self._stack("ego_tar", cfg.n_ego_tar_layer, x, "ln_ego_tar")
        T = scene.shape[0]
        q = add_pos(P, P["transformer.egoe.weight"][None].expand(T, -1, -1))
        # decoders never mix frames and only the last frame is read (UMGen.py:1002)
        q, scene = q[-1:], scene[-1:]
        for i in range(cfg.n_ego_ca_layer):
            q = ego_decoder(P, f"transformer.ego_cross_attn.{i}", q, scene, cfg.n_head)
        q = layer_norm(q, P["transformer.ln_ego.weight"])
        return F.linear(q[0], P["transformer.head_ego.weight"])

    def tar_feature(self, tokens: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Step 2 of _inference (UMGen.py:1482-1511): the [2207, C] feature of the LAST frame that
        conditions the OAR.  pose/image rows come from TAR, map rows from map_tar plus the warped-map
        prior of the map pass, bbox3d rows from box_tar."""
        cfg = self.cfg
        off = mod_offsets(MODS)
        x, warped = self.tar_inputs(tokens, TASKS["pose_map"], map_grid_pos=False, warp=True)
        f_map = self._stack("map_tar", cfg.n_map_tar_layer, x, "ln_map_tar")[-1]
        x, _ = self.tar_inputs(tokens, TASKS["pose_map_bbox3d"], map_grid_pos=False, warp=True)
        f_box = self._stack("box_tar", cfg.n_box_tar_layer, x, "ln_box_tar")[-1]
        x, _ = self.tar_inputs(tokens, MODS, map_grid_pos=True, warp=True)
        f_all = self._stack("TAR", cfg.n_tar_layer, x, "ln_tar")[-1]
        feat = f_all.clone()
        m0 = off["map"]
        feat[m0:m0 + TOKEN_LEN["map"]] = f_map[m0:m0 + TOKEN_LEN["map"]]
        feat[m0 + 1:m0 + 1 + CONTENT_LEN["map"]] += warped[-1]
        b0 = off["bbox3d"]
        feat[b0:b0 + TOKEN_LEN["bbox3d"]] = f_box[b0:b0 + TOKEN_LEN["bbox3d"]]
        return feat

    # -- OAR side -------------------------------------------------------------------------
    def embed_token(self, q: int, tok: int) -> torch.Tensor:
        """Embedding of new-frame token at 1-indexed position q as appended to ``out_tokens``
        (UMGen.py:1046-1137): bos/eos -> axe, pose -> fouier_pe, map/image -> codebook+GMLP, bbox -> be
        (no spatial sinusoid, no spe/tpe in the OAR)."""
        P = self.P
        forced = forced_positions()
        if q in forced:
            return P["transformer.axe.weight"][forced[q]].float()
        m = pos_mod(q)
        t = torch.tensor([tok])
        if m == "pose":
            return P["fouier_pe"][t][0].float()
        if m == "map":
            return gmlp(P, "map_mlp_pre", P["map_codebook.weight"][t])[0]
        if m == "image":
            return gmlp(P, "img_mlp_pre", P["img_codebook.weight"][t])[0]
        return P["transformer.be.weight"][t][0]

    def oar_frame(self, tar_feat: torch.Tensor, pose_tokens: torch.Tensor, prev_bbox: torch.Tensor,
                  control_slots: Optional[Sequence[int]] = None, teacher: Optional[torch.Tensor] = None,
                  trace: Optional[FrameTrace] = None, max_pos: int = SEQ_LEN,
                  given: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        """infer_oar_net + sample_next_token + rule_based_constraint (UMGen.py:1151-1383).

        tar_feat [2207, C]; pose_tokens [3] (given); prev_bbox [660] = last conditioning frame's box
        tokens; teacher [2207] optional full-frame token ids to force (parity tests); given = init
        tokens of further modalities ({"map": [1024]} or map + bbox3d) that extend the forced prefix
        (UMGen.py:1184-1201: causal prefill, no sampling and no rule check for them).  Returns the
        frame's [2207] ids (bos/eos positions hold the aux id)."""
        P, cfg, sc = self.P, self.cfg, self.sample
        forced = forced_positions()
        head = {"map": "transformer.head_ar_map.weight", "bbox3d": "transformer.head_ar_bbox3d.weight",
                "image": "transformer.head_ar_img.weight"}
        out = torch.zeros(SEQ_LEN + 1, dtype=torch.long)             # 1-indexed
        out[1], out[5] = BOS_EOS["pose"]
        out[2:5] = pose_tokens.long()
        prefix_end = 5
        off = mod_offsets(MODS)
        for m in ("map", "bbox3d"):
            if given and given.get(m) is not None:
                assert prefix_end == off[m], "given modalities must be a contiguous prefix of the frame"
                out[off[m] + 1], out[off[m] + TOKEN_LEN[m]] = BOS_EOS[m]
                out[off[m] + 2: off[m] + 2 + CONTENT_LEN[m]] = given[m].long().view(-1)
                prefix_end = off[m] + TOKEN_LEN[m]
        caches = [[None, None] for _ in range(cfg.n_oar_layer)]
        decoded_boxes: List[np.ndarray] = []
        ln_w = P["transformer.ln_oar.weight"]
        w_tar_bbox = P["transformer.head_tar_bbox3d.weight"]

        def run(x):
            for i in range(cfg.n_oar_layer):
                x = block_oar(P, f"transformer.OAR.{i}", x, caches[i], cfg.n_head)
            return layer_norm(x, ln_w)

        # positions 1..5 are given -> the first pass is a 6-token causal prefill (UMGen.py:1234-1235)
        emb = [P["transformer.tske.weight"][TASK_ID].float()] + [self.embed_token(q, int(out[q])) for q in range(1, 6)]
        x = torch.stack(emb)[None] + tar_feat[None, :6]
        h = run(x)[0, -1]
        for p in range(6, max_pos + 1):
            if p > 6:
                x = (self.embed_token(p - 1, int(out[p - 1])) + tar_feat[p - 1])[None, None]
                h = run(x)[0, -1]
            if p in forced:
                out[p] = forced[p]
                continue
            if p <= prefix_end:                                        # given token: fed forward as is
                continue
            m = pos_mod(p)
            logits = F.linear(h, P[head[m]])
            if trace is not None:
                trace.logits[p] = logits.clone()
            tok = int(sample_rows(logits[None], sc, sc.param(m), self.gen)[0])
            if m == "bbox3d":
                bidx = p - BBOX_FIRST_POS - 1                          # UMGen.py:1076-1081
                prev = int(prev_bbox[bidx])
                tar_logits = None
                if control_slots is not None and (p - BBOX_FIRST_POS) // N_ATTR in control_slots:
                    tar_logits = F.linear(tar_feat[p - 1], w_tar_bbox)   # UMGen.py:1083-1089
                    tar_logits[-1] = float("-inf")
                    tok = int(sample_rows(tar_logits[None], sc, sc.param(m), self.gen)[0])
                if tok == PAD_TOKEN and cfg.merage_ar_tar and prev != PAD_TOKEN:   # UMGen.py:1092-1104
                    if tar_logits is None:
                        tar_logits = F.linear(tar_feat[p - 1], w_tar_bbox)
                    tok = int(sample_rows(tar_logits[None], sc, sc.param(m), self.gen)[0])
                    if trace is not None:
                        trace.tar_resampled.append(p)
                if cfg.rule_constrain and tok != PAD_TOKEN and (p - BBOX_FIRST_POS) % N_ATTR == 0:
                    # UMGen.py:1275-1383 at the slot's 11th (category) token
                    attr = out[p - 10:p].numpy()
                    box = self.box_lut[attr, np.arange(10)]
                    if not decoded_boxes:
                        decoded_boxes.append(np.array([0, 0, 0, *EGO_LWH, 0, 0, 0, 0], dtype=np.float64))
                    decoded_boxes.append(box)
                    hit = check_collision(decoded_boxes)
                    was_pad = prev == PAD_TOKEN
                    if was_pad and (hit or len(decoded_boxes) > 30):
                        out[p - 10:p] = PAD_TOKEN                      # KV cache stays stale (quirk)
                        tok = PAD_TOKEN
                        decoded_boxes.pop()
                        if trace is not None:
                            trace.cleaned_slots.append((p - BBOX_FIRST_POS) // N_ATTR - 1)
            out[p] = tok
            if trace is not None:
                trace.stream[p] = tok
            if teacher is not None:
                if trace is not None:
                    trace.logits[-p] = torch.tensor(tok)               # what the oracle itself picked
                out[p] = int(teacher[p - 1])
        return out[1:]

    # -- frame loop -----------------------------------------------------------------------
    def frame(self, cond: Dict[str, torch.Tensor], init: Optional[Dict[str, torch.Tensor]] = None,
              control_test: bool = False, teacher: Optional[torch.Tensor] = None,
              max_pos: int = SEQ_LEN) -> Dict[str, torch.Tensor]:
        """_inference (UMGen.py:1406-1540) for cond tokens {mod: [T, S_mod]}; returns {mod: [S_mod]}.
        ``init`` = this frame's slice of the control dict ({pose: [3], bbox3d: [660] with -1 = free})."""
        tr = FrameTrace() if self.keep_trace else None
        cond = dict(cond)
        if init is not None and init.get("pose") is not None:
            pose_new = init["pose"].long().view(3)
        else:
            ego_logits = self.ego_net(cond)
            pose_new = sample_rows(ego_logits, self.sample, self.sample.param("pose"), self.gen)
            if tr is not None:
                tr.ego_logits = ego_logits
        cond["pose"] = torch.cat([cond["pose"], pose_new[None]], dim=0)[1:]      # UMGen.py:1445-1452
        control_slots = None
        if control_test and init is not None and init.get("bbox3d") is not None:
            valid = init["bbox3d"].view(-1) != -1                               # UMGen.py:1464-1472
            cond["bbox3d"][-1, valid] = init["bbox3d"].view(-1)[valid]          # in place, as the reference
            control_slots = set(np.where(valid.view(N_SLOTS, -1).any(dim=1).numpy())[0].tolist())
        feat = self.tar_feature(cond)
        if tr is not None:
            tr.pose_shifted = cond["pose"].clone()
            tr.tar_feat = feat
        given = None
        if init is not None:                                                    # UMGen.py:1474, 1515-1523, 1184-1201
            given = {m: init[m] for m in ("map", "bbox3d") if init.get(m) is not None and not (control_test and m == "bbox3d")}
        ids = self.oar_frame(feat, pose_new, cond["bbox3d"][-1], control_slots, teacher, tr, max_pos, given=given)
        if tr is not None:
            tr.tokens = ids
            self.trace.append(tr)
        off = mod_offsets(MODS)
        return {m: ids[off[m] + 1: off[m] + 1 + CONTENT_LEN[m]] for m in MODS}

    def inference(self, new_frames: int, cond_frames: int, input_cond_frames: int,
                  input_cond_tokens: Dict[str, torch.Tensor], init_tokens: Optional[Dict[str, torch.Tensor]] = None,
                  control_test: bool = False) -> Dict[str, np.ndarray]:
        """UMGen.inference (UMGen.py:1542-1671).  Tokens in/out carry the leading batch-1 axis."""
        if input_cond_frames == -1:
            input_cond_frames = cond_frames
        out = {m: input_cond_tokens[m][0, :input_cond_frames].clone() for m in MODS}
        cond = {m: input_cond_tokens[m][0, :input_cond_frames].clone() for m in MODS}
        for idx in range(new_frames):
            if cond["pose"].shape[0] > cond_frames:
                cond = {m: cond[m][-cond_frames:].clone() for m in MODS}
            init = None
            if init_tokens is not None:
                init = {m: (v[0, idx] if idx < v.shape[1] else None) for m, v in init_tokens.items()}
                if "pose" in init and init["pose"] is None:                      # UMGen.py:1613-1619
                    init_tokens, control_test, init = None, False, None
            new = self.frame(cond, init, control_test)
            for m in MODS:
                use_init = init_tokens is not None and m in init_tokens and not (control_test and m == "bbox3d")
                row = init[m].long().view(-1) if use_init else new[m]
                cond[m] = torch.cat([cond[m], row[None]], dim=0)
                out[m] = torch.cat([out[m], row[None]], dim=0)
        return {m: out[m][None].numpy() for m in MODS}
