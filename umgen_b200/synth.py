"""Seeded synthetic assets: a UMGen checkpoint with the reference's state_dict keys, tokenised
scenes and control dicts (SURVEY.md section 8d -- the released weights/scenes are not available
offline).

Every tensor is generated independently from ``(seed, key)`` so any subset (one layer, one head)
can be produced without materialising the 2.4 B-parameter model, and so the GPU box, the build
container and the golden-vector script all see bit-identical values.  Values follow PyTorch's
default initialisers (the reference never runs its ``_init_weights``, UMGen.py:274-285) and are
then rounded to fp16-representable numbers: the engine stores matrices in fp16 (the reference's
autocast dtype), so rounding at generation time makes that storage lossless and keeps greedy
token-id parity a statement about arithmetic, not about weight quantisation.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

from .config import ModelConfig, N_EMBD, N_SLOTS, N_ATTR, PAD_TOKEN

C = N_EMBD


def _attn_specs(pre: str) -> List[Tuple[str, Tuple[int, ...], str]]:
    return [(pre + ".scale", (), "scale"),
            (pre + ".c_attn.weight", (3 * C, C), "linear"), (pre + ".c_attn.bias", (3 * C,), "bias"),
            (pre + ".c_proj.weight", (C, C), "linear"), (pre + ".c_proj.bias", (C,), "bias")]


def _mlp_specs(pre: str):
    return [(pre + ".c_fc.weight", (4 * C, C), "linear"), (pre + ".c_proj.weight", (C, 4 * C), "linear")]


def _tar_block_specs(pre: str):
    s = [(pre + ".ln_1.weight", (C,), "ln")] + _attn_specs(pre + ".spatial_attn_1")
    s += [(pre + ".ln_2.weight", (C,), "ln")] + _mlp_specs(pre + ".mlp1")
    s += [(pre + ".ln_3.weight", (C,), "ln")] + _attn_specs(pre + ".temporal_attn")
    s += [(pre + ".ln_4.weight", (C,), "ln")] + _mlp_specs(pre + ".mlp2")
    s += [(pre + ".ln_5.weight", (C,), "ln")] + _attn_specs(pre + ".spatial_attn_2")
    s += [(pre + ".ln_6.weight", (C,), "ln")] + _mlp_specs(pre + ".mlp3")
    return s


def _oar_block_specs(pre: str):
    return ([(pre + ".ln_1.weight", (C,), "ln")] + _attn_specs(pre + ".temporal_attn")
            + [(pre + ".ln_2.weight", (C,), "ln")] + _mlp_specs(pre + ".mlp"))


def _ego_decoder_specs(pre: str):
    s = [(pre + ".ln_1.weight", (C,), "ln")] + _attn_specs(pre + ".self_attn")
    s += [(pre + ".ln_2.weight", (C,), "ln"), (pre + ".ln_3.weight", (C,), "ln"),
          (pre + ".cross_attn.scale", (), "scale")]
    for n in ("q_attn", "k_attn", "v_attn", "c_proj"):
        s += [(f"{pre}.cross_attn.{n}.weight", (C, C), "linear"), (f"{pre}.cross_attn.{n}.bias", (C,), "bias")]
    s += [(pre + ".ln_4.weight", (C,), "ln")] + _mlp_specs(pre + ".mlp1")
    return s


def param_specs(cfg: ModelConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) for every entry of the reference model's state_dict (UMGen.py:137-261)."""
    t = "transformer."
    s: List[Tuple[str, Tuple[int, ...], str]] = [
        ("fouier_pe", (1024, C), "sin0"), ("bbox3d_spatial_posi", (1030, C), "sin1024"),
        ("grid_center_posi_embedding", (1024, C), "gridpos"),
        ("map_mlp_pre.c_fc.weight", (4 * C, 16), "linear"), ("map_mlp_pre.c_proj.weight", (C, 4 * C), "linear"),
        ("img_mlp_pre.c_fc.weight", (4 * C, 16), "linear"), ("img_mlp_pre.c_proj.weight", (C, 4 * C), "linear"),
        (t + "egoe.weight", (3, C), "emb"), (t + "axe.weight", (8, C), "emb"), (t + "be.weight", (1028, C), "emb"),
        (t + "tpe.weight", (cfg.max_frame_len, C), "emb"), (t + "spe.weight", (2207, C), "emb"),
        (t + "tske.weight", (7, C), "emb"),
    ]
    for i in range(cfg.n_ego_tar_layer):
        s += _tar_block_specs(f"{t}ego_tar.{i}")
    s += [(t + "ln_ego_tar.weight", (C,), "ln"), (t + "ln_ego.weight", (C,), "ln")]
    for i in range(cfg.n_tar_layer):
        s += _tar_block_specs(f"{t}TAR.{i}")
    for i in range(cfg.n_oar_layer):
        s += _oar_block_specs(f"{t}OAR.{i}")
    s += [(t + "ln_tar.weight", (C,), "ln"), (t + "ln_oar.weight", (C,), "ln")]
    for name, v in (("head_tar_aux", 8), ("head_tar_pose", 1024), ("head_tar_map", 8192), ("head_ar_aux", 8),
                    ("head_ar_pose", 1024), ("head_ar_map", 8192), ("head_ar_bbox3d", 1028)):
        s.append((f"{t}{name}.weight", (v, C), "linear"))
    for i in range(cfg.n_ego_ca_layer):
        s += _ego_decoder_specs(f"{t}ego_cross_attn.{i}")
    s += [(t + "head_ego.weight", (1024, C), "linear"), (t + "head_tar_bbox3d.weight", (1028, C), "linear")]
    for i in range(cfg.n_map_tar_layer):
        s += _tar_block_specs(f"{t}map_tar.{i}")
    s += [(t + "ln_map_tar.weight", (C,), "ln"),
          (t + "head_tar_img.weight", (8192, C), "linear"), (t + "head_ar_img.weight", (8192, C), "linear")]
    for i in range(cfg.n_box_tar_layer):
        s += _tar_block_specs(f"{t}box_tar.{i}")
    s += [(t + "ln_box_tar.weight", (C,), "ln"),
          ("map_codebook.weight", (8192, 16), "codebook"), ("img_codebook.weight", (8192, 16), "codebook")]
    return s


def sinusoid_table(n_position: int, emb_dim: int, start_index: int = 0) -> torch.Tensor:
    """Fixed sinusoid table of module.py:746-768 (row 0 zero), bfloat16."""
    pos = np.arange(n_position, dtype=np.float64)[:, None] + start_index
    j = np.arange(emb_dim)
    ang = pos / np.power(10000.0, 2.0 * (j // 2) / emb_dim)[None, :]
    tab = np.zeros((n_position, emb_dim), dtype=np.float64)
    tab[1:, 0::2] = np.sin(ang[1:, 0::2])
    tab[1:, 1::2] = np.cos(ang[1:, 1::2])
    return torch.from_numpy(tab).to(torch.bfloat16)


def grid_center_embedding(spatial: torch.Tensor) -> torch.Tensor:
    """UMGen.py:140-153: sinusoid rows of the digitised centre of each 4 m map cell, x + y, bf16."""
    idx = np.arange(32, dtype=np.float32)
    c = -((idx + 0.5) * 4.0 - 64.0)
    gx, gy = np.meshgrid(c, c, indexing="ij")
    bins = np.linspace(0.0, 1.0, 1024)
    tx = np.digitize((gx + 64.0) / 128.0, bins).reshape(1024)
    ty = np.digitize((gy + 64.0) / 128.0, bins).reshape(1024)
    return spatial[torch.from_numpy(tx)] + spatial[torch.from_numpy(ty)]


def _gen(seed: int, key: str) -> torch.Generator:
    return torch.Generator().manual_seed((zlib.crc32(key.encode()) * 2654435761 + seed * 97 + 1) % (2 ** 63))


def _fp16_exact(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float16).to(torch.float32)


def make_param(key: str, shape: Tuple[int, ...], kind: str, seed: int) -> torch.Tensor:
    g = _gen(seed, key)
    if kind == "linear":
        b = 1.0 / math.sqrt(shape[1])
        return _fp16_exact((torch.rand(shape, generator=g) * 2 - 1) * b)
    if kind == "bias":
        b = 1.0 / math.sqrt(C)
        return _fp16_exact((torch.rand(shape, generator=g) * 2 - 1) * b)
    if kind == "emb":
        return _fp16_exact(torch.randn(shape, generator=g))
    if kind == "ln":
        return _fp16_exact(1.0 + 0.1 * torch.randn(shape, generator=g))
    if kind == "scale":
        return torch.tensor(1.0 / math.sqrt(48.0))
    if kind == "codebook":
        w = torch.randn(shape, generator=g)
        return _fp16_exact(w / w.norm(dim=1, keepdim=True))
    if kind == "sin0":
        return sinusoid_table(shape[0], shape[1], 0)
    if kind == "sin1024":
        return sinusoid_table(shape[0], shape[1], 1024)
    if kind == "gridpos":
        return grid_center_embedding(sinusoid_table(1030, shape[1], 1024))
    raise ValueError(kind)


def make_state_dict(cfg: ModelConfig, seed: int = 0, keys: Optional[Iterable[str]] = None) -> Dict[str, torch.Tensor]:
    want = set(keys) if keys is not None else None
    return {k: make_param(k, shp, kind, seed) for k, shp, kind in param_specs(cfg)
            if want is None or k in want}


class DeviceParams(dict):
    """state_dict-like mapping whose random tensors are drawn directly on a CUDA device on first access and
    NOT cached (each is read once by the weight packers).  Same distributions as make_param, different
    values: for benchmarks of the full-size model, where parity against CPU-generated goldens is not needed."""

    def __init__(self, cfg: ModelConfig, seed: int = 0, device="cuda:0"):
        super().__init__()
        self._spec = {k: (shp, kind) for k, shp, kind in param_specs(cfg)}
        self._seed, self._dev = seed, torch.device(device)

    def __contains__(self, key):
        return key in self._spec

    def __missing__(self, key):
        shp, kind = self._spec[key]
        if kind in ("sin0", "sin1024", "gridpos", "scale"):
            return make_param(key, shp, kind, self._seed).to(self._dev)
        g = torch.Generator(device=self._dev).manual_seed((zlib.crc32(key.encode()) + self._seed * 97 + 1) % (2 ** 63))
        if kind == "linear":
            return _fp16_exact((torch.rand(shp, generator=g, device=self._dev) * 2 - 1) / math.sqrt(shp[1]))
        if kind == "bias":
            return _fp16_exact((torch.rand(shp, generator=g, device=self._dev) * 2 - 1) / math.sqrt(C))
        if kind == "emb":
            return _fp16_exact(torch.randn(shp, generator=g, device=self._dev))
        if kind == "ln":
            return _fp16_exact(1.0 + 0.1 * torch.randn(shp, generator=g, device=self._dev))
        if kind == "codebook":
            w = torch.randn(shp, generator=g, device=self._dev)
            return _fp16_exact(w / w.norm(dim=1, keepdim=True))
        raise ValueError(kind)


class LazyParams(dict):
    """state_dict-like mapping that materialises a tensor on first access (for bounded CPU samples)."""

    def __init__(self, cfg: ModelConfig, seed: int = 0):
        super().__init__()
        self._spec = {k: (shp, kind) for k, shp, kind in param_specs(cfg)}
        self._seed = seed

    def __missing__(self, key):
        shp, kind = self._spec[key]
        v = make_param(key, shp, kind, self._seed)
        self[key] = v
        return v


def evaluation_namespace(cfg: Optional[ModelConfig] = None, **over):
    """The config Namespace evaluate.py hands to UMGen(config) (configs/UMGen_config_evaluation.py:344-430 after tools/infer_fun.py:84-159),
    reduced to the fields the drop-in module reads; depths from `cfg`, codebooks synthetic unless paths are given."""
    import argparse
    cfg = cfg or ModelConfig.large()
    ns = argparse.Namespace(
        task={"pose_map_bbox3d_image": ["pose", "map", "bbox3d", "image"], "pose_map_bbox3d": ["pose", "map", "bbox3d"], "pose_map": ["pose", "map"]},
        task_name_id={"pose_map_bbox3d_image": 6}, task_num=7, token_len={"pose": 5, "map": 1026, "bbox3d": 662, "image": 514},
        seq_len=2207, bos_eos={"pose": [0, 1], "map": [2, 3], "bbox3d": [4, 5], "image": [6, 7]}, cond_frame=cfg.cond_frame,
        max_frame_len=cfg.max_frame_len, sfmx_temp=1.0, top_k=5, top_k_map=5, p=0.4, sample_method="topk", rule_constrain=cfg.rule_constrain,
        merage_ar_tar=cfg.merage_ar_tar, n_embd=768, n_head=16, n_tar_layer=cfg.n_tar_layer, n_oar_layer=cfg.n_oar_layer,
        n_ego_tar_layer=cfg.n_ego_tar_layer, n_ego_ca_layer=cfg.n_ego_ca_layer, n_map_tar_layer=cfg.n_map_tar_layer,
        n_box_tar_layer=cfg.n_box_tar_layer, split_map_tar=True, split_box_tar=True, sample_img=True, map_transform=True,
        device_set=torch.device("cpu"), map_codebook=None, img_codebook=None)
    for k, v in over.items():
        setattr(ns, k, v)
    return ns


# ----------------------------------------------------------------------------------------
def make_scene(seed: int = 1, n_frames: int = 50, min_alive: int = 5, max_alive: int = 20) -> Dict[str, torch.Tensor]:
    """Synthetic tokenised scene with the shapes a DataLoader(batch_size=1) hands to the model
    (SURVEY.md section 3.6 / 8d): int64 pose [1,T,3], map [1,T,1024], bbox3d [1,T,660], image [1,T,512]."""
    g = torch.Generator().manual_seed(seed)
    T = n_frames
    pose = torch.randint(0, 1024, (1, T, 3), generator=g)
    mp = torch.randint(0, 8192, (1, T, 1024), generator=g)
    img = torch.randint(0, 8192, (1, T, 512), generator=g)
    box = torch.full((1, T, N_SLOTS, N_ATTR), PAD_TOKEN, dtype=torch.long)
    alive = int(torch.randint(min_alive, max_alive + 1, (1,), generator=g))
    box[0, :, :alive, :10] = torch.randint(100, 901, (T, alive, 10), generator=g)
    box[0, :, :alive, 10] = torch.randint(1024, 1027, (alive,), generator=g)[None, :].expand(T, -1)
    return {"pose": pose, "map": mp, "bbox3d": box.view(1, T, N_SLOTS * N_ATTR), "image": img}


def make_control(seed: int = 1, n_frames: int = 30, slot: int = 2) -> Dict[str, torch.Tensor]:
    """Control dict: forced ego poses plus one agent slot following a lateral
    shift; -1 marks free bbox3d tokens (UMGen.py:1464)."""
    g = torch.Generator().manual_seed(seed + 1000)
    pose = torch.randint(400, 624, (1, n_frames, 3), generator=g)
    box = torch.full((1, n_frames, N_SLOTS, N_ATTR), -1, dtype=torch.long)
    base = torch.randint(300, 700, (10,), generator=g)
    for t in range(n_frames):
        attrs = base.clone()
        attrs[1] = base[1] + 4 * t
        box[0, t, slot, :10] = attrs
        box[0, t, slot, 10] = 1024
    return {"pose": pose, "bbox3d": box.view(1, n_frames, N_SLOTS * N_ATTR)}


# ----------------------------------------------------------------------------------------
def make_vq_state_dict(kind: str, seed: int = 1) -> Dict[str, torch.Tensor]:
    """Seeded synthetic checkpoint of one VQ decoder ("map" | "image") with the reference's state_dict keys
    (tokenizer/vq_model.py, vq_modules.py): conv weights/biases ~ U(+-1/sqrt(fan_in)) (PyTorch default), GroupNorm
    affine ~ (1 + 0.1 n, 0.05 n), L2-normalised codebook; everything fp16-representable."""
    from .vq import VQ_CONFIGS
    cfg = VQ_CONFIGS[kind]
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k):
        g = _gen(seed, f"{kind}.{name}")
        b = 1.0 / math.sqrt(cin * k * k)
        sd[name + ".weight"] = _fp16_exact((torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * b)
        sd[name + ".bias"] = _fp16_exact((torch.rand(cout, generator=g) * 2 - 1) * b)

    def norm(name, c):
        g = _gen(seed, f"{kind}.{name}")
        sd[name + ".weight"] = _fp16_exact(1.0 + 0.1 * torch.randn(c, generator=g))
        sd[name + ".bias"] = _fp16_exact(0.05 * torch.randn(c, generator=g))

    def res(pre, cin, cout):
        norm(pre + ".norm1", cin); conv(pre + ".conv1", cout, cin, 3); norm(pre + ".norm2", cout); conv(pre + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(pre + ".nin_shortcut", cout, cin, 1)

    def attn(pre, c):
        norm(pre + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(f"{pre}.{n}", c, c, 1)

    g = _gen(seed, f"{kind}.codebook")
    w = torch.randn(8192, 16, generator=g)
    sd["quantize.embedding.weight"] = _fp16_exact(w / w.norm(dim=1, keepdim=True))
    conv("post_quant_conv", cfg["z_channels"], 16, cfg["post_quant_kernel"])
    nres = len(cfg["ch_mult"])
    block_in = cfg["ch"] * cfg["ch_mult"][-1]
    curr = cfg["resolution"] // 2 ** (nres - 1)
    conv("decoder.conv_in", block_in, cfg["z_channels"], 3)
    res("decoder.mid.block_1", block_in, block_in); attn("decoder.mid.attn_1", block_in); res("decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(nres)):
        block_out = cfg["ch"] * cfg["ch_mult"][lvl]
        for ib in range(cfg["num_res_blocks"] + 1):
            res(f"decoder.up.{lvl}.block.{ib}", block_in, block_out)
            block_in = block_out
            if curr in cfg["attn_resolutions"]:
                attn(f"decoder.up.{lvl}.attn.{ib}", block_in)
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3)
            curr *= 2
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", cfg["out_ch"], block_in, 3)
    return sd
