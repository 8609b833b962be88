"""Host side of the TAR encoders and the ego-action head: sequences the sm_100a kernels of
``csrc/gemm_sm100.cu`` / ``csrc/tar.cu`` exactly as the reference's forward passes do.

reference (paths relative to /root/reference/projects):
  models/UMGen.py:634-687   forward_ego_net          models/module.py:332-359  BlockTAR.forward_func
  models/UMGen.py:691-872   forward_tar_{net,for_map,for_box}   module.py:662-683 Decoder.forward_func
  models/UMGen.py:994-1005  infer_ego_net            models/UMGen.py:1482-1511 cascade + tar_emb assembly

Three ways to run a stack over a window of T frames (run_stack): "full" (all frames), "prefix" (the first frames of a window whose last
frame is not known yet; keeps the temporal qkv of every layer) and "suffix" (the last frame only, against those caches) -- spatial
attention is per frame and temporal attention causal over frames, so prefix + suffix is the same arithmetic as full.
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional

import torch

from . import capi, ops
from .config import ModelConfig, SampleConfig, SEQ_LEN
from .weights import pose_value_lut, token_table

C = 768
SUBS = (("ln_1", "spatial_attn_1", "ln_2", "mlp1", "spatial"),
        ("ln_3", "temporal_attn", "ln_4", "mlp2", "temporal"),
        ("ln_5", "spatial_attn_2", "ln_6", "mlp3", "spatial"))
TASK_S = {2: 1031, 3: 1693, 4: 2207}


def _h(t, dev):
    return t.detach().to(device=dev, dtype=torch.float16).contiguous()


def _f(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def pack_tar_block(sd: Mapping[str, torch.Tensor], pre: str, dev):
    out = []
    for ln_a, attn, ln_b, mlp, kind in SUBS:
        out.append(dict(
            kind=kind, ln_a=_f(sd[f"{pre}.{ln_a}.weight"], dev), w_qkv=_h(sd[f"{pre}.{attn}.c_attn.weight"], dev),
            b_qkv=_f(sd[f"{pre}.{attn}.c_attn.bias"], dev), w_proj=_h(sd[f"{pre}.{attn}.c_proj.weight"], dev),
            b_proj=_f(sd[f"{pre}.{attn}.c_proj.bias"], dev), ln_b=_f(sd[f"{pre}.{ln_b}.weight"], dev),
            w_fc=_h(sd[f"{pre}.{mlp}.c_fc.weight"], dev), w_proj2=_h(sd[f"{pre}.{mlp}.c_proj.weight"], dev)))
    return out


def pack_ego_decoder(sd, pre: str, dev):
    d = {f"ln_{i}": _f(sd[f"{pre}.ln_{i}.weight"], dev) for i in (1, 2, 3, 4)}
    d.update(w_qkv=_h(sd[f"{pre}.self_attn.c_attn.weight"], dev), b_qkv=_f(sd[f"{pre}.self_attn.c_attn.bias"], dev),
             w_proj=_h(sd[f"{pre}.self_attn.c_proj.weight"], dev), b_proj=_f(sd[f"{pre}.self_attn.c_proj.bias"], dev),
             w_fc=_h(sd[f"{pre}.mlp1.c_fc.weight"], dev), w_proj2=_h(sd[f"{pre}.mlp1.c_proj.weight"], dev))
    for n in ("q_attn", "k_attn", "v_attn", "c_proj"):
        d[f"w_{n}"] = _h(sd[f"{pre}.cross_attn.{n}.weight"], dev)
        d[f"b_{n}"] = _f(sd[f"{pre}.cross_attn.{n}.bias"], dev)
    return d


class TarEncoders:
    """Device-resident TAR stacks + ego head.  All activations live in buffers allocated once for the
    largest pass (cond_frame x 2207 rows)."""

    def __init__(self, sd: Mapping[str, torch.Tensor], cfg: ModelConfig, device="cuda:0", max_frames: Optional[int] = None):
        if not torch.cuda.is_available():
            raise capi.UmgenError("umgen_b200 needs a CUDA device (sm_100a); there is no CPU path")
        capi.lib()
        self.cfg, self.dev = cfg, torch.device(device)
        dev = self.dev
        capi.preload(dev)
        t = "transformer."
        self.stacks = {
            "ego_tar": [pack_tar_block(sd, f"{t}ego_tar.{i}", dev) for i in range(cfg.n_ego_tar_layer)],
            "map_tar": [pack_tar_block(sd, f"{t}map_tar.{i}", dev) for i in range(cfg.n_map_tar_layer)],
            "box_tar": [pack_tar_block(sd, f"{t}box_tar.{i}", dev) for i in range(cfg.n_box_tar_layer)],
            "TAR": [pack_tar_block(sd, f"{t}TAR.{i}", dev) for i in range(cfg.n_tar_layer)],
        }
        self.ego_dec = [pack_ego_decoder(sd, f"{t}ego_cross_attn.{i}", dev) for i in range(cfg.n_ego_ca_layer)]
        self.ln = {k: _f(sd[f"{t}{k}.weight"], dev) for k in ("ln_ego_tar", "ln_ego", "ln_tar", "ln_map_tar", "ln_box_tar")}
        self.head_ego = _h(sd[t + "head_ego.weight"], dev)
        self.egoe = _f(sd[t + "egoe.weight"], dev)
        self.tables = dict(fpe_f=_f(sd["fouier_pe"], dev), img_table_f=token_table(sd, "img", dev), be_f=_f(sd[t + "be.weight"], dev),
                           axe_f=_f(sd[t + "axe.weight"], dev), spe_f=_f(sd[t + "spe.weight"], dev), tpe_f=_f(sd[t + "tpe.weight"], dev),
                           spatial_f=_f(sd["bbox3d_spatial_posi"], dev))
        self.map_table = token_table(sd, "map", dev)
        self.grid_pos = _f(sd["grid_center_posi_embedding"], dev)
        self.pose_lut = torch.from_numpy(pose_value_lut()).to(dev)
        T = max_frames or cfg.cond_frame
        M = T * SEQ_LEN
        self.T_max = T
        self.x = torch.empty(M, C, dtype=torch.float32, device=dev)
        self.a_h = torch.empty(M, C, dtype=torch.float16, device=dev)
        self.y_h = torch.empty(M, C, dtype=torch.float16, device=dev)
        self.qkv_h = torch.empty(M, 3 * C, dtype=torch.float16, device=dev)
        self.h_h = torch.empty(M, 4 * C, dtype=torch.float16, device=dev)
        self.use_graphs = True
        self.parallel_suffix = True
        self._sbufs: list = []
        self._alloc_scene_state()

    def _alloc_scene_state(self):
        """Everything that belongs to ONE scene's rollout: map features, last-frame outputs, the look-ahead caches, CUDA graphs and their static
        token buffers, the ego decoder's scratch and outputs.  (Weights and the big working buffers x / a_h / y_h / qkv_h / h_h are shared by the
        scenes of a device, see for_scene.)"""
        dev, T = self.dev, self.T_max
        self.mf = [torch.empty(T * 1024, C, dtype=torch.float32, device=dev) for _ in range(2)]     # map feature without / with grid pos
        self.mw = [torch.empty(T * 1024, C, dtype=torch.float32, device=dev) for _ in range(2)]     # their warps
        self.f_last = {k: torch.empty(SEQ_LEN, C, dtype=torch.float32, device=dev) for k in ("ego", "map", "box", "all")}
        self.tar_feat = torch.empty(SEQ_LEN, C, dtype=torch.float32, device=dev)
        self.tcache: Dict[str, list] = {}      # temporal qkv caches of the look-ahead schedule (run_stack)
        self._graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        self._gtok: Dict[tuple, Dict[str, torch.Tensor]] = {}
        # ego decoder scratch
        self.q3 = torch.empty(3, C, dtype=torch.float32, device=dev)
        self.q3_h = torch.empty(3, C, dtype=torch.float16, device=dev)
        self.q3_qkv = torch.empty(3, 3 * C, dtype=torch.float16, device=dev)
        self.q3_y = torch.empty(3, C, dtype=torch.float16, device=dev)
        self.q3_q = torch.empty(3, C, dtype=torch.float16, device=dev)
        self.q3_hid = torch.empty(3, 4 * C, dtype=torch.float16, device=dev)
        self.scene_h = torch.empty(SEQ_LEN, C, dtype=torch.float16, device=dev)
        self.scene_k = torch.empty(SEQ_LEN, C, dtype=torch.float16, device=dev)
        self.scene_v = torch.empty(SEQ_LEN, C, dtype=torch.float16, device=dev)
        self.ego_logits = torch.empty(3, 1024, dtype=torch.float32, device=dev)
        self.ego_tok = torch.zeros(3, dtype=torch.int32, device=dev)

    def for_scene(self) -> "TarEncoders":
        """Encoders for one more scene on the same device (several scenes per GPU): shares the weights, the tables and the
        working buffers (the scenes' passes are stream-ordered one after the other), owns its scene state (_alloc_scene_state)."""
        t = TarEncoders.__new__(TarEncoders)
        t.__dict__.update(self.__dict__)
        t._alloc_scene_state()
        return t

    # ---- BlockTAR.forward_func (module.py:332-359) on self.x viewed as [T, S, 768] ---------------------------
    # cache (optional): the fused qkv activation [T_max * S, 2304] of this block's temporal sub-block, kept between calls.
    #   first = 0:  all T frames are computed (a whole window, or the first T frames of the NEXT window while the decode kernel runs)
    #   first = t:  only frame t is computed (self.x holds its S rows); its temporal attention reads the keys / values of frames 0..t-1
    #               from `cache` -- causal attention over frames makes those independent of frame t (module.py:342-345)
    def run_block(self, blk, T: int, S: int, cache: Optional[torch.Tensor] = None, first: int = 0, bufs=None):
        n = T - first if first else T          # frames computed now
        M = n * S
        B = bufs if bufs is not None else self
        x, a_h, y_h, qkv_h, h_h = B.x[:M], B.a_h[:M], B.y_h[:M], B.qkv_h[:M], B.h_h[:M]
        for sub in blk:
            ops.layernorm(x, sub["ln_a"], a_h)
            if sub["kind"] == "spatial":
                ops.gemm(a_h, sub["w_qkv"], sub["b_qkv"], qkv_h, ops.EPI_BIAS_F16)
                ops.spatial_attention(qkv_h, y_h, n, S)
                y_in = y_h
            else:       # causal over frames for every sequence position ("(b t) s c -> (b s) t c", module.py:342)
                full = cache[: T * S] if cache is not None else B.qkv_h[: T * S]
                ops.gemm(a_h, sub["w_qkv"], sub["b_qkv"], full[first * S: T * S], ops.EPI_BIAS_F16)
                y_full = B.y_h[: T * S]
                ops.small_attention(full, y_full, S, T, 1, S, True, q0=first)
                y_in = y_full[first * S: T * S]
            ops.gemm(y_in, sub["w_proj"], sub["b_proj"], x, ops.EPI_RESID_F32)
            ops.layernorm(x, sub["ln_b"], a_h)
            ops.gemm(a_h, sub["w_fc"], None, h_h, ops.EPI_GELU_F16)
            ops.gemm(h_h, sub["w_proj2"], None, x, ops.EPI_RESID_F32)

    def _caches(self, name: str, S: int):
        """Per-layer qkv caches of stack `name` (allocated on first use: 20 x S x 2304 fp16 per layer, 15.8 GB for UMGen_Large)."""
        c = self.tcache.get(name)
        if c is None:
            c = [torch.empty(self.T_max * S, 3 * C, dtype=torch.float16, device=self.dev) for _ in self.stacks[name]]
            self.tcache[name] = c
        return c

    def run_stack(self, name: str, ln: str, T: int, S: int, out_key: str, mode: str = "full", bufs=None) -> Optional[torch.Tensor]:
        """Runs stack `name` on self.x and leaves LayerNorm of the LAST frame in f_last[out_key]
        (only [:, -1] of every TAR output is consumed downstream, UMGen.py:1002,1228-1230).
        mode "full": self.x holds all T frames.  "prefix": self.x holds the first T frames of a longer window, the temporal qkv of every
        layer is kept, nothing is returned.  "suffix": self.x holds frame T-1 only, frames 0..T-2 come from the caches of a prefix run."""
        if mode == "full":
            for blk in self.stacks[name]:
                self.run_block(blk, T, S)
            last = self.x[(T - 1) * S: T * S]
        else:
            caches = self._caches(name, S)
            for blk, cache in zip(self.stacks[name], caches):
                self.run_block(blk, T, S, cache, first=(T - 1 if mode == "suffix" else 0), bufs=bufs)
            if mode == "prefix":
                return None
            last = (bufs if bufs is not None else self).x[:S]
        out = self.f_last[out_key][:S]
        ops.layernorm(last, self.ln[ln], out)
        return out

    def _embed(self, tok: Dict[str, torch.Tensor], n_mods: int, mf: torch.Tensor, mw: Optional[torch.Tensor], first: int = 0, bufs=None):
        """Embeds frames first..T-1 of `tok` into x[: (T - first) * S] (first = 0: the whole window)."""
        T = tok["pose"].shape[0]
        S = TASK_S[n_mods]
        if first:
            sl = {m: v[first:] for m, v in tok.items()}
            ops.embed_sequence(sl, self.tables, mf[first * 1024:], None if mw is None else mw[first * 1024:],
                               (bufs if bufs is not None else self).x[: (T - first) * S], n_mods, t_offset=first)
        else:
            ops.embed_sequence(tok, self.tables, mf, mw, self.x[: T * S], n_mods)
        return T, S

    @staticmethod
    def to_device_tokens(cond: Mapping[str, torch.Tensor], dev) -> Dict[str, torch.Tensor]:
        return {m: cond[m].to(device=dev, dtype=torch.int32).contiguous() for m in ("pose", "map", "bbox3d", "image")}

    # ---- infer_ego_net (UMGen.py:994-1005) ------------------------------------------------------------------
    def ego_action(self, tok: Dict[str, torch.Tensor], sample: SampleConfig, frame_index: int, mode: str = "full") -> torch.Tensor:
        """tok: device int32 tokens of the conditioning window (pose NOT yet shifted).  Returns [3] int32.
        mode "suffix": frames 0..T-2 were run by ego_prefix (look-ahead schedule), only the last frame is computed."""
        T = tok["pose"].shape[0]
        if mode == "suffix" and self.use_graphs:
            self._replay("ego", T, tok, lambda st: self._ego_logits(st, "suffix"))
        else:
            self._ego_logits(tok, mode)
        # token_sampler(logits, sample_param) (UMGen.py:1001-1004): sample_top_p(p) under sample_method "topp", topk(top_k) otherwise
        if sample.method == "topp":
            ops.sample_rows(self.ego_logits, 1, sample.temp, sample.seed, frame_index, self.ego_tok, top_p=sample.p)
        elif sample.method == "topk":
            ops.sample_rows(self.ego_logits, sample.top_k, sample.temp, sample.seed, frame_index, self.ego_tok)
        else:
            raise capi.UmgenError(f"unknown sample_method {sample.method!r}")
        return self.ego_tok

    def _ego_logits(self, tok: Dict[str, torch.Tensor], mode: str):
        T = tok["pose"].shape[0]
        first = T - 1 if mode == "suffix" else 0
        ops.map_feature(tok["map"][first:].reshape(-1), self.map_table, None, self.mf[0][first * 1024: T * 1024])
        _, S = self._embed(tok, 4, self.mf[0], None, first=first)
        scene = self.run_stack("ego_tar", "ln_ego_tar", T, S, "ego", mode)     # [2207, 768] fp32, last frame
        # ego queries of the last frame: egoe + spe[:3] + tpe[T-1] (UMGen.py:672-677, 503-510)
        self.q3.copy_(self.egoe + self.tables["spe_f"][:3] + self.tables["tpe_f"][T - 1][None])
        q3 = self.q3
        for d in self.ego_dec:
            ops.layernorm(q3, d["ln_1"], self.q3_h)
            ops.gemm(self.q3_h, d["w_qkv"], d["b_qkv"], self.q3_qkv, ops.EPI_BIAS_F16)
            ops.small_attention(self.q3_qkv, self.q3_y, 1, 3, 0, 1, False)
            ops.gemm(self.q3_y, d["w_proj"], d["b_proj"], q3, ops.EPI_RESID_F32)
            ops.layernorm(q3, d["ln_2"], self.q3_h)
            ops.layernorm(scene, d["ln_3"], self.scene_h)
            ops.gemm(self.q3_h, d["w_q_attn"], d["b_q_attn"], self.q3_q, ops.EPI_BIAS_F16)
            ops.gemm(self.scene_h, d["w_k_attn"], d["b_k_attn"], self.scene_k, ops.EPI_BIAS_F16)
            ops.gemm(self.scene_h, d["w_v_attn"], d["b_v_attn"], self.scene_v, ops.EPI_BIAS_F16)
            ops.cross_attention(self.q3_q, self.scene_k, self.scene_v, self.q3_y)
            ops.gemm(self.q3_y, d["w_c_proj"], d["b_c_proj"], q3, ops.EPI_RESID_F32)
            ops.layernorm(q3, d["ln_4"], self.q3_h)
            ops.gemm(self.q3_h, d["w_fc"], None, self.q3_hid, ops.EPI_GELU_F16)
            ops.gemm(self.q3_hid, d["w_proj2"], None, q3, ops.EPI_RESID_F32)
        ops.layernorm(q3, self.ln["ln_ego"], self.q3_h)
        ops.gemm(self.q3_h, self.head_ego, None, self.ego_logits, ops.EPI_STORE_F32)

    # The last-frame passes of the look-ahead schedule are ~2100 launches of kernels that run for a few microseconds each: replayed from a
    # CUDA graph (captured once per window length, every buffer is static; the tokens are copied into static buffers first).
    def _replay(self, kind: str, T: int, tok: Dict[str, torch.Tensor], body):
        key = (kind, T)
        st = self._gtok.get(key)
        if st is None:
            st = {m: torch.zeros_like(tok[m]) for m in ("pose", "map", "bbox3d", "image")}
            self._gtok[key] = st
        for m, v in st.items():
            v.copy_(tok[m])
        g = self._graphs.get(key)
        if g is None:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body(st)
            self._graphs[key] = g
        g.replay()

    def ego_prefix(self, tok: Dict[str, torch.Tensor]):
        """Look-ahead: the ego stack over the first T frames of the next window (pose NOT shifted); keeps the temporal qkv of every layer."""
        T = tok["pose"].shape[0]
        ops.map_feature(tok["map"].reshape(-1), self.map_table, None, self.mf[0][: T * 1024])
        _, S = self._embed(tok, 4, self.mf[0], None)
        self.run_stack("ego_tar", "ln_ego_tar", T, S, "ego", "prefix")

    # ---- cascade of _inference step 2 (UMGen.py:1482-1511) --------------------------------------------------
    BOX_ROW0, BOX_ROW1 = 1031, 1693        # rows of tar_feat that come from the box_tar pass (UMGEN_TAR_LATE_ROW0)

    def conditioning_feature(self, tok: Dict[str, torch.Tensor]) -> torch.Tensor:
        """tok: device int32 tokens with the pose stream already shifted.  Returns tar_feat [2207, 768] fp32."""
        self.conditioning_early(tok)
        self.conditioning_late(tok)
        return self.tar_feat

    def conditioning_early(self, tok: Dict[str, torch.Tensor]) -> torch.Tensor:
        """The full and the map pass and every row of tar_feat except the bbox3d block: all the OAR decode needs for its first 1030 steps.
        (The three passes are independent of each other, UMGen.py:1482-1495: only their order of execution changes.)"""
        T = tok["pose"].shape[0]
        n = T * 1024
        mf0, mf1, mw0, mw1 = self.mf[0][:n], self.mf[1][:n], self.mw[0][:n], self.mw[1][:n]
        ops.map_feature(tok["map"].view(-1), self.map_table, None, mf0)
        ops.map_feature(tok["map"].view(-1), self.map_table, self.grid_pos, mf1)
        ops.map_warp(mf0.view(T, 1024, C), tok["pose"], self.pose_lut, mw0.view(T, 1024, C))
        ops.map_warp(mf1.view(T, 1024, C), tok["pose"], self.pose_lut, mw1.view(T, 1024, C))
        Tt, S = self._embed(tok, 4, mf1, mw1)
        self.run_stack("TAR", "ln_tar", Tt, S, "all")
        Tt, S = self._embed(tok, 2, mf0, mw0)
        self.run_stack("map_tar", "ln_map_tar", Tt, S, "map")
        last = mw0[(T - 1) * 1024:]
        ops.assemble_tar_feat(self.f_last["all"], self.f_last["map"], self.f_last["box"], last, self.tar_feat, 0, self.BOX_ROW0)
        ops.assemble_tar_feat(self.f_last["all"], self.f_last["map"], self.f_last["box"], last, self.tar_feat, self.BOX_ROW1, SEQ_LEN)
        return self.tar_feat

    def conditioning_late(self, tok: Dict[str, torch.Tensor]) -> torch.Tensor:
        """The box pass and the bbox3d rows of tar_feat (needs conditioning_early's map features)."""
        T = tok["pose"].shape[0]
        n = T * 1024
        mf0, mw0 = self.mf[0][:n], self.mw[0][:n]
        Tt, S = self._embed(tok, 3, mf0, mw0)
        self.run_stack("box_tar", "ln_box_tar", Tt, S, "box")
        ops.assemble_tar_feat(self.f_last["all"], self.f_last["map"], self.f_last["box"], mw0[(T - 1) * 1024:], self.tar_feat,
                              self.BOX_ROW0, self.BOX_ROW1)
        return self.tar_feat

    # ---- look-ahead schedule: frames 0..T-2 of a window do not depend on its last frame (causal temporal attention, per-frame spatial
    # attention), so they are computed while the decode kernel of the previous frame runs (conditioning_prefix), and only the last frame
    # is computed on the critical path (conditioning_suffix).  Same kernels, same arithmetic per row: bit-identical features.
    def _map_features(self, tok: Dict[str, torch.Tensor], first: int):
        T = tok["pose"].shape[0]
        a, b = first * 1024, T * 1024
        mf0, mf1, mw0, mw1 = self.mf[0][a:b], self.mf[1][a:b], self.mw[0][a:b], self.mw[1][a:b]
        mtok = tok["map"][first:].reshape(-1)
        ops.map_feature(mtok, self.map_table, None, mf0)
        ops.map_feature(mtok, self.map_table, self.grid_pos, mf1)
        ops.map_warp(mf0.view(T - first, 1024, C), tok["pose"][first:], self.pose_lut, mw0.view(T - first, 1024, C))
        ops.map_warp(mf1.view(T - first, 1024, C), tok["pose"][first:], self.pose_lut, mw1.view(T - first, 1024, C))

    def conditioning_prefix(self, tok: Dict[str, torch.Tensor]):
        """tok: the first T frames of the NEXT window, pose stream already shifted.  Runs the three passes over them and keeps the temporal qkv."""
        T = tok["pose"].shape[0]
        n = T * 1024
        self._map_features(tok, 0)
        _, S = self._embed(tok, 4, self.mf[1][:n], self.mw[1][:n])
        self.run_stack("TAR", "ln_tar", T, S, "all", "prefix")
        _, S = self._embed(tok, 2, self.mf[0][:n], self.mw[0][:n])
        self.run_stack("map_tar", "ln_map_tar", T, S, "map", "prefix")
        _, S = self._embed(tok, 3, self.mf[0][:n], self.mw[0][:n])
        self.run_stack("box_tar", "ln_box_tar", T, S, "box", "prefix")

    def conditioning_suffix(self, tok: Dict[str, torch.Tensor]) -> torch.Tensor:
        """tok: the whole window (pose shifted) whose first T-1 frames went through conditioning_prefix.  Returns tar_feat [2207, 768]."""
        if self.use_graphs:
            self._replay("cond", tok["pose"].shape[0], tok, self._conditioning_suffix)
        else:
            self._conditioning_suffix(tok)
        return self.tar_feat

    def _suffix_bufs(self, k: int):
        """Working buffers of one last-frame pass (the three passes run on three streams): one frame of rows, except y (the temporal attention
        writes the last frame's rows at their place in the window)."""
        while len(self._sbufs) <= k:
            dev = self.dev
            b = type("Bufs", (), {})()
            b.x = torch.empty(SEQ_LEN, C, dtype=torch.float32, device=dev)
            b.a_h = torch.empty(SEQ_LEN, C, dtype=torch.float16, device=dev)
            b.y_h = torch.empty(self.T_max * SEQ_LEN, C, dtype=torch.float16, device=dev)
            b.qkv_h = torch.empty(SEQ_LEN, 3 * C, dtype=torch.float16, device=dev)
            b.h_h = torch.empty(SEQ_LEN, 4 * C, dtype=torch.float16, device=dev)
            b.stream = torch.cuda.Stream(device=dev)
            self._sbufs.append(b)
        return self._sbufs[k]

    def _conditioning_suffix(self, tok: Dict[str, torch.Tensor]):
        T = tok["pose"].shape[0]
        n = T * 1024
        self._map_features(tok, T - 1)
        passes = ((4, 1, "TAR", "ln_tar", "all"), (2, 0, "map_tar", "ln_map_tar", "map"), (3, 0, "box_tar", "ln_box_tar", "box"))
        if not self.parallel_suffix:
            for n_mods, w, name, ln, key in passes:
                _, S = self._embed(tok, n_mods, self.mf[w][:n], self.mw[w][:n], first=T - 1)
                self.run_stack(name, ln, T, S, key, "suffix")
        else:
            # one frame of rows fills a third of the GPU (a c_proj GEMM is 54 tiles for 148 SMs): the three passes are independent, so each runs
            # on its own stream with its own working buffers (inside the CUDA graph when one is being captured)
            main = torch.cuda.current_stream(self.dev)
            fork = torch.cuda.Event()
            fork.record(main)
            joins = []
            for k, (n_mods, w, name, ln, key) in enumerate(passes):
                b = self._suffix_bufs(k)
                b.stream.wait_event(fork)
                with torch.cuda.stream(b.stream):
                    _, S = self._embed(tok, n_mods, self.mf[w][:n], self.mw[w][:n], first=T - 1, bufs=b)
                    self.run_stack(name, ln, T, S, key, "suffix", bufs=b)
                    j = torch.cuda.Event()
                    j.record(b.stream)
                joins.append(j)
            for j in joins:
                main.wait_event(j)
        ops.assemble_tar_feat(self.f_last["all"], self.f_last["map"], self.f_last["box"], self.mw[0][(T - 1) * 1024: n], self.tar_feat)
