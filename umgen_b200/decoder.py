"""Host side of the OAR frame decoder: owns the device buffers (weights, KV cache, scratch) as torch
tensors and calls ``umgen_decode_frame`` through the C ABI.

Mirrors the argument meaning of the reference's ``UMGen.infer_oar_net`` (models/UMGen.py:1151-1273):
conditioning feature of the last frame, the frame's pose tokens, the previous frame's bbox3d tokens,
the controlled agent slots; returns the per-modality token ids."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Iterable, List, Mapping, Optional, Sequence

import torch

from . import capi
from .config import CONTENT_LEN, MOD_OFFSET, MODS, SEQ_LEN, ModelConfig, SampleConfig
from .weights import pack_oar

KV_ROWS = 2304


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


@dataclass
class DecodeResult:
    tokens: torch.Tensor            # [2207] int32 (device): ids of the frame incl. bos/eos aux ids
    picks: torch.Tensor             # [2207] int32: the sampler's own choice per position
    status: torch.Tensor            # [8] int32
    logits: Optional[torch.Tensor]  # [2207, 8192] fp32 when requested

    def by_modality(self) -> Dict[str, torch.Tensor]:
        return {m: self.tokens[MOD_OFFSET[m] + 1: MOD_OFFSET[m] + 1 + CONTENT_LEN[m]] for m in MODS}


class FrameDecoder:
    def __init__(self, state_dict: Mapping[str, torch.Tensor], cfg: ModelConfig, device="cuda:0",
                 packed: Optional[Dict[str, torch.Tensor]] = None, use_cluster: Optional[bool] = None):
        if not torch.cuda.is_available():
            raise capi.UmgenError("umgen_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = capi.lib()
        self.cfg = cfg
        self.dev = torch.device(device)
        self.w = packed if packed is not None else pack_oar(state_dict, cfg, self.dev)
        L = cfg.n_oar_layer
        self.kv = torch.zeros(L, 2, 16, KV_ROWS, 48, dtype=torch.float16, device=self.dev)
        self.scratch = torch.zeros(int(self.lib.umgen_decode_scratch_floats()), dtype=torch.float32, device=self.dev)
        self.tar_bbox_logits = torch.zeros(660, 1028, dtype=torch.float32, device=self.dev)
        self.out_tokens = torch.zeros(SEQ_LEN, dtype=torch.int32, device=self.dev)
        self.picks = torch.zeros(SEQ_LEN, dtype=torch.int32, device=self.dev)
        self.status = torch.zeros(96, dtype=torch.int32, device=self.dev)
        # kernel choice: 0 = 8-cluster kernel when its 8 clusters fit, else the L2-exchange kernel; 1 / 2 force the L2-exchange / 8-cluster kernel
        self.mode = 0
        self.grid = 0
        with torch.cuda.device(self.dev):
            self.cluster_capacity = int(self.lib.umgen_decode_cluster_capacity())
        self.use_cluster = False
        if use_cluster is None:
            use_cluster = self.cluster_capacity >= 8
        if use_cluster:
            self.pack_cluster()
        capi.preload(self.dev)
        self.debug = None          # optional [grid,16] int64 tensor for the timeline probe

    def pack_cluster(self):
        """Pack the OAR matrices in the cluster kernel's fragment order (include/umgen.h: umgen_pack_oar_cluster)."""
        if self.cluster_capacity < 8:
            raise capi.UmgenError(f"device holds {self.cluster_capacity} clusters of 8 CTAs; 8 needed")
        with torch.cuda.device(self.dev):
            if "oar_cl_h" not in self.w:
                self.w["oar_cl_h"] = torch.empty_like(self.w["oar_h"])
            capi.check(self.lib.umgen_pack_oar_cluster(self.w["oar_h"].data_ptr(), self.w["oar_cl_h"].data_ptr(), self.cfg.n_oar_layer,
                                                       torch.cuda.current_stream(self.dev).cuda_stream), "umgen_pack_oar_cluster")
        self.use_cluster = True

    @property
    def kernel_name(self) -> str:
        """The kernel umgen_decode_frame picks for the current mode."""
        if self.mode == 2 or (self.mode == 0 and self.use_cluster):
            return "decode_cluster_kernel"
        return "decode_frame_kernel"

    def tar_head_logits(self, tar_feat: torch.Tensor):
        """head_tar_bbox3d over the bbox3d rows of tar_feat (UMGen.py:1087,1103) on the current stream."""
        capi.check(self.lib.umgen_tar_bbox_logits(tar_feat.data_ptr(), self.w["head_tar_bbox_h"].data_ptr(), self.tar_bbox_logits.data_ptr(),
                                                  torch.cuda.current_stream(self.dev).cuda_stream), "umgen_tar_bbox_logits")

    def signal_ready(self, flag: torch.Tensor, value: int):
        capi.check(self.lib.umgen_signal_ready(flag.data_ptr(), int(value), torch.cuda.current_stream(self.dev).cuda_stream), "umgen_signal_ready")

    def for_scene(self) -> "FrameDecoder":
        """A decoder for one more scene on the same device: shares the packed weights and the scratch buffer, owns its KV cache, TAR-head
        logits, outputs and status (what differs per scene in umgen_decode_frames, include/umgen.h)."""
        d = FrameDecoder.__new__(FrameDecoder)
        d.__dict__.update(self.__dict__)
        L = self.cfg.n_oar_layer
        d.kv = torch.zeros(L, 2, 16, KV_ROWS, 48, dtype=torch.float16, device=self.dev)
        d.tar_bbox_logits = torch.zeros_like(self.tar_bbox_logits)
        d.out_tokens = torch.zeros_like(self.out_tokens)
        d.picks = torch.zeros_like(self.picks)
        d.status = torch.zeros_like(self.status)
        return d

    def _args(self, tar_feat, pose_tok, prev_bbox, sample, frame_index, control_slots, teacher, want_logits, n_steps, tar_ready, prefix_len):
        """Fills UmgenDecodeArgs for one frame of this decoder's scene.  Returns (args, logits tensor or None, tensors to keep alive)."""
        dev = self.dev
        if sample.method not in ("topk", "topp"):
            raise capi.UmgenError(f"unknown sample_method {sample.method!r}")
        tar_feat = tar_feat.to(device=dev, dtype=torch.float32).contiguous()
        assert tar_feat.shape == (SEQ_LEN, 768)
        pose_i = pose_tok.to(device=dev, dtype=torch.int32).contiguous().view(3)
        prev_i = prev_bbox.to(device=dev, dtype=torch.int32).contiguous().view(660)
        teach_i = None if teacher is None else teacher.to(device=dev, dtype=torch.int32).contiguous().view(SEQ_LEN)
        logits = torch.zeros(SEQ_LEN, 8192, dtype=torch.float32, device=dev) if want_logits else None
        mask = 0
        for s in (control_slots or ()):
            mask |= 1 << int(s)
        w = self.w
        if tar_ready is None:
            self.tar_head_logits(tar_feat)
        a = capi.UmgenDecodeArgs()
        a.n_layer = self.cfg.n_oar_layer
        for k in ("oar_h", "oar_f", "ln_oar_f", "head_map_h", "head_bbox_h", "head_img_h", "map_table_f", "img_table_f",
                  "be_f", "axe_f", "tske_f", "fpe_f", "box_lut_d"):
            setattr(a, k, w[k].data_ptr())
        a.tar_feat_f = tar_feat.data_ptr()
        a.tar_bbox_logits_f = self.tar_bbox_logits.data_ptr()
        a.pose_tok_i32 = pose_i.data_ptr()
        a.prev_bbox_i32 = prev_i.data_ptr()
        a.teacher_i32 = _ptr(teach_i)
        a.prefix_len = int(prefix_len)
        a.control_mask = mask
        if sample.method == "topp":
            a.top_k_map = a.top_k_bbox = a.top_k_img = 1          # ignored by the kernel in top-p mode
        else:
            a.top_k_map, a.top_k_bbox, a.top_k_img = int(sample.top_k_map), int(sample.top_k), int(sample.top_k_image)
        a.sample_topp = int(sample.method == "topp")
        # reference quirk (UMGen.py:1133): in topp mode the image branch passes topk_image (16) as p -> nothing is cut
        a.top_p_map, a.top_p_bbox, a.top_p_img = float(sample.p_map), float(sample.p), float(sample.top_k_image)
        a.temperature = float(sample.temp)
        a.seed = int(sample.seed)
        a.frame_index = int(frame_index)
        a.merge_ar_tar = int(self.cfg.merage_ar_tar)
        a.rule_constrain = int(self.cfg.rule_constrain)
        a.kv_h = self.kv.data_ptr()
        a.scratch_f = self.scratch.data_ptr()
        a.out_tokens_i32 = self.out_tokens.data_ptr()
        a.picks_i32 = self.picks.data_ptr()
        a.logits_dump_f = _ptr(logits)
        a.status_i32 = self.status.data_ptr()
        a.n_steps = int(n_steps)
        a.mode = int(self.mode)
        a.grid = int(self.grid)
        a.debug_u64 = _ptr(self.debug)
        a.tar_ready_i32 = None if tar_ready is None else tar_ready[0].data_ptr()
        a.tar_ready_value = 0 if tar_ready is None else int(tar_ready[1])
        a.oar_cl_h = _ptr(w.get("oar_cl_h")) if self.use_cluster else None
        return a, logits, (tar_feat, pose_i, prev_i, teach_i)

    def _check(self):
        st = self.status.cpu()
        if int(st[0]) != 0:
            raise capi.UmgenError(f"decode kernel aborted with code {int(st[0])} (a cross-CTA wait timed out)")

    def decode(self, tar_feat: torch.Tensor, pose_tok: torch.Tensor, prev_bbox: torch.Tensor,
               sample: SampleConfig, frame_index: int = 0, control_slots: Optional[Iterable[int]] = None,
               teacher: Optional[torch.Tensor] = None, want_logits: bool = False, n_steps: int = SEQ_LEN - 1,
               check: bool = True, tar_ready=None, prefix_len: int = 0) -> DecodeResult:
        """One frame.  tar_feat [2207,768] fp32, pose_tok [3], prev_bbox [660] (device or host ints).
        teacher [2207]: ids forced into the stream after each pick; positions 1..prefix_len of it are a GIVEN prefix (init_tokens of
        UMGen.py:1184-1201: no head, no sampling, no rule check there).
        tar_ready = (flag int32 tensor, value): the bbox3d rows of tar_feat (>= 1031) and the TAR-head logits are still being produced on another
        stream; the caller computes them (tar_head_logits) and then calls signal_ready (8-cluster kernel only, include/umgen.h)."""
        a, logits, keep = self._args(tar_feat, pose_tok, prev_bbox, sample, frame_index, control_slots, teacher, want_logits, n_steps,
                                     tar_ready, prefix_len)
        capi.check(self.lib.umgen_decode_frame(C.byref(a), torch.cuda.current_stream(self.dev).cuda_stream), "umgen_decode_frame")
        self._keepalive = keep
        res = DecodeResult(self.out_tokens, self.picks, self.status, logits)
        if check:
            self._check()
        return res

    @staticmethod
    def decode_batch(decoders: Sequence["FrameDecoder"], frames: Sequence[Mapping], sample: SampleConfig, want_logits: bool = False,
                     n_steps: int = SEQ_LEN - 1, check: bool = True, prefix_len: int = 0) -> List[DecodeResult]:
        """One frame of several scenes in ONE launch (umgen_decode_frames, include/umgen.h): decoders[s] (scene s: the first decoder and its
        for_scene() siblings) decodes frames[s] = {tar_feat, pose_tok, prev_bbox, frame_index, control_slots, teacher, tar_ready, seed}.  The scenes
        run in lockstep and share every weight byte read from HBM; the ids of each scene equal those of decode() on the same inputs."""
        d0 = decoders[0]
        n = len(decoders)
        if n != len(frames) or n < 1:
            raise capi.UmgenError("decode_batch needs one frame per decoder")
        if n > int(d0.lib.umgen_decode_max_scenes()):
            raise capi.UmgenError(f"{n} scenes per launch; this build takes at most {int(d0.lib.umgen_decode_max_scenes())}")
        arr = (capi.UmgenDecodeArgs * n)()
        out, keep = [], []
        for s, (d, f) in enumerate(zip(decoders, frames)):
            if d.w is not d0.w or d.scratch is not d0.scratch:
                raise capi.UmgenError("the decoders of a batch must share weights and scratch (FrameDecoder.for_scene)")
            a, logits, k = d._args(f["tar_feat"], f["pose_tok"], f["prev_bbox"], sample, f.get("frame_index", 0), f.get("control_slots"),
                                   f.get("teacher"), want_logits, n_steps, f.get("tar_ready"), prefix_len)
            if f.get("seed") is not None:
                a.seed = int(f["seed"])
            arr[s] = a
            keep.append(k)
            out.append(DecodeResult(d.out_tokens, d.picks, d.status, logits))
        capi.check(d0.lib.umgen_decode_frames(arr, n, torch.cuda.current_stream(d0.dev).cuda_stream), "umgen_decode_frames")
        d0._keepalive = keep
        if check:
            for d in decoders:
                d._check()
        return out
