"""Next-scene prediction engine: the B200-native body of ``UMGen.inference`` / ``UMGen._inference``
(reference models/UMGen.py:1406-1671).  Same arguments, same return value; the per-frame work runs in
hand-written sm_100a kernels behind the C ABI (TAR encoders: tar.py; OAR decode: decoder.py).

Schedule of a frame: the decode kernel runs on a second stream and leaves 84 SMs free; the TAR stacks run beside
it over the frames of the NEXT window that are already final (look-ahead), so that after a decode only the window's last frame goes
through the stacks.  Every schedule produces bit-identical tokens, logits and features (tests/test_engine_gpu.py)."""
from __future__ import annotations

import dataclasses
import os
from dataclasses import dataclass
from typing import Dict, List, Mapping, Optional, Sequence

import numpy as np
import torch

from . import capi
from .config import CONTENT_LEN, MOD_OFFSET, MODS, N_SLOTS, SEQ_LEN, ModelConfig, SampleConfig
from .decoder import FrameDecoder
from .tar import TarEncoders


@dataclass
class FrameTrace:
    """Device tensors kept for parity tests when ``engine.keep_trace`` is set."""
    ego_logits: Optional[torch.Tensor] = None
    tar_feat: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    tokens: Optional[torch.Tensor] = None       # [2207] output ids (slots wiped by the rule check read <pad>)
    picks: Optional[torch.Tensor] = None        # [2207] the decode stream: what was appended at each step
    status: Optional[List[int]] = None


class _Pending:
    """One frame of one scene between its conditioning passes and the end of its decode (UMGenEngine._begin_frame .. _end_frame)."""
    tr = cond = tok = pose_unshifted = pose_new = control_slots = teacher = feat = prev_bbox = res = None
    fidx = prefix_len = 0
    la_ok = suffix = False


def next_window_start(T: int, window: int) -> int:
    """A rollout keeps at most `window` conditioning frames (UMGen.py:1598-1601): the next window drops the first frame of the current one
    once it is `window` long.  Returns the index of the first current frame that is also in the next window."""
    return 1 if T >= window else 0


def window_continues(assumed: Optional[dict], cond: Mapping[str, torch.Tensor]) -> bool:
    """Is `cond` ([T, S_mod] tokens per modality, host or device) the window the look-ahead passes of the previous frame assumed?
    assumed = {"T": length, "host": {mod: LongTensor [T-1, S_mod]} (the frames known in advance), "pose_new": the ego action the previous
    frame was decoded with -- it must be the pose of the window's last frame}."""
    if assumed is None or assumed.get("host") is None or cond["pose"].shape[0] != assumed["T"]:
        return False
    P = assumed["T"] - 1
    for m in MODS:
        if not torch.equal(cond[m][:P].cpu().long(), assumed["host"][m]):
            return False
    return torch.equal(cond["pose"][P].cpu().long().view(3), assumed["pose_new"].cpu().long().view(3))


class UMGenEngine:
    def __init__(self, state_dict: Mapping[str, torch.Tensor], cfg: ModelConfig, sample: Optional[SampleConfig] = None,
                 device="cuda:0"):
        if not torch.cuda.is_available():
            raise capi.UmgenError("umgen_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.cfg = cfg
        self.sample = sample or SampleConfig()
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        self.tar = TarEncoders(state_dict, cfg, device)
        self.dec = FrameDecoder(state_dict, cfg, device)
        self.keep_trace = False
        self.want_logits = False
        self.trace: List[FrameTrace] = []
        self.frame_counter = 0
        self.check_status = True       # read back the decode kernel's abort word after every frame
        # Run the box_tar pass beside the decode kernel: the 8-cluster kernel holds 64 of the 148 SMs for ~1.3 s and needs the bbox3d rows of the
        # conditioning feature only from step 1030 on (include/umgen.h: tar_ready_i32).  Same arithmetic, different schedule.
        self.overlap = True
        self.dec_stream = torch.cuda.Stream(device=self.dev, priority=-1)
        self.ready_flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.n_sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
        # A kernel launched for the first time in a process is loaded lazily, and that load waits for the device to go idle -- which never happens
        # while the decode kernel spins on tar_ready.  So the first frame of an engine runs the sequential schedule (it launches every kernel of
        # the late path) and the signal kernel is launched once here.
        self._late_path_loaded = False
        self.dec.signal_ready(self.ready_flag, 0)
        # Look-ahead schedule (supersedes `overlap` when a frame continues the previous one): frames 0..T-2 of the NEXT window are final before
        # the current frame is decoded and -- causal temporal attention, per-frame spatial attention -- independent of the window's last frame,
        # so all four TAR stacks run over them beside the decode kernel (tar.ego_prefix / conditioning_prefix keep the temporal qkv of every
        # layer); after the decode only the last frame of the window goes through the stacks (1/20 of the work).
        self.lookahead = True
        # cap on the GEMM CTAs of the look-ahead passes (0 = every free SM).  The passes have the whole decode (~0.95 s) to finish; on fewer SMs they
        # disturb the decode kernel's L2 exchanges less.  Measured with the round-2 GEMM (decode kernel / passes beside it, ms): 64 CTAs 985 / 475,
        # 48 CTAs 958 / 529, 32 CTAs 954 / 684 (the decode kernel alone: 926)
        self.lookahead_sms = int(os.environ.get("UMGEN_LOOKAHEAD_SMS", "32"))
        self._la = None                # what the prefix run beside the last decode assumed about the next window
        self.time_lookahead = False    # record CUDA events around the look-ahead passes (bench.py): self.la_events = (start, passes done, decode done)
        self.la_events = None
        self.window = cfg.cond_frame   # frames a rollout keeps as conditioning (inference() sets it to its cond_frames argument)

    def for_scene(self, k: int) -> "UMGenEngine":
        """An engine for the k-th further scene on the same device (SceneBatchEngine): shares the weights, the working buffers and the decode
        stream; owns its scene state (look-ahead caches, KV cache, outputs, frame counter) and draws from its own random stream (seed + k)."""
        e = UMGenEngine.__new__(UMGenEngine)
        e.__dict__.update(self.__dict__)
        e.tar = self.tar.for_scene()
        e.dec = self.dec.for_scene()
        e.sample = dataclasses.replace(self.sample, seed=int(self.sample.seed) + int(k))
        e.trace = []
        e.frame_counter = 0
        e._la = None
        e.la_events = None
        e.ready_flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        return e

    # one new frame: _inference (UMGen.py:1406-1540).  cond: {mod: LongTensor [T, S_mod]} on any device.
    def frame(self, cond: Dict[str, torch.Tensor], init: Optional[Dict[str, Optional[torch.Tensor]]] = None,
              control_test: bool = False, teacher: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        with torch.cuda.device(self.dev):
            tok = TarEncoders.to_device_tokens(cond, self.dev)      # H2D of the conditioning window
            return self.frame_device(tok, cond, init, control_test, teacher, continues=self._continues(cond))

    def _continues(self, cond: Dict[str, torch.Tensor]) -> bool:
        """Does this window continue the previous frame the way the look-ahead run assumed?  (host compare, see window_continues)"""
        return window_continues(self._la, cond)

    def frame_device(self, tok: Dict[str, torch.Tensor], cond: Optional[Dict[str, torch.Tensor]] = None,
                     init: Optional[Dict[str, Optional[torch.Tensor]]] = None, control_test: bool = False,
                     teacher: Optional[torch.Tensor] = None, continues: bool = False) -> Dict[str, torch.Tensor]:
        with torch.cuda.device(self.dev):       # every launch below goes to this engine's device whatever the caller's current device is
            return self._frame_device(tok, cond, init, control_test, teacher, continues)

    def _given_prefix(self, init, control_test: bool, pose_new: torch.Tensor):
        """init modalities other than pose are a GIVEN prefix of the new frame (UMGen.py:1184-1201, after _inference drops bbox3d in control
        mode :1474 and everything but pose/map/bbox3d :1515-1523): returns (teacher [2207] int32 on the device, prefix_len) or (None, 0)."""
        if init is None:
            return None, 0
        given = [m for m in ("map", "bbox3d") if init.get(m) is not None and not (control_test and m == "bbox3d")]
        if not given:
            return None, 0
        if given == ["bbox3d"]:
            raise capi.UmgenError("init_tokens with bbox3d but without map is not a contiguous prefix of the frame (pose, map, bbox3d, image)")
        t = torch.zeros(SEQ_LEN, dtype=torch.int32, device=self.dev)
        t[1:4] = pose_new.to(torch.int32)
        end = 0
        for m in given:
            v = init[m].to(device=self.dev, dtype=torch.int32).view(-1)
            if v.numel() != CONTENT_LEN[m]:
                raise capi.UmgenError(f"init_tokens[{m!r}] has {v.numel()} tokens per frame, expected {CONTENT_LEN[m]}")
            t[MOD_OFFSET[m] + 1: MOD_OFFSET[m] + 1 + CONTENT_LEN[m]] = v
            end = MOD_OFFSET[m] + CONTENT_LEN[m] + 2          # 1-indexed position of the modality's eos
        return t, end

    def _frame_device(self, tok, cond, init, control_test, teacher, continues):
        """Same as frame() with the conditioning tokens already resident on the device (int32 [T, S_mod]).
        Returns device int64 tokens; no host synchronisation except the decode status check.
        continues: the caller asserts that this window is the previous one extended by the frame this engine returned last (sliding to
        cond_frame frames) -- the look-ahead schedule then computes only the last frame of the window (frame() checks this itself)."""
        p = self._begin_frame(tok, cond, init, control_test, teacher, continues)
        tok = p.tok
        if p.la_ok:
            self._decode_lookahead(p)
        elif self.overlap and self._late_path_loaded and self.dec.kernel_name == "decode_cluster_kernel":
            p.res, p.feat = self._frame_overlapped(tok, p.pose_new, p.fidx, p.control_slots, p.teacher, p.prefix_len)
        else:
            self._late_path_loaded = True
            # Step 2: TAR cascade -> conditioning feature of the last frame
            p.feat = self.tar.conditioning_feature(tok)
            # Step 3: OAR decode of the frame
            p.res = self.dec.decode(p.feat, p.pose_new, tok["bbox3d"][-1], self.sample, frame_index=p.fidx, control_slots=p.control_slots,
                                    teacher=p.teacher, want_logits=self.want_logits, check=self.check_status, prefix_len=p.prefix_len)
        return self._end_frame(p)

    def _begin_frame(self, tok, cond, init, control_test, teacher, continues) -> "_Pending":
        """Step 1 of _inference (ego action, UMGen.py:1440-1455) and what decides how steps 2 + 3 run; with the look-ahead schedule also step 2
        (the conditioning feature), so that the decode of this frame -- alone or together with other scenes' frames -- can be launched next."""
        dev = self.dev
        p = _Pending()
        p.tr = FrameTrace() if self.keep_trace else None
        tok = dict(tok)
        p.cond = cond
        p.fidx = self.frame_counter
        self.frame_counter += 1
        T = tok["pose"].shape[0]
        p.la_ok = self.lookahead and self.dec.kernel_name != "decode_frame_kernel"
        p.suffix = p.la_ok and continues and self._la is not None and self._la["T"] == T and T > 1
        p.pose_unshifted = tok["pose"]
        if init is not None and init.get("pose") is not None:
            p.pose_new = init["pose"].to(device=dev, dtype=torch.int32).view(3)
        else:
            p.pose_new = self.tar.ego_action(tok, self.sample, p.fidx, "suffix" if p.suffix else "full").clone()
            if p.tr is not None:
                p.tr.ego_logits = self.tar.ego_logits.clone()
        tok["pose"] = torch.cat([tok["pose"], p.pose_new[None]], dim=0)[1:].contiguous()
        # controlled agent slots (UMGen.py:1459-1475): overwrite the last conditioning frame in place
        p.control_slots = None
        if control_test and init is not None and init.get("bbox3d") is not None:
            ctrl = init["bbox3d"].view(-1)
            valid = ctrl != -1
            cond["bbox3d"][-1, valid.to(cond["bbox3d"].device)] = ctrl[valid].to(cond["bbox3d"])
            tok["bbox3d"] = cond["bbox3d"].to(device=dev, dtype=torch.int32).contiguous()
            p.control_slots = np.where(valid.view(N_SLOTS, -1).any(dim=1).cpu().numpy())[0].tolist()
        p.prefix_len = 0
        if teacher is None:
            teacher, p.prefix_len = self._given_prefix(init, control_test, p.pose_new)
        p.teacher = teacher
        p.tok = tok
        if p.la_ok:
            p.feat = self.tar.conditioning_suffix(tok) if p.suffix else self.tar.conditioning_feature(tok)
            p.prev_bbox = tok["bbox3d"][-1].contiguous()
        return p

    def _end_frame(self, p: "_Pending") -> Dict[str, torch.Tensor]:
        res = p.res
        ids = res.tokens.to(torch.int64)
        if p.tr is not None:
            tr = p.tr
            tr.tar_feat = p.feat.clone()
            tr.logits = res.logits
            tr.tokens = ids.clone()
            tr.picks = res.picks.to(torch.int64).clone()
            tr.status = res.status.cpu().tolist()
            self.trace.append(tr)
        return {m: ids[MOD_OFFSET[m] + 1: MOD_OFFSET[m] + 1 + CONTENT_LEN[m]] for m in MODS}

    def _decode_lookahead(self, p: "_Pending"):
        """Step 3 of _inference with the look-ahead schedule (see __init__): the decode kernel on a second stream while the first frames of the
        NEXT window go through the stacks on the SMs it leaves free (the conditioning feature was computed by _begin_frame: from the last frame
        only when the first T-1 frames of this window went through the stacks beside the previous decode)."""
        cur = torch.cuda.current_stream(self.dev)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.dec_stream.wait_event(ev)
        with torch.cuda.stream(self.dec_stream):
            p.res = self.dec.decode(p.feat, p.pose_new, p.prev_bbox, self.sample, frame_index=p.fidx, control_slots=p.control_slots,
                                    teacher=p.teacher, want_logits=self.want_logits, check=False, prefix_len=p.prefix_len)
            done = torch.cuda.Event()
            done.record(self.dec_stream)
        t_ev = None
        if self.time_lookahead:
            t_ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            t_ev[0].record(cur)
        self._lookahead_passes(p)
        if t_ev is not None:
            t_ev[1].record(cur)
        cur.wait_event(done)
        if t_ev is not None:
            t_ev[2].record(cur)
            self.la_events = t_ev
        if self.check_status:
            self.dec._check()

    def _lookahead_passes(self, p: "_Pending", sm_limit: Optional[int] = None):
        """The next window = this one (without its first frame once it is cond_frame long) + the frame being decoded: its first frames go
        through the four stacks now, beside the decode kernel."""
        tok = p.tok
        T = tok["pose"].shape[0]
        s = next_window_start(T, self.window)
        self._la = None
        if T - s >= 1 and T - s + 1 <= self.tar.T_max:
            nxt = {m: tok[m][s:].contiguous() for m in MODS}            # pose stream shifted: its last row is pose_new
            nxt_ego = dict(nxt)
            nxt_ego["pose"] = p.pose_unshifted[s:].contiguous()
            lib = capi.lib()
            free = self.n_sms - (64 if self.dec.kernel_name == "decode_cluster_kernel" else 16)
            lib.umgen_gemm_set_sm_limit(max(min(free, (self.lookahead_sms if sm_limit is None else sm_limit) or free), 1))
            try:
                self.tar.ego_prefix(nxt_ego)
                self.tar.conditioning_prefix(nxt)
            finally:
                lib.umgen_gemm_set_sm_limit(0)
            host = None
            if p.cond is not None:
                host = {m: p.cond[m][s:].clone().cpu().long() for m in MODS}
            self._la = {"T": T - s + 1, "host": host, "pose_new": p.pose_new}

    def _frame_overlapped(self, tok, pose_new, fidx, control_slots, teacher, prefix_len=0):
        """Steps 2 + 3 of _inference with the box_tar pass running beside the decode kernel (see __init__)."""
        cur = torch.cuda.current_stream(self.dev)
        seq = fidx + 1
        feat = self.tar.conditioning_early(tok)
        prev_bbox = tok["bbox3d"][-1].contiguous()
        ev = torch.cuda.Event()
        ev.record(cur)
        self.dec_stream.wait_event(ev)
        with torch.cuda.stream(self.dec_stream):
            res = self.dec.decode(feat, pose_new, prev_bbox, self.sample, frame_index=fidx, control_slots=control_slots, teacher=teacher,
                                  want_logits=self.want_logits, check=False, tar_ready=(self.ready_flag, seq), prefix_len=prefix_len)
            done = torch.cuda.Event()
            done.record(self.dec_stream)
        lib = capi.lib()
        lib.umgen_gemm_set_sm_limit(max(self.n_sms - 64, 1))      # the decode kernel's 64 CTAs each hold a whole SM
        try:
            self.tar.conditioning_late(tok)
            self.dec.tar_head_logits(feat)
            self.dec.signal_ready(self.ready_flag, seq)
        finally:
            lib.umgen_gemm_set_sm_limit(0)
        cur.wait_event(done)
        if self.check_status:
            st = res.status.cpu()
            if int(st[0]) != 0:
                raise capi.UmgenError(f"decode kernel aborted with code {int(st[0])} (a cross-CTA wait timed out)")
        return res, feat

    # UMGen.inference (UMGen.py:1542-1671): tokens carry the leading batch-1 axis; returns numpy int64
    def inference(self, new_frames: int, cond_frames: int = 1, input_cond_frames: int = -1, pred_task: str = "pose_map_bbox3d_image",
                  input_cond_tokens: Optional[Dict[str, torch.Tensor]] = None, init_tokens: Optional[Dict[str, torch.Tensor]] = None,
                  cond_on_tar: bool = False, test_map_affine: bool = False, max_objects=100, control_test: bool = False,
                  **kwargs) -> Dict[str, np.ndarray]:
        if pred_task != "pose_map_bbox3d_image":
            raise capi.UmgenError(f"pred_task {pred_task!r} is not supported (the evaluation config defines only pose_map_bbox3d_image)")
        new_frames, cond_frames, input_cond_frames = int(new_frames), int(cond_frames), int(input_cond_frames)      # the harness may hand over 1-element tensors (model_pl.py:166-169)
        if input_cond_frames == -1:
            input_cond_frames = cond_frames
        if cond_frames > self.tar.T_max:
            raise capi.UmgenError(f"cond_frames {cond_frames} exceeds the engine's window {self.tar.T_max}")
        self.window = cond_frames
        self._la = None
        self.frame_counter = 0          # the random stream of a rollout is a function of (seed, frame index, position, draw) only
        out = {m: input_cond_tokens[m][0, :input_cond_frames].clone().cpu().long() for m in MODS}
        cond = {m: input_cond_tokens[m][0, :input_cond_frames].clone().cpu().long() for m in MODS}
        for idx in range(new_frames):
            if cond["pose"].shape[0] > cond_frames:
                cond = {m: cond[m][-cond_frames:].clone() for m in MODS}
            init = None
            if init_tokens is not None:
                init = {m: (v[0, idx].cpu() if idx < v.shape[1] else None) for m, v in init_tokens.items()}
                if "pose" in init and init["pose"] is None:                        # UMGen.py:1613-1619: the control horizon is over
                    init_tokens, control_test, init = None, False, None
            new = self.frame(cond, init, control_test)
            for m in MODS:
                use_init = init_tokens is not None and m in init_tokens and not (control_test and m == "bbox3d")
                row = init[m].long().view(-1) if use_init else new[m].cpu()
                cond[m] = torch.cat([cond[m], row[None]], dim=0)
                out[m] = torch.cat([out[m], row[None]], dim=0)
        return {m: out[m][None].numpy() for m in MODS}


class SceneBatchEngine:
    """Several scenes per GPU.  The reference generates one scene at a time (UMGen.py:907,1093 are hard-wired to batch 1);
    at batch 1 the OAR decode is bound by the latency of its per-layer exchanges, not by HBM, and 79 % of its bytes are weights.  Here B scenes
    advance in lockstep: ONE launch of the 8-cluster kernel decodes the B frames (umgen_decode_frames: the scenes share every weight fragment and
    every exchange), while the look-ahead passes of all B next windows run beside it on the free SMs.  Scene k is bit-identical to a UMGenEngine
    run on the same inputs with sample.seed + k (tests/test_engine_gpu.py)."""

    def __init__(self, state_dict: Mapping[str, torch.Tensor], cfg: ModelConfig, sample: Optional[SampleConfig] = None, device="cuda:0",
                 scenes: int = 2):
        self._init(UMGenEngine(state_dict, cfg, sample, device), scenes)

    @classmethod
    def around(cls, engine: UMGenEngine, scenes: int) -> "SceneBatchEngine":
        """A batch engine whose first scene is an existing one-scene engine (weights, working buffers and streams are shared, not copied)."""
        self = cls.__new__(cls)
        self._init(engine, scenes)
        return self

    def _init(self, e0: UMGenEngine, scenes: int):
        cfg = e0.cfg
        if e0.dec.kernel_name != "decode_cluster_kernel":
            raise capi.UmgenError("several scenes per launch need the 8-cluster decode kernel (umgen_decode_cluster_capacity() >= 8)")
        max_scenes = int(capi.lib().umgen_decode_max_scenes())
        if not 1 <= scenes <= max_scenes:
            raise capi.UmgenError(f"scenes per GPU must be in [1, {max_scenes}] (got {scenes})")
        self.engines: List[UMGenEngine] = [e0] + [e0.for_scene(k) for k in range(1, scenes)]
        self.dev = e0.dev
        self.cfg = cfg
        # cap on the GEMM CTAs of the look-ahead passes, as for one scene: on all 84 free SMs the passes of 2 / 3 scenes finish sooner (0.93 / 1.38 s)
        # but slow the decode kernel's L2 exchanges down; at 48 they take 1.24 / 1.84 s, still inside the decode (1.29 / 1.90 s), and the step is
        # 5 % / 2.5 % shorter (measured: 3072 -> 3230 and 3198 -> 3279 tokens/s)
        self.lookahead_sms = int(os.environ.get("UMGEN_LOOKAHEAD_SMS_BATCH", "48"))
        self.check_status = True
        self.time_lookahead = False
        self.la_events = None

    @property
    def scenes(self) -> int:
        return len(self.engines)

    def frames(self, conds: Sequence[Dict[str, torch.Tensor]], inits: Optional[Sequence[Optional[Dict]]] = None,
               control_test: bool = False) -> List[Dict[str, torch.Tensor]]:
        """One new frame of every scene.  conds[k]: {mod: LongTensor [T, S_mod]} of scene k (any device)."""
        with torch.cuda.device(self.dev):
            toks = [TarEncoders.to_device_tokens(c, self.dev) for c in conds]
            cont = [e._continues(c) for e, c in zip(self.engines, conds)]
            return self.frames_device(toks, conds, inits, control_test, continues=cont)

    def frames_device(self, toks, conds=None, inits=None, control_test: bool = False, teachers=None, continues=None):
        E = self.engines
        n = len(E)
        if len(toks) != n:
            raise capi.UmgenError(f"{len(toks)} windows for {n} scenes")
        conds = conds if conds is not None else [None] * n
        inits = inits if inits is not None else [None] * n
        teachers = teachers if teachers is not None else [None] * n
        continues = continues if continues is not None else [False] * n
        with torch.cuda.device(self.dev):
            pend = [e._begin_frame(t, c, i, control_test, th, bool(k)) for e, t, c, i, th, k in zip(E, toks, conds, inits, teachers, continues)]
            if len({p.prefix_len for p in pend}) != 1:
                raise capi.UmgenError("the scenes of a launch must be given the same modalities (init_tokens): prefix lengths differ")
            e0 = E[0]
            cur = torch.cuda.current_stream(self.dev)
            ev = torch.cuda.Event()
            ev.record(cur)
            e0.dec_stream.wait_event(ev)
            with torch.cuda.stream(e0.dec_stream):
                fr = [dict(tar_feat=p.feat, pose_tok=p.pose_new, prev_bbox=p.prev_bbox, frame_index=p.fidx, control_slots=p.control_slots,
                           teacher=p.teacher, seed=e.sample.seed) for e, p in zip(E, pend)]
                res = FrameDecoder.decode_batch([e.dec for e in E], fr, e0.sample, want_logits=e0.want_logits, check=False,
                                                prefix_len=pend[0].prefix_len)
                done = torch.cuda.Event()
                done.record(e0.dec_stream)
            t_ev = None
            if self.time_lookahead:
                t_ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                t_ev[0].record(cur)
            for e, p, r in zip(E, pend, res):
                p.res = r
                e._lookahead_passes(p, sm_limit=self.lookahead_sms)
            if t_ev is not None:
                t_ev[1].record(cur)
            cur.wait_event(done)
            if t_ev is not None:
                t_ev[2].record(cur)
                self.la_events = t_ev
            if self.check_status:
                for e in E:
                    e.dec._check()
            return [e._end_frame(p) for e, p in zip(E, pend)]

    def inference(self, new_frames: int, cond_frames: int = 1, input_cond_frames: int = -1, pred_task: str = "pose_map_bbox3d_image",
                  input_cond_tokens: Optional[Dict[str, torch.Tensor]] = None, init_tokens: Optional[Dict[str, torch.Tensor]] = None,
                  cond_on_tar: bool = False, test_map_affine: bool = False, max_objects=100, control_test: bool = False,
                  **kwargs) -> Dict[str, np.ndarray]:
        """UMGen.inference (UMGen.py:1542-1671) for B scenes at once: tokens carry a leading axis of B = self.scenes; returns numpy int64
        [B, input_cond_frames + new_frames, S_mod].  Row k equals what UMGenEngine.inference returns for scene k alone (with sample.seed + k)."""
        B = self.scenes
        if pred_task != "pose_map_bbox3d_image":
            raise capi.UmgenError(f"pred_task {pred_task!r} is not supported (the evaluation config defines only pose_map_bbox3d_image)")
        if input_cond_tokens["pose"].shape[0] != B:
            raise capi.UmgenError(f"input_cond_tokens hold {input_cond_tokens['pose'].shape[0]} scenes, this engine decodes {B} per launch")
        new_frames, cond_frames, input_cond_frames = int(new_frames), int(cond_frames), int(input_cond_frames)      # the harness may hand over 1-element tensors (model_pl.py:166-169)
        if input_cond_frames == -1:
            input_cond_frames = cond_frames
        if cond_frames > self.engines[0].tar.T_max:
            raise capi.UmgenError(f"cond_frames {cond_frames} exceeds the engine's window {self.engines[0].tar.T_max}")
        for e in self.engines:
            e.window = cond_frames
            e._la = None
            e.frame_counter = 0
        out = [{m: input_cond_tokens[m][k, :input_cond_frames].clone().cpu().long() for m in MODS} for k in range(B)]
        cond = [{m: input_cond_tokens[m][k, :input_cond_frames].clone().cpu().long() for m in MODS} for k in range(B)]
        for idx in range(new_frames):
            inits = [None] * B
            for k in range(B):
                if cond[k]["pose"].shape[0] > cond_frames:
                    cond[k] = {m: cond[k][m][-cond_frames:].clone() for m in MODS}
            if init_tokens is not None:
                inits = [{m: (v[k, idx].cpu() if idx < v.shape[1] else None) for m, v in init_tokens.items()} for k in range(B)]
                if "pose" in inits[0] and inits[0]["pose"] is None:                # UMGen.py:1613-1619: the control horizon is over
                    init_tokens, control_test, inits = None, False, [None] * B
            new = self.frames(cond, inits, control_test)
            for k in range(B):
                for m in MODS:
                    use_init = init_tokens is not None and m in init_tokens and not (control_test and m == "bbox3d")
                    row = inits[k][m].long().view(-1) if use_init else new[k][m].cpu()
                    cond[k][m] = torch.cat([cond[k][m], row[None]], dim=0)
                    out[k][m] = torch.cat([out[k][m], row[None]], dim=0)
        return {m: np.stack([out[k][m].numpy() for k in range(B)]) for m in MODS}
