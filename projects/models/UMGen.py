"""Drop-in ``projects.models.UMGen.UMGen`` (reference projects/models/UMGen.py:51-270, 1542-1671).

Same registration, constructor argument (the config Namespace of configs/UMGen_config_evaluation.py:344-430),
parameter names / shapes / dtypes (so ``load_state_dict(ckpt["module"], strict=False)`` of tools/infer_fun.py:43-50
loads ``UMGen_Large.pt``) and ``inference()`` signature / return value.  The forward computation is not
re-stated in PyTorch: ``inference`` hands the module's ``state_dict`` to the B200 engine (``umgen_b200``), which
packs it once and runs the hand-written sm_100a kernels.  There is no CPU path: without CUDA ``inference`` raises."""
from __future__ import annotations

import dataclasses
import os
from typing import Dict, Optional

import numpy as np
import torch
from torch import nn

from projects.registry import MODELS
from umgen_b200 import synth
from umgen_b200.config import ModelConfig, SampleConfig

__all__ = ["UMGen"]


class _Tree(nn.Module):
    """Nested container that reproduces the reference's module hierarchy by name only."""

    def add(self, dotted: str, tensor: torch.Tensor, buffer: bool):
        head, _, rest = dotted.partition(".")
        if rest:
            if head not in self._modules:
                self.add_module(head, _Tree())
            self._modules[head].add(rest, tensor, buffer)
        elif buffer:
            self.register_buffer(head, tensor)
        else:
            self.register_parameter(head, nn.Parameter(tensor, requires_grad=tensor.is_floating_point() and tensor.dtype != torch.bfloat16))


@MODELS.register_module()
class UMGen(nn.Module):
    def __init__(self, config):
        super().__init__()
        g = lambda k, d=None: getattr(config, k, d)
        self.config = config
        # fields of UMGen.__init__ that the harness or callers read back (UMGen.py:56-173)
        self.task = config.task
        self.task_name_id = config.task_name_id
        self.token_len, self.seq_len, self.bos_eos = config.token_len, config.seq_len, config.bos_eos
        self.cond_frame, self.max_frame_len = config.cond_frame, config.max_frame_len
        self.sfmx_temp, self.top_k = config.sfmx_temp, config.top_k
        self.top_k_map = g("top_k_map", self.top_k)
        self.topk_image = 16                                        # hard-coded in the reference (UMGen.py:103)
        self.p, self.p_map = config.p, g("p_map", config.p)
        assert config.sample_method in ("topk", "topp")
        topk = config.sample_method == "topk"
        self.sample_param = self.top_k if topk else self.p          # UMGen.py:118-126
        self.sample_param_map = self.top_k_map if topk else self.p_map
        self.rule_constrain = bool(g("rule_constrain", False))
        self.ego_tokenlizer, self.ego_norm = g("ego_tokenlizer"), g("ego_norm")
        self.box3d_tokenlizer, self.agent_norm = g("box3d_tokenlizer"), g("agent_norm")
        if config.seq_len != 2207 or config.n_embd != 768 or config.n_head != 16:
            raise ValueError("the B200 engine is built for the 2207-token frame with n_embd 768 / 16 heads")
        if not (g("split_map_tar", True) and g("split_box_tar", True) and g("sample_img", True) and g("map_transform", True)):
            raise ValueError("only the evaluation configuration (split map/box TAR, image sampling, map transform) is supported")
        self.model_cfg = ModelConfig(
            n_tar_layer=config.n_tar_layer, n_oar_layer=config.n_oar_layer, n_ego_tar_layer=config.n_ego_tar_layer,
            n_ego_ca_layer=config.n_ego_ca_layer, n_map_tar_layer=config.n_map_tar_layer, n_box_tar_layer=config.n_box_tar_layer,
            cond_frame=config.cond_frame, max_frame_len=config.max_frame_len, rule_constrain=self.rule_constrain,
            merage_ar_tar=bool(g("merage_ar_tar", True)))
        cpu_params = g("device_set", None) == torch.device("cpu")   # evaluate.py:182 -> sinusoid tables become Parameters
        codebooks = {}
        for key, path in (("map_codebook.weight", g("map_codebook")), ("img_codebook.weight", g("img_codebook"))):
            if path is not None:
                codebooks[key] = torch.load(path, map_location="cpu").float()       # UMGen.py:248-253
        tree = _Tree()
        self._fixed: Dict[str, torch.Tensor] = {}
        # config.skip_init (an extension): parameters that a load_state_dict / weight broadcast will overwrite anyway are created as
        # zero-stride views of one zero, so a 2.4 B-parameter module costs neither the random draws nor 9.8 GB of host memory
        skip_init = self._skip_init = bool(g("skip_init", os.environ.get("UMGEN_SKIP_INIT") == "1"))
        for key, shape, kind in synth.param_specs(self.model_cfg):
            if key in codebooks:
                t = codebooks[key]
            elif skip_init and kind in ("linear", "bias", "emb", "ln", "codebook"):
                t = torch.zeros((), dtype=torch.float32).expand(shape)
            else:
                t = synth.make_param(key, shape, kind, seed=0)
            if kind in ("sin0", "sin1024", "gridpos") and not cpu_params:
                self._fixed[key] = t                                 # plain tensors in the reference when built on cuda
                continue
            tree.add(key, t, buffer=(kind == "scale"))
        # expose the hierarchy at the top level so state_dict keys carry no extra prefix
        for name, mod in tree._modules.items():
            self.add_module(name, mod)
        for name, prm in tree._parameters.items():
            self.register_parameter(name, prm)
        self._engine = None
        self._batch_engines = {}
        self._rollouts = 0
        print("number of parameters: %.2fB" % (sum(p.numel() for p in self.parameters()) / 1e9))

    # ---- engine management ------------------------------------------------------------------------------------
    def _rollout_seed(self, seed: Optional[int]) -> int:
        """Seed of one inference() call's counter-based random stream (Philox keyed by seed; counters = frame index, position, draw).
        `seed` argument > `self.seed` attribute > torch.initial_seed() mixed with the number of rollouts this module has run and the
        process rank, so different scenes and data-parallel ranks draw different uniforms while `torch.manual_seed` keeps runs repeatable."""
        if seed is None:
            seed = getattr(self, "seed", None)
        if seed is None:
            rank = int(os.environ.get("RANK", "0"))
            seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * (self._rollouts + 1) + 0xBF58476D1CE4E5B9 * rank) % (1 << 64)
        self._rollouts += 1
        return int(seed)

    def _sample_config(self, seed: int = 0) -> SampleConfig:
        method = self.config.sample_method
        if method == "topk":
            return SampleConfig(method="topk", top_k=int(self.sample_param), top_k_map=int(self.sample_param_map),
                                top_k_image=int(self.topk_image), temp=float(self.sfmx_temp), seed=seed)
        return SampleConfig(method="topp", p=float(self.sample_param), p_map=float(self.sample_param_map),
                            top_k_image=self.topk_image, temp=float(self.sfmx_temp), seed=seed)

    def _get_engine(self, seed: int = 0, scenes: int = 1):
        """The engine for `scenes` scenes per launch (1: UMGenEngine, the reference's batch-1 loop; more: SceneBatchEngine)."""
        from umgen_b200.engine import SceneBatchEngine, UMGenEngine
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else dev
        if scenes > 1:
            eng = self._batch_engines.get(scenes)
            if eng is None or eng.dev != dev:
                one = self._get_engine(seed)          # shares its weights, working buffers and decode stream with the batch engine's first scene
                eng = SceneBatchEngine.around(one, scenes)
                self._batch_engines[scenes] = eng
            for k, e in enumerate(eng.engines):
                e.sample = dataclasses.replace(self._sample_config(seed), seed=int(seed) + k)
            return eng
        if self._engine is None or self._engine.dev != dev:
            sd = dict(self.state_dict())
            sd.update(self._fixed)
            self._engine = UMGenEngine(sd, self.model_cfg, self._sample_config(seed), device=dev)
            self._batch_engines = {}
        self._engine.sample = self._sample_config(seed)             # attributes may be edited after construction (greedy recipe)
        return self._engine

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._engine = None                                          # packed device copies are stale
        self._batch_engines = {}
        if self._skip_init:                                          # placeholders cannot be copied into: adopt the checkpoint's tensors
            kw.setdefault("assign", True)
        return super().load_state_dict(state_dict, strict=strict, **kw)

    # ---- UMGen.inference (UMGen.py:1542-1671) -----------------------------------------------------------------
    def inference(self, new_frames: int, cond_frames: int = 1, input_cond_frames: int = -1, pred_task: str = "image",
                  input_cond_tokens: Optional[Dict[str, torch.Tensor]] = None, init_tokens: Optional[Dict[str, torch.Tensor]] = None,
                  cond_on_tar: bool = False, test_map_affine: bool = False, max_objects=100, control_test=False,
                  seed: Optional[int] = None, **kwargs) -> Dict[str, np.ndarray]:
        """`seed` (an extension; the reference draws from torch's global generator): see _rollout_seed.
        Tokens with a leading batch axis B > 1 (another extension: the reference asserts batch 1, UMGen.py:907,1093) are B scenes generated in
        lockstep, one decode launch per frame for all of them (umgen_b200.engine.SceneBatchEngine); scene k draws from the stream of seed + k."""
        assert pred_task in self.task_name_id
        scenes = int(input_cond_tokens["pose"].shape[0])
        return self._get_engine(self._rollout_seed(seed), scenes).inference(new_frames, cond_frames, input_cond_frames, pred_task, input_cond_tokens, init_tokens,
                                            cond_on_tar, test_map_affine, max_objects, control_test, **kwargs)
