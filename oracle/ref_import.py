"""Import the UNMODIFIED reference (``/root/reference``) on CPU.  TEST INFRASTRUCTURE ONLY.

Used in the build container to (a) validate ``oracle/umgen_oracle.py`` against the real
implementation and (b) generate the golden vectors under ``tests/golden/``
(``oracle/make_golden.py``).  ``/root/reference`` does not exist on the GPU box, so nothing
that runs there imports this module.

Recipe: three stub modules for packages that are not
installed (``mmcv``, ``deepspeed``, ``torchmetrics``) and two monkey-patches:
``torch.Tensor.cuda`` -> identity (the model hard-codes ``.cuda()``), and
``projects.models.module.flash_attn_func`` -> fp32 math attention with bottom-right causal
alignment (flash-attn is CUDA-only).  No reference source is copied.
"""
from __future__ import annotations

import contextlib
import copy
import os
import sys
import types
from argparse import Namespace

import torch

REF_ROOT = os.environ.get("UMGEN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "projects", "models"))


class _Registry:
    def __init__(self, name):
        self.name, self.module_dict = name, {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        return deco(module) if module is not None else deco


def _build_from_cfg(cfg, registry, default_args=None):
    cfg = dict(cfg)
    t = cfg.pop("type")
    cls = registry.module_dict[t] if isinstance(t, str) else t
    return cls(**cfg, **(default_args or {}))


def _install_stubs():
    if "mmcv" not in sys.modules:
        mmcv = types.ModuleType("mmcv")
        utils = types.ModuleType("mmcv.utils")
        utils.Registry, utils.build_from_cfg = _Registry, _build_from_cfg
        mmcv.utils = utils
        mmcv.imfrombytes = lambda *a, **k: None
        sys.modules["mmcv"], sys.modules["mmcv.utils"] = mmcv, utils
    if "deepspeed" not in sys.modules:
        ds = types.ModuleType("deepspeed")
        ck = types.ModuleType("deepspeed.checkpointing")
        ck.is_configured = lambda: False
        ds.checkpointing = ck
        sys.modules["deepspeed"], sys.modules["deepspeed.checkpointing"] = ds, ck
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")
        tm.Metric = type("Metric", (), {})
        sys.modules["torchmetrics"] = tm
    if "flash_attn" not in sys.modules:
        try:
            import flash_attn  # noqa: F401
        except Exception:
            fa = types.ModuleType("flash_attn")
            fa.flash_attn_func = None
            sys.modules["flash_attn"] = fa


def _math_attention(q, k, v, dropout_p=0.0, softmax_scale=None, causal=False, **_):
    """(B, T, H, D) in/out; bottom-right aligned causal mask like flash-attn >= 2.1."""
    qf, kf, vf = (t.float().transpose(1, 2) for t in (q, k, v))
    att = (qf @ kf.transpose(-1, -2)) * float(softmax_scale)
    if causal:
        tq, tk = qf.shape[-2], kf.shape[-2]
        keep = torch.ones(tq, tk, dtype=torch.bool).tril(diagonal=tk - tq)
        att = att.masked_fill(~keep, float("-inf"))
    return (torch.softmax(att, dim=-1) @ vf).transpose(1, 2).contiguous().to(q.dtype)


@contextlib.contextmanager
def reference_cwd():
    old = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        yield
    finally:
        os.chdir(old)


_loaded = {}


def load(native: bool = False):
    """Returns a namespace with the reference's modules (imported once).  native=True (a box that has both the reference tree and a GPU:
    bench.py --impl reference-gpu) leaves the reference's own CUDA path alone: no `.cuda()` patch, flash_attn as installed."""
    if _loaded:
        assert _loaded["native"] == native, "the reference was already imported in the other mode in this process"
        return _loaded["ns"]
    assert available(), f"reference tree not found at {REF_ROOT}"
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # our own repo also has a `projects` package; make sure the reference's wins in this process
    for k in [k for k in sys.modules if k == "projects" or k.startswith("projects.")]:
        del sys.modules[k]
    with reference_cwd():
        import projects.models.module as ref_module
        import projects.models.UMGen as ref_umgen
        import projects.plugin.misc.misc as ref_misc
        from projects.plugin.data.transforms import normalize as ref_norm
        from projects.plugin.data.transforms import tokenizer as ref_tok
    if not native:
        ref_module.flash_attn_func = _math_attention
        torch.Tensor.cuda = lambda self, *a, **k: self
    ns = Namespace(module=ref_module, umgen=ref_umgen, misc=ref_misc, norm=ref_norm, tok=ref_tok)
    _loaded["ns"], _loaded["native"] = ns, native
    return ns


def reference_config(layers=None, native: bool = False, **over) -> Namespace:
    """The Namespace evaluate.py hands to UMGen(config) for ``--model_scale larger``
    (configs/UMGen_config_evaluation.py:344-430 after tools/infer_fun.py:84-159), rebuilt from the
    reference's own tokenizer/normaliser classes.  ``layers`` overrides every stack depth."""
    ns = load(native) if not _loaded else _loaded["ns"]
    with reference_cwd():
        ego_tok = ns.tok.DigitalBinsTokenizer(bins=[(-1.0, 1.0, 1024)], data_key="pose", seq_len=3,
                                              special_tokens=None, start=0)
        box_tok = ns.tok.BBox3DTokenizer(bins=[(0.0, 1.0, 1024)], category_file="projects/configs/category.txt",
                                         start=0, special_tokens=[], pad_to_length=60, target_key=["bbox3d"],
                                         shift_object_order_pro=0)
    data_key = ("bbox_posi_x", "bbox_posi_y", "bbox_posi_z", "bbox_wlh_l", "bbox_wlh_w", "bbox_wlh_h",
                "bbox_yaw", "bbox_speed_x", "bbox_speed_y", "bbox_speed_z")
    rng = {"bbox_posi_x": (-64, 64), "bbox_posi_y": (-64, 64), "bbox_posi_z": (-5, 5), "bbox_wlh_l": (0, 15),
           "bbox_wlh_w": (0, 4), "bbox_wlh_h": (0, 5), "bbox_yaw": (-3.14, 3.14), "bbox_speed_x": (-20, 20),
           "bbox_speed_y": (-15, 15), "bbox_speed_z": (-0.3, 0.3)}
    agent_norm = ns.norm.Normalize(data_key=data_key, max_min=rng, min_max_standard_key=[])
    ego_norm = ns.norm.Normalize_Standard(data_key="pose", mean=[0, 0, 0], std=[10.0, 4.0, 1.0])
    token_len = {"bbox3d": box_tok.seq_len + 2, "map": 1026, "pose": ego_tok.seq_len + 2, "image": 514}
    cfg = Namespace(
        pred_task="pose_map_bbox3d_image", max_frame_len=100, cond_frame=20,
        pose_vocab_size=1024, map_vocab_size=8192, img_vocab_size=8192, bbox3d_vocab_size=1028,
        bos_eos={"pose": [0, 1], "map": [2, 3], "bbox3d": [4, 5], "image": [6, 7]}, aux_vocab_size=8,
        box3d_tokenlizer=box_tok, agent_norm=agent_norm, ego_tokenlizer=ego_tok, ego_norm=ego_norm,
        task={"pose_map_bbox3d_image": ["pose", "map", "bbox3d", "image"],
              "pose_map_bbox3d": ["pose", "map", "bbox3d"], "pose_map": ["pose", "map"]},
        task_prob=None, task_name_id={"pose_map_bbox3d_image": 6}, task_num=7,
        vocab_len={"bbox3d": len(box_tok), "map": 2, "pose": len(ego_tok) + 2, "image": 2},
        token_len=token_len, seq_len=2207,
        map_codebook="projects/tokenizer/weights/map_codebook.pth",
        img_codebook="projects/tokenizer/weights/img_codebook.pth",
        pad_to_length=60, n_tar_layer=36, n_oar_layer=36, n_ego_tar_layer=12, n_ego_ca_layer=12,
        n_map_tar_layer=24, n_box_tar_layer=24, n_head=16, n_embd=768, n_img_embd=16, n_map_embd=16,
        dropout=0, ar_dropout=0, add_posi_embedd=True, add_spatial_pos_embedd_on_map=True, bias=False,
        top_k=5, top_k_map=5, sample_method="topk", p=0.4, sfmx_temp=1.0, flash_attention=True,
        cond_prob=1, box_transform=False, add_t_pos=False, split_map_tar=True, split_map_ar=False,
        split_box_tar=True, split_image_ar=False, only_ar=False, sample_img=True, map_transform=True,
        n_posiembed=0, n_step=1, block_size=21, merage_ar_tar=True, train_only_ego=False,
        rule_constrain=True, num_attritube=10, device_set=torch.device("cpu"),
    )
    if layers is not None:
        for k in ("n_tar_layer", "n_oar_layer", "n_ego_tar_layer", "n_ego_ca_layer", "n_map_tar_layer",
                  "n_box_tar_layer"):
            setattr(cfg, k, layers)
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def build_reference_model(cfg: Namespace, state_dict=None, greedy: bool = False):
    ns = _loaded["ns"] if _loaded else load()
    with reference_cwd():
        model = ns.umgen.UMGen(copy.copy(cfg)).eval()
    if state_dict is not None:
        missing = model.load_state_dict(state_dict, strict=False)
        assert not missing.unexpected_keys, missing.unexpected_keys[:5]
        assert not missing.missing_keys, missing.missing_keys[:5]
    if greedy:                      # greedy == top-k with k = 1 everywhere
        model.top_k = model.sample_param = model.sample_param_map = model.topk_image = 1
    return model
