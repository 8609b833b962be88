"""Static geometry of the UMGen scene-token stream and the model depth configuration.

Mirrors configs/UMGen_config_evaluation.py:27-38, 284-290 and tools/infer_fun.py:84-159 of the
reference (paths relative to /root/reference/projects)."""
from __future__ import annotations

from dataclasses import dataclass, asdict

MODS = ("pose", "map", "bbox3d", "image")
CONTENT_LEN = {"pose": 3, "map": 1024, "bbox3d": 660, "image": 512}
TOKEN_LEN = {m: CONTENT_LEN[m] + 2 for m in MODS}
BOS_EOS = {"pose": (0, 1), "map": (2, 3), "bbox3d": (4, 5), "image": (6, 7)}
SEQ_LEN = 2207
MOD_OFFSET = {"pose": 0, "map": 5, "bbox3d": 1031, "image": 1693}     # 0-indexed start (bos) of each modality
N_SLOTS, N_ATTR = 60, 11
PAD_TOKEN = 1027
TASK_ID = 6
N_EMBD, N_HEAD, HEAD_DIM = 768, 16, 48
VOCAB = {"pose": 1024, "map": 8192, "bbox3d": 1028, "image": 8192}
TASK_MODS = {"ego": MODS, "map": ("pose", "map"), "box": ("pose", "map", "bbox3d"), "full": MODS}
TASK_SEQ = {"ego": 2207, "map": 1031, "box": 1693, "full": 2207}


@dataclass(frozen=True)
class ModelConfig:
    n_tar_layer: int = 36
    n_oar_layer: int = 36
    n_ego_tar_layer: int = 12
    n_ego_ca_layer: int = 12
    n_map_tar_layer: int = 24
    n_box_tar_layer: int = 24
    cond_frame: int = 20
    max_frame_len: int = 100
    rule_constrain: bool = True
    merage_ar_tar: bool = True

    @staticmethod
    def large() -> "ModelConfig":
        """UMGen_Large: evaluate.py --model_scale larger."""
        return ModelConfig()

    @staticmethod
    def tiny(layers: int = 1, **kw) -> "ModelConfig":
        return ModelConfig(n_tar_layer=layers, n_oar_layer=layers, n_ego_tar_layer=layers,
                           n_ego_ca_layer=layers, n_map_tar_layer=layers, n_box_tar_layer=layers, **kw)

    def to_dict(self):
        return asdict(self)


@dataclass
class SampleConfig:
    """Sampling knobs of UMGen.__init__ (UMGen.py:99-126)."""
    method: str = "topk"
    top_k: int = 5
    top_k_map: int = 5
    top_k_image: int = 16
    p: float = 0.4
    p_map: float = 0.4
    temp: float = 1.0
    seed: int = 0

    @staticmethod
    def greedy() -> "SampleConfig":
        return SampleConfig(top_k=1, top_k_map=1, top_k_image=1)
