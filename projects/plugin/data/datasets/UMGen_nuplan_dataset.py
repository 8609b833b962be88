"""Drop-in for the reference's ``projects/plugin/data/datasets/UMGen_nuplan_dataset.py``: the ``NuPlanTokenDataset`` that
``configs/UMGen_config_evaluation.py:168-170`` imports, ``evaluate.py:154`` names as ``dataset_type`` and ``tools/infer_fun.py:189-213``
configures, backed by ``umgen_b200.dataset.NuPlanTokenScenes`` (bit-identical tokens: tests/test_dataset.py).

The constructor keeps the reference's signature.  ``transforms`` is accepted and must be the evaluation list (SplitAttriute, Normalize,
MergeAttribute, Normalize_Standard, BBox3DTokenizer, DigitalBinsTokenizer, ToTensor) or None: its arithmetic is built into the scenes
class; the image-loading options of the training pipeline (``return_ori_image``, ``img_transform``, lidar boxes) are not part of the
inference path and raise."""
from typing import List, Optional, Union

from torch.utils import data

from projects.registry import DATASETS
from umgen_b200.dataset import NuPlanTokenScenes

EVAL_TRANSFORMS = ["SplitAttriute", "Normalize", "MergeAttribute", "Normalize_Standard", "BBox3DTokenizer", "DigitalBinsTokenizer", "ToTensor"]


@DATASETS.register_module()
class NuPlanTokenDataset(data.Dataset):
    def __init__(self, data_root: Union[str, List[str]], training: bool, block_size: int, views: List[str],
                 categories_file: str = "projects/configs/category.txt", sampling_gap: Optional[int] = 1, transforms: Optional[List] = None,
                 img_transform=None, inference_flag=False, start_index=10, sp_list=None, sample_img=False, return_ori_image=False,
                 sample_others=True, return_scene_name=False, return_lidar_box=False, return_lidar_box_id=False, control_test=False,
                 vae_token_path=None, **kwargs):
        super().__init__()
        if return_ori_image or return_lidar_box or return_lidar_box_id or not sample_others:
            raise NotImplementedError("NuPlanTokenDataset drop-in: only the token outputs of the inference path are provided")
        if transforms is not None:
            names = [type(t).__name__ for t in getattr(transforms, "transforms", transforms)]
            if names != EVAL_TRANSFORMS:
                raise NotImplementedError(f"NuPlanTokenDataset drop-in implements the evaluation transform list {EVAL_TRANSFORMS}, got {names}")
        self.training, self.block_size, self.sampling_gap, self.control_test = training, block_size, sampling_gap, control_test
        self._scenes = NuPlanTokenScenes(data_root, block_size=block_size, sampling_gap=sampling_gap, start_index=start_index,
                                         inference_flag=inference_flag, views=views, categories_file=categories_file, sample_img=sample_img,
                                         control_test=control_test, return_scene_name=return_scene_name)
        self.files = self._scenes.files

    @property
    def mode(self):
        return "train" if self.training else "test"

    def __len__(self):
        return len(self._scenes)

    def __getitem__(self, idx):
        return self._scenes[idx]
