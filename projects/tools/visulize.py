"""Drop-in ``Visulizer`` (reference projects/tools/visulize.py:425-577, 1607-1715): the constructor and the two methods ``UMGen_PL`` calls
(``visulize`` from generate_videos, ``vis_pred_video`` from generate_compare_videos; model_pl.py:61-73, 283-331), backed by
``umgen_b200.visualize.SceneVideo`` -- frames bit-identical to the reference's (tests/test_visualize.py)."""
from __future__ import annotations

from umgen_b200.visualize import SceneVideo, frame_label as add_frame_number, write_video_single  # noqa: F401  (decode_map.py's helpers live there too)


class Visulizer(SceneVideo):
    def __init__(self, video_save_path="output/videos/", video_pretext="test", width=256, height=256, project_name="test", spe_text="p=0.5",
                 save_video=True, addtion_ego=False, resort_attritube=None, bbox3d_arrow_length_scale=1, rotate_speed=False, map_type="token",
                 dataset="nuplan", cond_frames=20, put_text=True):
        if resort_attritube is not None or rotate_speed or map_type != "token" or dataset != "nuplan":
            raise NotImplementedError("only the configuration tools/model_pl.py builds is supported: nuPlan, token maps, velocities in the ego frame")
        super().__init__(video_save_path, video_pretext, width, height, project_name, spe_text, save_video, addtion_ego, bbox3d_arrow_length_scale,
                         cond_frames, put_text)
