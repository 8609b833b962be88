"""What follows a rollout, per scene: token pickle + value decode, the VQ pixel decoders in 6-frame pieces (GPU), the scene
video (host).  50-frame scene (20 conditioning + 30 generated) at evaluate.py's sizes.  python tools/bench_scene_tail.py [frames]"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from umgen_b200 import postprocess, runner, synth  # noqa: E402
from umgen_b200.visualize import SceneVideo  # noqa: E402
from umgen_b200.vq import Imagedecoder, Mapdecoder  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 50
tmp = tempfile.mkdtemp()
scene = synth.make_scene(seed=5, n_frames=T)
tokens = {m: scene[m][:, :T].numpy().astype(np.int64) for m in ("pose", "map", "bbox3d", "image")}
md, idec = Mapdecoder(synth.make_vq_state_dict("map", seed=1)), Imagedecoder(synth.make_vq_state_dict("image", seed=1))
video = SceneVideo(video_save_path=os.path.join(tmp, "clips/"), video_pretext="UMGen", width=512, height=512, project_name="UMGen_infer", spe_text="tail",
                   addtion_ego=True, cond_frames=20, put_text=True)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.time()
    postprocess.save_tokens(tokens, os.path.join(tmp, "tokens"), f"scene{rep}")
    t1 = time.time()
    decoded = postprocess.decode_tokens(dict(tokens), {m: tokens[m] for m in ("pose", "bbox3d")}, md, idec)
    torch.cuda.synchronize()
    t2 = time.time()
    path = runner.write_scene_video(video, decoded, f"scene{rep}")
    t3 = time.time()
    print(f"rep {rep}: {T} frames: token pickle {1e3 * (t1 - t0):.1f} ms, decode_tokens (values + map / image pixels, results on the host) {1e3 * (t2 - t1):.1f} ms, "
          f"scene video {1e3 * (t3 - t2):.1f} ms ({os.path.getsize(path) / 1e6:.1f} MB) -> {t3 - t0:.2f} s per scene", flush=True)
