"""Scene compositor throughput (umgen_b200/visualize.py) on a 50-frame scene at evaluate.py's sizes (512-pixel canvas, 256 x 256 map, 256 x 512 camera
image), beside the reference's Visulizer on the same inputs where the reference tree is mounted (TEST INFRASTRUCTURE: oracle/ref_import).  Host code,
one thread.  python tools/bench_visualize.py [frames]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tests._cases as C  # noqa: E402
from umgen_b200 import visualize as V  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 50
C.VISUALIZE_CASES["bench"] = (31, T, 512, 20, True, 256, (256, 512), 20)
d = C.visualize_inputs("bench")
tmp = tempfile.mkdtemp()
os.chdir(tmp)
vis = V.SceneVideo(video_save_path=os.path.join(tmp, "ours/"), video_pretext="UMGen", width=512, height=512, project_name="UMGen_infer", spe_text="bench",
                   addtion_ego=True, cond_frames=20, put_text=True)
for rep in range(2):
    t0 = time.time()
    frames = vis.compose(d["boxes"], d["pose"], d["real_pose"], d["maps"], d["image"], "bench")
    t1 = time.time()
    V.write_mp4(frames, os.path.join(tmp, "ours", "UMGen_bench.mp4"))
    t2 = time.time()
agents = sum(len(V.live_slots(b)) for b in d["boxes"])
print(f"ours      : {T} frames, {agents} agents: compose {1e3 * (t1 - t0):.0f} ms + mp4 {1e3 * (t2 - t1):.0f} ms = {T / (t2 - t0):.1f} frames/s")

from oracle import ref_import as R  # noqa: E402
if R.available():
    import cv2
    R.load()
    sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))
    with R.reference_cwd():
        import projects.tools.visulize as ref_vis
    cv2.destroyAllWindows = lambda: None
    rv = ref_vis.Visulizer(video_save_path=os.path.join(tmp, "ref/"), video_pretext="UMGen", width=512, height=512, project_name="UMGen_infer", spe_text="bench",
                           save_video=True, addtion_ego=True, cond_frames=20, put_text=True)
    for rep in range(2):
        t0 = time.time()
        rv.visulize(box=np.array([b.copy() for b in d["boxes"]], dtype=object), scene_name="bench", pose=d["pose"].copy(), real_pose=d["real_pose"].copy(),
                    maps={"map": d["maps"].clone()}, decoded_image=d["image"].clone())
        t1 = time.time()
    same = open(os.path.join(tmp, "ref", "UMGen_bench.mp4"), "rb").read() == open(os.path.join(tmp, "ours", "UMGen_bench.mp4"), "rb").read()
    print(f"reference : visulize() {1e3 * (t1 - t0):.0f} ms = {T / (t1 - t0):.1f} frames/s; mp4 files identical: {same}")
else:
    print("reference : tree not mounted")
