import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from tests._cases import ROLLOUT_CASES
from tests.test_engine_gpu import build, sampled_positions
from umgen_b200 import synth
name = "video_L1"
spec = ROLLOUT_CASES[name]
g = np.load(f"tests/golden/rollout_{name}.npz")
eng = build(spec)
scene = synth.make_scene(seed=spec["scene_seed"], n_frames=spec["input_frames"])
out = eng.inference(1, spec["cond_frames"], spec["input_cond_frames"], input_cond_tokens=scene)
tr = eng.trace[0]
pos = sampled_positions()
picks = tr.tokens.cpu().numpy()[[p - 1 for p in pos]]
stream = g["input_stream"][0]
bad = np.nonzero(picks != stream)[0]
print("n bad", bad.size, "first", [pos[i] for i in bad[:10]])
logits = tr.logits.cpu()
for i in list(range(bad[0] - 3, bad[0] + 2)):
    p = pos[i]
    V = 1028 if 1033 <= p <= 1692 else 8192
    tv, ti = torch.topk(logits[p - 1, :V], 4)
    print(p, "mine", ti.tolist(), [round(x, 4) for x in tv.tolist()], "pick", picks[i], "| gold", g["ar_top_ids"][0][i][:4].tolist(),
          [round(float(x), 4) for x in g["ar_top_vals"][0][i][:4]], "stream", stream[i])
print("status", tr.status[:4], "gold n_tar calls", g["n_tar_bbox_calls"])
prev = scene["bbox3d"][0, spec["input_cond_frames"] - 1]
print("prev tokens slot 21/22:", prev[21 * 11: 23 * 11].tolist())
print("mine bbox out slots 20-22", out["bbox3d"][0, -1][20 * 11: 23 * 11].tolist())
print("gold bbox out slots 20-22", g["out_bbox3d"][0, spec["input_cond_frames"]][20 * 11: 23 * 11].tolist())
