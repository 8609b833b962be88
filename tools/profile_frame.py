"""One generated frame of UMGen_Large with the sequential schedule between cudaProfilerStart / Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/profile_frame.py
(the profiler serialises kernels, so the look-ahead / overlap schedules are switched off).  Optional argument: layers per stack (debug)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umgen_b200 import synth  # noqa: E402
from umgen_b200.config import MODS, ModelConfig, SampleConfig  # noqa: E402
from umgen_b200.engine import UMGenEngine  # noqa: E402

cfg = ModelConfig.tiny(int(sys.argv[1])) if len(sys.argv) > 1 else ModelConfig.large()
dev = torch.device("cuda:0")
eng = UMGenEngine(synth.DeviceParams(cfg, seed=0, device=dev), cfg, SampleConfig.greedy(), device=dev)
eng.lookahead = eng.overlap = False
eng.check_status = False
scene = synth.make_scene(seed=1, n_frames=cfg.cond_frame)
tok = {m: scene[m][0].to(torch.int32).to(dev) for m in MODS}
eng.frame_device(tok)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.frame_device(tok)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one frame")
