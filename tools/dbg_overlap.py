"""Debug aid: does the box_tar pass really run beside the decode kernel?  Prints when the late TAR work and the decode kernel finish."""
import sys
import torch
sys.path.insert(0, ".")
from umgen_b200 import synth
from umgen_b200.config import MODS, ModelConfig, SampleConfig
from umgen_b200.engine import UMGenEngine
from umgen_b200.tar import TarEncoders

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = ModelConfig.tiny(layers, cond_frame=T)
sd = synth.make_state_dict(cfg, seed=0)
eng = UMGenEngine(sd, cfg, SampleConfig.greedy())
eng.check_status = False
eng.lookahead = False          # this tool looks at the box-pass schedule (engine.overlap)
scene = synth.make_scene(seed=1, n_frames=T)
tok = TarEncoders.to_device_tokens({m: scene[m][0] for m in MODS}, eng.dev)
orig_late = eng.tar.conditioning_late
orig_sig = eng.dec.signal_ready
marks = {}

def late(t):
    marks["late0"] = torch.cuda.Event(enable_timing=True); marks["late0"].record()
    r = orig_late(t)
    return r

def sig(flag, v):
    orig_sig(flag, v)
    marks["late1"] = torch.cuda.Event(enable_timing=True); marks["late1"].record()

eng.tar.conditioning_late = late
eng.dec.signal_ready = sig
for it in range(3):
    t0 = torch.cuda.Event(enable_timing=True); t0.record()
    eng.frame_device(tok)
    t1 = torch.cuda.Event(enable_timing=True); t1.record()
    torch.cuda.synchronize()
    late = (f"late pass starts at {t0.elapsed_time(marks['late0']):.1f} ms, signalled at {t0.elapsed_time(marks['late1']):.1f} ms" if "late1" in marks
            else "sequential schedule (an engine's first frame)")
    marks.clear()
    print(f"iter {it}: frame {t0.elapsed_time(t1):.1f} ms; {late}; status {eng.dec.status[:4].tolist()}", flush=True)
