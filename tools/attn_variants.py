"""Time the tcgen05 attention kernel in experiment builds (python -m umgen_b200.build --variant TAG DEFINE ...): one subprocess per library."""
import glob, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
from umgen_b200 import ops
T, S = 20, 2207
qkv = torch.randn(T * S, 2304, device="cuda").half()
y = torch.zeros(T * S, 768, dtype=torch.float16, device="cuda")
for _ in range(3): ops.spatial_attention(qkv, y, T, S)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.spatial_attention(qkv, y, T, S)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("%%.3f ms  %%.0f TFLOP/s" %% (ms, 4.0 * S * S * 768 * T / ms / 1e9))
''' % root
libs = [None] + sorted(glob.glob(os.path.join(root, "umgen_b200/lib/libumgen_sm100.*.so")))
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["UMGEN_LIB"] = lib
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(os.path.basename(lib) if lib else "default", "->", (r.stdout.strip() or r.stderr.strip()[-300:]))
