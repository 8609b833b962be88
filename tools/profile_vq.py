#!/usr/bin/env python
"""Where a VQ decode goes: every library call of one eager decode_code (4 frames) bracketed by CUDA events on the launching stream,
summed by (call, shape).  python tools/profile_vq.py [map|image] [frames]"""
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umgen_b200 import ops, synth  # noqa: E402
from umgen_b200.vq import VQDecoder  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "image"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
log = []


def wrap(name, label):
    fn = getattr(ops, name)

    def inner(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **kw)
        e1.record()
        log.append((label(*a, **kw), e0, e1))
        return r
    setattr(ops, name, inner)


wrap("gemm", lambda a, w, b, out, epi, resid=None: f"gemm M={a.shape[0]} N={w.shape[0]} K={a.shape[1]} epi={epi}")
wrap("conv3x3", lambda x, w, b, out, B, H, W, Cin, epi=0, resid=None: f"conv3x3 {B}x{H}x{W} {Cin}->{w.shape[0]} epi={epi}")
wrap("im2col3x3", lambda x, a, B, H, W, Cin, kp, up: f"im2col {B}x{H}x{W} {Cin}")
wrap("upsample2x", lambda x, out, B, H, W, Cc: f"upsample2x {B}x{H}x{W} {Cc}")
wrap("groupnorm_slab", lambda x, g, b, y, sc, B, HW, Cc, sw: f"groupnorm {B}x{HW} {Cc}")
wrap("groupnorm", lambda x, g, b, y, sc, B, HW, Cc, sw: f"groupnorm(old) {B}x{HW} {Cc}")
wrap("softmax_rows", lambda s, p, sc: f"softmax {s.shape[0]}x{s.shape[1]}")
wrap("transpose_f16", lambda x, out: f"transpose {x.shape[0]}x{x.shape[1]}")
wrap("conv_out3x3", lambda x, w, b, out, B, H, W, Cin, Cout: f"conv_out {B}x{H}x{W} {Cin}->{Cout}")
wrap("vq_gather", lambda idx, t, out: "vq_gather")

dec = VQDecoder(synth.make_vq_state_dict(kind, seed=1), kind)
dec.use_graph = False
h, w = dec.cfg["grid"]
code = torch.randint(0, 8192, (frames, h, w), generator=torch.Generator().manual_seed(3))
dec.decode_code(code)
dec.decode_code(code)
torch.cuda.synchronize()
log.clear()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
dec.decode_code(code)
t1.record()
torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
for label, e0, e1 in log:
    agg[label][0] += 1
    agg[label][1] += e0.elapsed_time(e1)
tot = t0.elapsed_time(t1)
inside = sum(v[1] for v in agg.values())
print(f"# {kind} decoder, {frames} frames, eager launches: {tot:.3f} ms wall on the stream, {inside:.3f} ms inside library calls, {len(log)} calls")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  n={n:3d}  {1e3 * ms / n:8.1f} us/call  {k}")
