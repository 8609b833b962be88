"""Microbenchmark: per-SM HBM -> shared-memory streaming rate with few SMs active (debug aid, see csrc/stream_bench.cu)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from umgen_b200 import capi

lib = capi.lib()
lib.umgen_debug_stream_bench.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
out = torch.zeros(4, dtype=torch.int64, device="cuda")
sink = torch.zeros(4, dtype=torch.int32, device="cuda")
PER = 36864 * 1280
src = torch.empty(64 * PER, dtype=torch.uint8, device="cuda")
src.fill_(1)
CASES = [(cs, ncl, chunk, ring, 0) for cs, ncl in ((8, 1), (16, 1), (8, 2), (16, 2), (8, 4), (8, 8))
         for chunk, ring in ((8192, 192), (16384, 192), (32768, 192), (65536, 192), (16384, 96), (4096, 192))]
if len(sys.argv) > 1 and sys.argv[1] == "prefetch":      # the one-cluster decode kernel's ring (4 x 36 KB) with L2 prefetch distances
    CASES = [(16, 1, 36864, 144, pf) for pf in (0, 4, 8, 16, 32, 64)] + [(16, 1, 18432, 144, pf) for pf in (0, 16, 64)] + \
            [(16, 1, 36864, 108, pf) for pf in (0, 16)] + [(16, 1, 36864, 180, pf) for pf in (0, 16)]
for cs, ncl, chunk, ring, pf in CASES:
    if True:
        nst = min(32, ring * 1024 // chunk)
        rc = lib.umgen_debug_stream_bench(src.data_ptr(), PER, chunk, nst, cs, ncl, out.data_ptr(), sink.data_ptr(), pf, None)
        if rc != 0:
            print(f"cluster {cs} x {ncl} chunk {chunk}: {lib.umgen_last_error().decode()}", flush=True)
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.umgen_debug_stream_bench(src.data_ptr(), PER, chunk, nst, cs, ncl, out.data_ptr(), sink.data_ptr(), pf, None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        o = out.cpu().tolist()
        print(f"cluster {cs} x {ncl} clusters, chunk {chunk:6d} x {nst:2d} stages, L2 prefetch {pf:2d} ahead: slowest CTA {PER / o[0]:.1f} B/clk/SM, CTA0 {PER / o[1]:.1f} B/clk; "
              f"{cs * ncl * PER / ms / 1e6:.0f} GB/s total, {PER / ms / 1e6:.1f} GB/s per SM (max active clusters {o[2]})", flush=True)
