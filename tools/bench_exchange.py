"""Microbenchmark of the cross-CTA exchange primitive (debug aid): load/store flavours, replication, HBM load."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from umgen_b200 import capi

lib = capi.lib()
lib.umgen_debug_exchange_bench.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
lib.umgen_debug_pingpong.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
dev = torch.device("cuda:0")
src = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
out = torch.zeros(160, dtype=torch.int64, device=dev)
names = {0: "relaxed.gpu", 1: "cg", 2: "cv/wt", 3: "volatile"}
for fl in range(4):
    buf = torch.zeros(256, dtype=torch.float32, device=dev)
    rc = lib.umgen_debug_pingpong(buf.data_ptr(), 20000, fl, out.data_ptr(), None)
    assert rc == 0, lib.umgen_last_error()
    torch.cuda.synchronize()
    print(f"pingpong {names[fl]:12s}: {out[0].item() / 20000:.0f} ns per round trip", flush=True)
iters = 2000
for nvals in (768, 3072):
    for fl in range(4):
        for variant, krep, sbytes in [(0, 16, 0), (0, 16, 49152)]:
            buf = torch.zeros(64 * 2 * nvals + 64, dtype=torch.float32, device=dev)
            flags = torch.zeros(160 * 32, dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            rc = lib.umgen_debug_exchange_bench(buf.data_ptr(), flags.data_ptr(), iters, krep, variant | (fl << 4), src.data_ptr(), sbytes, nvals, out.data_ptr(), None)
            assert rc == 0, lib.umgen_last_error()
            torch.cuda.synchronize()
            per = out[:148].double().median().item() / iters
            print(f"nvals={nvals} {names[fl]:12s} krep={krep:2d} stream={sbytes:6d} B/iter/SM: {per:7.0f} ns/exchange", flush=True)
