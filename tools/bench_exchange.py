"""Microbenchmark of the cross-CTA exchange primitive (debug aid)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from umgen_b200 import capi

lib = capi.lib()
lib.umgen_debug_exchange_bench.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
dev = torch.device("cuda:0")
src = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
out = torch.zeros(160, dtype=torch.int64, device=dev)
iters = 2000
for nvals in (768, 3072):
    for variant, krep, sbytes in [(0, 1, 0), (0, 4, 0), (0, 16, 0), (0, 32, 0), (1, 1, 0), (0, 16, 16384), (0, 16, 49152), (0, 1, 49152), (1, 1, 49152)]:
        buf = torch.zeros(64 * 2 * nvals + 64, dtype=torch.float32, device=dev)
        flags = torch.zeros(160 * 32, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        rc = lib.umgen_debug_exchange_bench(buf.data_ptr(), flags.data_ptr(), iters, krep, variant, src.data_ptr(), sbytes, nvals, out.data_ptr(), None)
        assert rc == 0, lib.umgen_last_error()
        torch.cuda.synchronize()
        ns = out[:148].double()
        per = ns.median().item() / iters
        bw = 148 * sbytes / per if per > 0 else 0
        print(f"nvals={nvals} variant={'LL' if variant == 0 else 'flag'} krep={krep:2d} stream={sbytes:6d} B/iter/SM: {per:7.0f} ns/exchange  (stream {bw:6.0f} GB/s)", flush=True)
