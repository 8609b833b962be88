"""Microbenchmark of the intra-cluster exchange primitives (debug aid, see csrc/dsmem_bench.cu)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from umgen_b200 import capi

lib = capi.lib()
lib.umgen_debug_dsmem_bench.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
out = torch.zeros(4, dtype=torch.int64, device="cuda")
names = {0: "ping-pong st.shared::cluster line + local poll", 1: "ping-pong st.async complete_tx + try_wait", 2: "ping-pong remote arrive.release + try_wait.acquire",
         3: "all-to-all 48 lines/pair, remote stores + tight local polls", 4: "all-to-all, polls paused 64 cycles", 5: "all-to-all, stores + barrier + remote arrive",
         6: "remote store per thread + 2 block barriers (does bar.sync wait for the store to be acknowledged?)", 7: "2 block barriers alone",
         8: "all-to-all, scattered pattern (thread u -> rank u % cluster size), tight local polls"}
for ncl in (1, 8, -1):      # -1: one cluster of 16 CTAs
    for mode in range(9):
        rc = lib.umgen_debug_dsmem_bench(out.data_ptr(), 2000, mode, ncl, None)
        assert rc == 0, lib.umgen_last_error()
        torch.cuda.synchronize()
        print(f"clusters={abs(ncl)} x {16 if ncl < 0 else 8} CTAs mode {mode} ({names[mode]}): {out[0].item()} cycles per {'round trip' if mode < 3 else 'exchange'}", flush=True)
