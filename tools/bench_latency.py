"""Instruction latencies on the decode kernel's serial path (tools/csrc/lat_bench.cu).  python -m umgen_b200.build --tools first."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umgen_b200 import build
L = ctypes.CDLL(build.TOOLS_LIB)
L.umgen_tools_lat_bench.argtypes = [ctypes.c_void_p] * 3
out = torch.zeros(16, dtype=torch.int64, device="cuda")
sink = torch.zeros(4, device="cuda")
for _ in range(2):
    assert L.umgen_tools_lat_bench(out.data_ptr(), sink.data_ptr(), None) == 0
    torch.cuda.synchronize()
names = ["HMMA.16816 dependent chain", "HMMA.16816 4 independent chains (per MMA)", "shfl dependent", "ld.shared dependent (pointer chase)",
         "FFMA dependent", "ex2 dependent", "bar.sync 12 warps back to back", "LayerNorm-style round (STS, bar, 12 LDS, bar)"]
for n, v in zip(names, out.cpu().tolist()):
    print(f"{n:55s} {v:5d} cycles")
