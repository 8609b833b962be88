"""Per-tile timeline of one softmax thread of the tcgen05 attention kernel (experiment build -DATTN_PROF, UMGEN_LIB=...prof.so)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umgen_b200 import ops
T, S = 20, 2207
qkv = torch.randn(T * S, 2304, device="cuda").half()
y = torch.zeros(T * S, 768, dtype=torch.float16, device="cuda")
dbg = torch.zeros(128 * 128 + 128 + 128 * 48, device="cuda")
for _ in range(2):
    ops.spatial_attention(qkv, y, T, S, dbg=dbg)
torch.cuda.synchronize()
st = dbg.view(torch.int64)[:64 * 8].view(64, 8).cpu()
names = ["s_full wait", "ld", "max", "exp", "wait st", "arrive"]
tot = [0] * 6
for j in range(4, 30):
    d = [int(st[j, k + 1]) - int(st[j, k]) for k in range(6)]
    for k in range(6):
        tot[k] += d[k]
    if j < 12:
        print(f"tile {j}: " + "  ".join(f"{n}={v}" for n, v in zip(names, d)) + f"  | period {int(st[j + 1, 0]) - int(st[j, 0])}")
print("mean over tiles 4..29: " + "  ".join(f"{n}={v / 26:.0f}" for n, v in zip(names, tot)) + f"  | period {(int(st[30, 0]) - int(st[4, 0])) / 26:.0f}")
