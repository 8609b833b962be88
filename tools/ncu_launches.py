"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total time, share.
Usage: python tools/ncu_launches.py launches.csv [title]"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= iv or not r[iv]:
        continue
    v = float(r[iv].replace(",", ""))
    if v != v:          # nan: ncu could not time the launch (the cooperative cluster launch of the decode kernel without UMGEN_DECODE_NO_COOP=1)
        agg[re.sub(r"\(.*", "", r[ik]) + "  [not timed by ncu]"][0] += 1
        continue
    ms = v / 1e6 if r[iu].startswith("ns") or r[iu] == "nsecond" else (v / 1e3 if r[iu].startswith("us") else v)
    name = re.sub(r"\(.*", "", r[ik])
    if not ("umgen" in name or "decode_cluster_kernel" in name or "decode_frame_kernel" in name or name.startswith(("cl::", "fa::", "attn::", "gemm::", "c16::", "void gemm", "void umgen", "void fa", "void attn"))):
        continue          # torch kernels of the synthetic-weight generator etc.
    agg[name][0] += 1
    agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}")
print(f"# library kernels only; cold-cache serialised launch times: compare SHARES, not absolutes; total {tot:.1f} ms")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:10.2f} ms  {100 * ms / tot:5.1f}%  n={n:6d}  {k}")
