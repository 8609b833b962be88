"""Per-object counts of the SASS mnemonics that tell a Blackwell-native kernel from a recompiled one (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor load / store), UBLKCP (1-D bulk copy), HMMA (mma.sync).
python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libdir = os.path.join(root, "umgen_b200", "lib")
pats = [("UTC*MMA", r"\bUTC\w*MMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
        ("UBLKCP", r"\bUBLKCP"), ("HMMA", r"\bHMMA"), ("MUFU.EX2", r"\bMUFU\.EX2"), ("LDGSTS", r"\bLDGSTS")]
print("# SASS mnemonic counts per object of libumgen_sm100.so (cuobjdump -sass umgen_b200/lib/<obj>.o), per kernel")
print("# kernel".ljust(64) + "".join(n.rjust(9) for n, _ in pats))
sys.path.insert(0, root)
from umgen_b200.build import SOURCES
for obj in [f.replace(".cu", ".o") for f in SOURCES]:
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(libdir, obj)], capture_output=True, text=True).stdout
    cur, counts = None, {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = [0] * len(pats)
            continue
        if cur:
            for i, (_, p) in enumerate(pats):
                if re.search(p, line):
                    counts[cur][i] += 1
    print(f"## {obj}")
    for k, v in counts.items():
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0][-60:]
        print(("  " + name).ljust(64) + "".join(str(x).rjust(9) for x in v))
