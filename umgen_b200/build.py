"""Build libumgen_sm100.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libumgen_sm100.so")
SOURCES = ["capi.cu", "decode.cu", "decode_cluster.cu", "decode_c16.cu", "gemm_sm100.cu", "tar.cu", "vq.cu", "exch_bench.cu", "dsmem_bench.cu", "stream_bench.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"] + os.environ.get("UMGEN_NVCC_EXTRA", "").split()


def _stamp() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(" ".join(NVCC_FLAGS + SOURCES).encode())
    return h.hexdigest()


def build_variant(tag: str, defines, verbose: bool = False) -> str:
    """Experiment builds (tools/): libumgen_sm100.<tag>.so with extra -D flags, selected at run time with UMGEN_LIB=<path>."""
    out = os.path.join(LIBDIR, f"libumgen_sm100.{tag}.so")
    return build(force=True, verbose=verbose, out=out, extra=[f"-D{d}" for d in defines], objdir=os.path.join(LIBDIR, f"obj_{tag}"))


def build(force: bool = False, verbose: bool = False, out: str = LIB, extra=(), objdir: str = LIBDIR) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(objdir, exist_ok=True)
    stamp_file = os.path.join(LIBDIR, "build.stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        log, _ = p.communicate()
        if verbose or p.returncode:
            print(log)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, "-shared", "-o", out, *objs, "-lcudart"]
    subprocess.check_call(cmd)
    if out == LIB:
        open(stamp_file, "w").write(stamp)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:          # python -m umgen_b200.build --variant TAG DEFINE[=V] ...
        k = sys.argv.index("--variant")
        print(build_variant(sys.argv[k + 1], [d for d in sys.argv[k + 2:] if not d.startswith("-")], verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
