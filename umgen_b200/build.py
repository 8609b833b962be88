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
SOURCES = ["capi.cu", "preload.cu", "decode.cu", "decode_cluster.cu", "gemm_sm100.cu", "attn_sm100.cu", "tar.cu", "vq.cu"]
# Microbenchmarks and the one-cluster decode study (profiles/r1_*_microbench.txt, r1_c16_study.txt): a separate library for tools/, never loaded by the product
TOOLS_DIR = os.path.join(HERE, "..", "tools", "csrc")
TOOLS_LIB = os.path.join(LIBDIR, "libumgen_tools.so")
TOOLS_SOURCES = ["exch_bench.cu", "dsmem_bench.cu", "stream_bench.cu", "decode_c16.cu", "lat_bench.cu"]
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


def have_nvcc() -> bool:
    return os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"))


def stamp_matches() -> bool:
    f = os.path.join(LIBDIR, "build.stamp")
    return os.path.exists(f) and open(f).read() == _stamp()


def build_variant(tag: str, defines, verbose: bool = False, force: bool = True) -> str:
    """Experiment / profiling builds: libumgen_sm100.<tag>.so with extra -D flags, selected at run time with UMGEN_LIB=<path>."""
    out = os.path.join(LIBDIR, f"libumgen_sm100.{tag}.so")
    stamp_file = os.path.join(LIBDIR, f"build.{tag}.stamp")
    stamp = _stamp() + " " + " ".join(defines)
    if not force and os.path.exists(out) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return out
    build(force=True, verbose=verbose, out=out, extra=[f"-D{d}" for d in defines], objdir=os.path.join(LIBDIR, f"obj_{tag}"))
    open(stamp_file, "w").write(stamp)
    return out


def build_profiling() -> str:
    """The profiling build bench.py's attention-path roofline reads (same sources, -DUMGEN_DECODE_PROFILE=1: per-phase clocks in the decode kernel)."""
    return build_variant("prof1", ["UMGEN_DECODE_PROFILE=1"], force=False)


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


def build_tools(verbose: bool = False) -> str:
    """libumgen_tools.so: tools/csrc/*.cu + capi.cu (error plumbing)."""
    os.makedirs(os.path.join(LIBDIR, "obj_tools"), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    for d, src in [(CSRC, "capi.cu")] + [(TOOLS_DIR, f) for f in TOOLS_SOURCES]:
        obj = os.path.join(LIBDIR, "obj_tools", src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(d, src), "-o", obj] + (["-Xptxas=-v"] if verbose else [])
        subprocess.check_call(cmd)
        objs.append(obj)
    subprocess.check_call([nvcc, "-shared", "-o", TOOLS_LIB, *objs, "-lcudart"])
    return TOOLS_LIB


if __name__ == "__main__":
    if "--tools" in sys.argv:
        print(build_tools(verbose="-v" in sys.argv))
    elif "--variant" in sys.argv:          # python -m umgen_b200.build --variant TAG DEFINE[=V] ...
        k = sys.argv.index("--variant")
        print(build_variant(sys.argv[k + 1], [d for d in sys.argv[k + 2:] if not d.startswith("-")], verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
