"""CPU-side checks of the C-ABI boundary: the library builds, loads and exports every symbol that
include/umgen.h declares; the ctypes mirror of UmgenDecodeArgs has the header's field order."""
import ctypes
import os
import re

import pytest

from umgen_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    return open(os.path.join(ROOT, "include", "umgen.h")).read()


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = set(re.findall(r"\b(umgen_[a-z0-9_]+)\s*\(", _header()))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/umgen.h but not exported"
    assert set(capi.EXPORTS) <= declared
    assert capi.lib().umgen_abi_version() == capi.ABI_VERSION
    m = re.search(r"#define UMGEN_ABI_VERSION (\d+)", _header())
    assert int(m.group(1)) == capi.ABI_VERSION


def test_decode_args_mirror_matches_header():
    body = re.search(r"typedef struct UmgenDecodeArgs \{(.*?)\} UmgenDecodeArgs;", _header(), re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        typ_and_names = decl.replace("const ", "").replace("*", " ")
        parts = typ_and_names.split(None, 1)
        for n in parts[1].split(","):
            names.append(n.strip())
    assert names == [f[0] for f in capi.UmgenDecodeArgs._fields_]
    assert ctypes.sizeof(capi.UmgenDecodeArgs) == 8 * len(names)


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from umgen_b200.config import ModelConfig
    from umgen_b200.decoder import FrameDecoder
    with pytest.raises(capi.UmgenError):
        FrameDecoder({}, ModelConfig.tiny(1))
