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


def test_conv3x3_geometry_rules_and_argument_checks():
    """Host logic of the implicit-GEMM convolution (include/umgen.h: umgen_conv3x3_f16): which images tile into 128-pixel boxes, and that the
    C entry points refuse bad arguments before touching the device (no GPU needed: the checks precede every CUDA call)."""
    import ctypes as C
    from umgen_b200 import ops
    ok = ops.conv3x3_supported
    # every resolution of the two decoders (tokenizer/vq_model.py:150-202): map 32..256 square, image 16x32 .. 256x512
    for H, W in [(32, 32), (64, 64), (128, 128), (256, 256), (16, 32), (32, 64), (64, 128), (128, 256), (256, 512)]:
        for cin, cout in [(128, 128), (256, 128), (256, 256), (512, 256), (512, 512)]:
            assert ok(H, W, cin, cout), (H, W, cin, cout)
    assert not ok(32, 32, 16, 512)          # conv_in of the map decoder: 16 channels, below one 64-channel K block -> im2col path
    assert not ok(32, 32, 96, 128) and not ok(32, 32, 128, 64) and not ok(32, 32, 128, 192)
    assert not ok(6, 32, 64, 128)           # 4-row boxes do not tile 6 rows
    assert not ok(16, 48, 64, 128) and not ok(8, 192, 64, 128) and not ok(64, 4, 64, 128)
    L = capi.lib()
    L.umgen_conv3x3_f16.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    L.umgen_conv3x3_nchw_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.umgen_groupnorm_scratch_floats.argtypes = [C.c_int64, C.c_int64]
    L.umgen_groupnorm_scratch_floats.restype = C.c_int64
    fake = 0x1000      # an aligned non-null address that is never dereferenced: every call below fails its argument checks first
    assert L.umgen_conv3x3_f16(None, 1, 32, 32, 64, fake, None, fake, None, 128, 0, None) == -1 and b"null" in L.umgen_last_error()
    assert L.umgen_conv3x3_f16(fake, 1, 32, 32, 64, fake, None, fake, None, 128, 1, None) == -1 and b"epilogue" in L.umgen_last_error()      # GELU is not a conv epilogue
    assert L.umgen_conv3x3_f16(fake, 1, 32, 32, 64, fake, None, fake, None, 128, 4, None) == -1 and b"resid" in L.umgen_last_error()
    assert L.umgen_conv3x3_f16(fake, 1, 32, 32, 16, fake, None, fake, None, 128, 0, None) == -1 and b"Cin" in L.umgen_last_error()
    assert L.umgen_conv3x3_f16(fake, 1, 6, 32, 64, fake, None, fake, None, 128, 0, None) == -1 and b"tile" in L.umgen_last_error()
    assert L.umgen_conv3x3_f16(fake + 4, 1, 32, 32, 64, fake, None, fake, None, 128, 0, None) == -1 and b"aligned" in L.umgen_last_error()
    assert L.umgen_conv3x3_nchw_f32(fake, 1, 32, 32, 64, fake, fake, fake, 33, None) == -1 and b"n_out" in L.umgen_last_error()
    # scratch of the slab GroupNorm: mean / rstd + one partial (sum, sum of squares) per slab and group + one ticket per image
    assert L.umgen_groupnorm_scratch_floats(4, 131072) == 4 * 64 + 4 * 128 * 64 + 4
    assert L.umgen_groupnorm_scratch_floats(2, 512) == 2 * 64 + 2 * 8 * 64 + 2
    assert L.umgen_groupnorm_scratch_floats(1, 1000) == 64 + 16 * 64 + 1
