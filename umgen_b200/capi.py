"""ctypes binding of include/umgen.h.  Loading fails loudly if the library is missing -- there is no
fallback path (the oracle is test infrastructure and is never imported from here)."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_i64, _u64, _p, _f64 = C.c_int64, C.c_uint64, C.c_void_p, C.c_double


class UmgenDecodeArgs(C.Structure):
    _fields_ = [
        ("n_layer", _i64),
        ("oar_h", _p), ("oar_f", _p), ("ln_oar_f", _p),
        ("head_map_h", _p), ("head_bbox_h", _p), ("head_img_h", _p),
        ("map_table_f", _p), ("img_table_f", _p),
        ("be_f", _p), ("axe_f", _p), ("tske_f", _p), ("fpe_f", _p), ("box_lut_d", _p),
        ("tar_feat_f", _p), ("tar_bbox_logits_f", _p), ("pose_tok_i32", _p), ("prev_bbox_i32", _p),
        ("teacher_i32", _p), ("prefix_len", _i64), ("control_mask", _u64),
        ("top_k_map", _i64), ("top_k_bbox", _i64), ("top_k_img", _i64),
        ("sample_topp", _i64), ("top_p_map", _f64), ("top_p_bbox", _f64), ("top_p_img", _f64),
        ("temperature", _f64), ("seed", _u64), ("frame_index", _i64),
        ("merge_ar_tar", _i64), ("rule_constrain", _i64),
        ("kv_h", _p), ("scratch_f", _p),
        ("out_tokens_i32", _p), ("picks_i32", _p), ("logits_dump_f", _p), ("status_i32", _p),
        ("n_steps", _i64), ("mode", _i64), ("grid", _i64), ("debug_u64", _p), ("oar_cl_h", _p),
        ("tar_ready_i32", _p), ("tar_ready_value", _i64),
    ]


ABI_VERSION = 16
_lib = None


class UmgenError(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB


def lib():
    """The loaded libumgen_sm100.so.  With nvcc present the build is always consulted (a no-op when the source stamp matches, a rebuild when
    csrc/ or include/ changed); without nvcc (the GPU box) a library whose stamp does not match the sources is refused."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("UMGEN_LIB") or _build.LIB      # UMGEN_LIB: an experiment build (build.build_variant), tools only
    if path == _build.LIB:
        if _build.have_nvcc():
            path = _build.build(force=bool(os.environ.get("UMGEN_REBUILD")))
        elif os.path.exists(path) and not _build.stamp_matches():
            raise UmgenError(f"{path} was built from different sources than umgen_b200/csrc (stale binary) and there is no nvcc to rebuild it")
    try:
        L = C.CDLL(path)
    except OSError as e:
        raise UmgenError(f"cannot load {path}: {e}. Build it with `python -m umgen_b200.build`.") from e
    L.umgen_abi_version.restype = C.c_int
    L.umgen_last_error.restype = C.c_char_p
    L.umgen_launch_count.restype = _i64
    L.umgen_decode_scratch_floats.restype = _i64
    L.umgen_decode_frame.argtypes = [C.POINTER(UmgenDecodeArgs), _p]
    L.umgen_decode_frame.restype = C.c_int
    if hasattr(L, "umgen_decode_frames") or not os.environ.get("UMGEN_ABI_ANY"):
        L.umgen_decode_frames.argtypes = [C.POINTER(UmgenDecodeArgs), _i64, _p]
        L.umgen_decode_frames.restype = C.c_int
        L.umgen_decode_max_scenes.restype = C.c_int
    L.umgen_tar_bbox_logits.argtypes = [_p, _p, _p, _p]
    L.umgen_tar_bbox_logits.restype = C.c_int
    L.umgen_decode_cluster_capacity.restype = C.c_int
    L.umgen_pack_oar_cluster.argtypes = [_p, _p, _i64, _p]
    L.umgen_pack_oar_cluster.restype = C.c_int
    L.umgen_signal_ready.argtypes = [_p, _i64, _p]
    L.umgen_signal_ready.restype = C.c_int
    L.umgen_gemm_set_sm_limit.argtypes = [C.c_int]
    L.umgen_check_collision.argtypes = [_p, _p, _i64, _p, _p]
    L.umgen_check_collision.restype = C.c_int
    L.umgen_preload.restype = C.c_int
    if L.umgen_abi_version() != ABI_VERSION and not (os.environ.get("UMGEN_LIB") and os.environ.get("UMGEN_ABI_ANY")):      # tools may load an older experiment build
        raise UmgenError(f"ABI mismatch: library {L.umgen_abi_version()} vs binding {ABI_VERSION}; rebuild")
    _lib = L
    return L


_preloaded = set()


def preload(device) -> None:
    """Load every kernel of the library on `device` now (include/umgen.h: umgen_preload): lazy loading would otherwise stall a first launch
    that happens while the persistent decode kernel is resident."""
    import torch
    dev = torch.device(device)
    if dev.index in _preloaded:
        return
    with torch.cuda.device(dev):
        check(lib().umgen_preload(), "umgen_preload")
    _preloaded.add(dev.index)


def check(rc: int, what: str):
    if rc != 0:
        raise UmgenError(f"{what} failed ({rc}): {lib().umgen_last_error().decode()}")


EXPORTS = ["umgen_abi_version", "umgen_last_error", "umgen_launch_count", "umgen_decode_scratch_floats",
           "umgen_decode_frame", "umgen_decode_frames", "umgen_decode_max_scenes", "umgen_tar_bbox_logits", "umgen_decode_cluster_capacity", "umgen_pack_oar_cluster",
           "umgen_signal_ready", "umgen_gemm_set_sm_limit", "umgen_check_collision", "umgen_preload"]
