"""Thin torch-tensor wrappers over the C-ABI entry points of the TAR kernels (include/umgen.h).
Every wrapper launches on the current torch CUDA stream and raises UmgenError on failure."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import capi

EPI_BIAS_F16, EPI_GELU_F16, EPI_RESID_F32, EPI_STORE_F32, EPI_RESID_F16 = 0, 1, 2, 3, 4
_i64, _p = C.c_int64, C.c_void_p


class UmgenEmbedArgs(C.Structure):
    _fields_ = [(n, _p) for n in ("pose_i32", "map_i32", "bbox_i32", "image_i32", "fpe_f", "img_table_f", "be_f", "axe_f",
                                  "spe_f", "tpe_f", "spatial_f", "map_feat_f", "map_warped_f", "out_f")] + [("T", _i64), ("n_mods", _i64), ("t_offset", _i64)]


_bound = False


def _lib():
    global _bound
    L = capi.lib()
    if not _bound:
        L.umgen_gemm_f16.argtypes = [_p, _i64, _p, _p, _p, _i64, _i64, _i64, _i64, C.c_int, _p]
        L.umgen_gemm_f16_ex.argtypes = [_p, _i64, _p, _p, _p, _i64, _p, _i64, _i64, _i64, _i64, C.c_int, _p]
        L.umgen_vq_gather.argtypes = [_p, _p, _p, _i64, _p]
        L.umgen_im2col3x3.argtypes = [_p, _p, _i64, _i64, _i64, _i64, _i64, C.c_int, _p]
        L.umgen_groupnorm_nhwc.argtypes = [_p, _p, _p, _p, _p, _i64, _i64, _i64, C.c_int, _p]
        L.umgen_groupnorm_nhwc_slab.argtypes = [_p, _p, _p, _p, _p, _i64, _i64, _i64, C.c_int, _p]
        L.umgen_groupnorm_scratch_floats.argtypes = [_i64, _i64]
        L.umgen_groupnorm_scratch_floats.restype = _i64
        L.umgen_conv3x3_f16.argtypes = [_p, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _i64, C.c_int, _p]
        L.umgen_conv3x3_nchw_f32.argtypes = [_p, _i64, _i64, _i64, _i64, _p, _p, _p, _i64, _p]
        L.umgen_upsample2x_nhwc.argtypes = [_p, _p, _i64, _i64, _i64, _i64, _p]
        L.umgen_softmax_rows.argtypes = [_p, _p, _i64, _i64, C.c_double, _p]
        L.umgen_transpose_f16.argtypes = [_p, _p, _i64, _i64, _p]
        L.umgen_conv_out3x3.argtypes = [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _p]
        L.umgen_to_rgb.argtypes = [_p, _p, _p, _p, _i64, _i64, _i64, _p]
        L.umgen_layernorm.argtypes = [_p, _p, _p, _i64, C.c_int, _p]
        L.umgen_cast_f16.argtypes = [_p, _p, _i64, _p]
        L.umgen_map_feature.argtypes = [_p, _p, _p, _p, _i64, _p]
        L.umgen_map_warp.argtypes = [_p, _p, _p, _p, _i64, _p]
        L.umgen_embed_sequence.argtypes = [C.POINTER(UmgenEmbedArgs), _p]
        L.umgen_small_attention.argtypes = [_p, _p, _i64, _i64, _i64, _i64, C.c_int, _p]
        L.umgen_small_attention_from.argtypes = [_p, _p, _i64, _i64, _i64, _i64, C.c_int, _i64, _p]
        L.umgen_spatial_attention.argtypes = [_p, _p, _i64, _i64, _p]
        L.umgen_spatial_attention_tc.argtypes = [_p, _p, _i64, _i64, _p, _p]
        L.umgen_cross_attention.argtypes = [_p, _p, _p, _p, _i64, _i64, _p]
        L.umgen_sample_rows.argtypes = [_p, _i64, _i64, _i64, C.c_double, C.c_double, C.c_uint64, _i64, _p, _p]
        L.umgen_assemble_tar_feat.argtypes = [_p, _p, _p, _p, _p, _i64, _i64, _p]
        _bound = True
    return L


def _s():
    """Stream of torch's CURRENT device: the engine enters `torch.cuda.device(engine.dev)` around every public call, so this is the engine's device
    whatever the caller's current device was."""
    return torch.cuda.current_stream().cuda_stream


def _dp(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, epilogue: int,
         resid: Optional[torch.Tensor] = None):
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T (+bias) (+resid)).  a, w fp16; out fp16 (EPI 0/1/4) or fp32 (EPI 2/3)."""
    M, K = a.shape
    N = w.shape[0]
    assert a.dtype == torch.float16 and w.dtype == torch.float16 and w.shape[1] == K and a.stride(1) == 1 and w.is_contiguous()
    assert out.shape == (M, N) and out.stride(1) == 1
    assert out.dtype == (torch.float16 if epilogue in (EPI_BIAS_F16, EPI_GELU_F16, EPI_RESID_F16) else torch.float32)
    capi.check(_lib().umgen_gemm_f16_ex(a.data_ptr(), a.stride(0), w.data_ptr(), _dp(bias), out.data_ptr(), out.stride(0), _dp(resid),
                                        0 if resid is None else resid.stride(0), M, N, K, epilogue, _s()), "umgen_gemm_f16_ex")
    return out


def vq_gather(idx, table, out):
    capi.check(_lib().umgen_vq_gather(idx.data_ptr(), table.data_ptr(), out.data_ptr(), idx.numel(), _s()), "umgen_vq_gather")
    return out


def im2col3x3(x, a, B, H, W, Cin, k_pad, upsample):
    capi.check(_lib().umgen_im2col3x3(x.data_ptr(), a.data_ptr(), B, H, W, Cin, k_pad, int(upsample), _s()), "umgen_im2col3x3")
    return a


def conv3x3_supported(H: int, W: int, Cin: int, Cout: int) -> bool:
    """Geometry umgen_conv3x3_f16 takes (include/umgen.h): 64-channel K blocks, 128-pixel boxes of one image."""
    bw = min(W, 128)
    return Cin >= 64 and Cin % 64 == 0 and Cout >= 128 and Cout % 128 == 0 and W >= 8 and 128 % bw == 0 and W % bw == 0 and H % (128 // bw) == 0


def conv3x3(x, w, bias, out, B, H, W, Cin, epilogue: int = EPI_BIAS_F16, resid: Optional[torch.Tensor] = None):
    """out[B*H*W, Cout] = conv3x3(x[B,H,W,Cin]; w[Cout, 9*Cin]) + bias (+ resid): implicit GEMM, csrc/gemm_sm100.cu."""
    Cout = w.shape[0]
    assert x.dtype == torch.float16 and w.dtype == torch.float16 and x.is_contiguous() and w.is_contiguous() and w.shape[1] == 9 * Cin
    assert x.numel() == B * H * W * Cin and out.dtype == torch.float16 and out.is_contiguous() and out.numel() == B * H * W * Cout
    assert resid is None or (resid.dtype == torch.float16 and resid.is_contiguous() and resid.numel() == out.numel())
    capi.check(_lib().umgen_conv3x3_f16(x.data_ptr(), B, H, W, Cin, w.data_ptr(), _dp(bias), out.data_ptr(), _dp(resid), Cout, epilogue, _s()),
               "umgen_conv3x3_f16")
    return out


def conv3x3_nchw(x, w, bias, out, B, H, W, Cin, n_out):
    """out[B, n_out, H, W] fp32 = conv3x3(x[B,H,W,Cin]; w[128, 9*Cin], rows >= n_out zero) + bias: conv_out of the VQ decoders."""
    assert x.dtype == torch.float16 and w.dtype == torch.float16 and x.is_contiguous() and w.is_contiguous() and w.shape == (128, 9 * Cin)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == B * n_out * H * W and bias.numel() == 128
    capi.check(_lib().umgen_conv3x3_nchw_f32(x.data_ptr(), B, H, W, Cin, w.data_ptr(), bias.data_ptr(), out.data_ptr(), n_out, _s()), "umgen_conv3x3_nchw_f32")
    return out


def upsample2x(x, out, B, H, W, Cc):
    capi.check(_lib().umgen_upsample2x_nhwc(x.data_ptr(), out.data_ptr(), B, H, W, Cc, _s()), "umgen_upsample2x_nhwc")
    return out


def groupnorm_scratch_floats(B: int, HW: int) -> int:
    return int(_lib().umgen_groupnorm_scratch_floats(B, HW))


def groupnorm_slab(x, gamma, beta, y, scratch, B, HW, Cc, swish):
    """GroupNorm(32) (+ swish) with the coalesced statistics pass; `scratch` as umgen_groupnorm_nhwc_slab wants it (tickets zeroed by the caller)."""
    capi.check(_lib().umgen_groupnorm_nhwc_slab(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), scratch.data_ptr(), B, HW, Cc, int(swish),
                                                _s()), "umgen_groupnorm_nhwc_slab")
    return y


def groupnorm(x, gamma, beta, y, stats, B, HW, Cc, swish):
    capi.check(_lib().umgen_groupnorm_nhwc(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), stats.data_ptr(), B, HW, Cc, int(swish), _s()),
               "umgen_groupnorm_nhwc")
    return y


def softmax_rows(s, p, scale):
    capi.check(_lib().umgen_softmax_rows(s.data_ptr(), p.data_ptr(), s.shape[0], s.shape[1], float(scale), _s()), "umgen_softmax_rows")
    return p


def transpose_f16(x, out):
    capi.check(_lib().umgen_transpose_f16(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], _s()), "umgen_transpose_f16")
    return out


def conv_out3x3(x, w, bias, out, B, H, W, Cin, Cout):
    capi.check(_lib().umgen_conv_out3x3(x.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), B, H, W, Cin, Cout, _s()), "umgen_conv_out3x3")
    return out


def to_rgb(x, w, out, mm, B, Cin, HW):
    capi.check(_lib().umgen_to_rgb(x.data_ptr(), w.data_ptr(), out.data_ptr(), mm.data_ptr(), B, Cin, HW, _s()), "umgen_to_rgb")
    return out


def layernorm(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor):
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[-1] == 768 and out.is_contiguous()
    rows = x.numel() // 768
    capi.check(_lib().umgen_layernorm(x.data_ptr(), w.data_ptr(), out.data_ptr(), rows, int(out.dtype == torch.float16), _s()), "umgen_layernorm")
    return out


def cast_f16(x: torch.Tensor, out: torch.Tensor):
    capi.check(_lib().umgen_cast_f16(x.data_ptr(), out.data_ptr(), x.numel(), _s()), "umgen_cast_f16")
    return out


def map_feature(tok: torch.Tensor, table: torch.Tensor, grid_pos: Optional[torch.Tensor], out: torch.Tensor):
    capi.check(_lib().umgen_map_feature(tok.data_ptr(), table.data_ptr(), _dp(grid_pos), out.data_ptr(), tok.numel(), _s()), "umgen_map_feature")
    return out


def map_warp(feat: torch.Tensor, pose_tok: torch.Tensor, pose_lut: torch.Tensor, out: torch.Tensor):
    capi.check(_lib().umgen_map_warp(feat.data_ptr(), pose_tok.data_ptr(), pose_lut.data_ptr(), out.data_ptr(), feat.shape[0], _s()), "umgen_map_warp")
    return out


def embed_sequence(tokens, tables, map_feat, map_warped, out, n_mods: int, t_offset: int = 0):
    a = UmgenEmbedArgs()
    a.pose_i32, a.map_i32, a.bbox_i32, a.image_i32 = (tokens[m].data_ptr() for m in ("pose", "map", "bbox3d", "image"))
    for k in ("fpe_f", "img_table_f", "be_f", "axe_f", "spe_f", "tpe_f", "spatial_f"):
        setattr(a, k, tables[k].data_ptr())
    a.map_feat_f, a.map_warped_f, a.out_f = map_feat.data_ptr(), _dp(map_warped), out.data_ptr()
    a.T, a.n_mods, a.t_offset = tokens["pose"].shape[0], n_mods, t_offset
    capi.check(_lib().umgen_embed_sequence(C.byref(a), _s()), "umgen_embed_sequence")
    return out


def small_attention(qkv, y, n_groups, n_tok, group_stride, tok_stride, causal, q0: int = 0):
    capi.check(_lib().umgen_small_attention_from(qkv.data_ptr(), y.data_ptr(), n_groups, n_tok, group_stride, tok_stride, int(causal), q0, _s()),
               "umgen_small_attention_from")
    return y


def spatial_attention(qkv, y, T, S, dbg=None):
    """tcgen05 / TMEM flash attention (csrc/attn_sm100.cu)."""
    capi.check(_lib().umgen_spatial_attention_tc(qkv.data_ptr(), y.data_ptr(), T, S, _dp(dbg), _s()), "umgen_spatial_attention_tc")
    return y


def spatial_attention_mma(qkv, y, T, S):
    """The mma.sync kernel of round 1 (csrc/tar.cu), kept as a cross-check."""
    capi.check(_lib().umgen_spatial_attention(qkv.data_ptr(), y.data_ptr(), T, S, _s()), "umgen_spatial_attention")
    return y


def cross_attention(q, k, v, y):
    capi.check(_lib().umgen_cross_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), y.data_ptr(), q.shape[0], k.shape[0], _s()),
               "umgen_cross_attention")
    return y


def sample_rows(logits, top_k, temperature, seed, frame_index, out, top_p: float = 0.0):
    """top_p > 0: nucleus sampling (sample_top_p, UMGen.py:915-965); else top-k (UMGen.py:899-913)."""
    rows, V = logits.shape
    capi.check(_lib().umgen_sample_rows(logits.data_ptr(), rows, V, int(top_k), float(top_p), float(temperature), int(seed), int(frame_index),
                                        out.data_ptr(), _s()), "umgen_sample_rows")
    return out


def assemble_tar_feat(f_all, f_map, f_box, warped_last, out, row0: int = 0, row1: int = 2207):
    capi.check(_lib().umgen_assemble_tar_feat(f_all.data_ptr(), f_map.data_ptr(), f_box.data_ptr(), warped_last.data_ptr(), out.data_ptr(),
                                              row0, row1, _s()), "umgen_assemble_tar_feat")
    return out
