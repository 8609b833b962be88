"""Host side of the VQ pixel decoders: same surface as the reference's tokenizer wrappers
(tools/decode_map.py:110-183 Mapdecoder / Imagedecoder; tokenizer/vq_model.py:87-101 decode_code /
indices_to_quant / decode), sequencing the sm_100a kernels of csrc/vq.cu + csrc/gemm_sm100.cu exactly as
vq_modules.Decoder.forward does (tokenizer/vq_modules.py:384-415)."""
from __future__ import annotations

import math
from typing import Dict, List, Mapping, Optional

import numpy as np
import torch

from . import capi, ops

VQ_CONFIGS = {      # tokenizer/vq_model.py:150-202
    "map": dict(z_channels=16, ch=128, ch_mult=(1, 2, 2, 4), num_res_blocks=2, attn_resolutions=(16,), resolution=256, out_ch=5,
                post_quant_kernel=1, grid=(32, 32)),
    "image": dict(z_channels=256, ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolutions=(32,), resolution=512, out_ch=3,
                  post_quant_kernel=3, grid=(16, 32)),
}


def _kpad(cin: int) -> int:
    return (9 * cin + 63) // 64 * 64


def _conv3_weight(w: torch.Tensor, dev) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> fp16 [Cout, k_pad] in (ky, kx, c) order, zero padded."""
    cout, cin = w.shape[:2]
    m = w.detach().permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    out = torch.zeros(cout, _kpad(cin), dtype=torch.float16, device=dev)
    out[:, :9 * cin] = m.to(dev)
    return out


class VQDecoder:
    """One NormVQModel decoder (map or image) resident on the device."""

    def __init__(self, state_dict: Mapping[str, torch.Tensor], kind: str, device="cuda:0"):
        if not torch.cuda.is_available():
            raise capi.UmgenError("umgen_b200 needs a CUDA device (sm_100a); there is no CPU path")
        capi.lib()
        self.kind, self.cfg, self.dev = kind, VQ_CONFIGS[kind], torch.device(device)
        dev, sd, cfg = self.dev, state_dict, self.cfg
        f32 = lambda k: sd[k].detach().to(device=dev, dtype=torch.float32).contiguous()
        self.codebook = f32("quantize.embedding.weight")
        self.raw_codebook = self.codebook          # [8192, 16] as stored (indices_to_quant hands these rows out)
        if cfg["post_quant_kernel"] == 1:
            # 1x1 post_quant_conv folded into the codebook (weights-only): quant' = W . e + b
            w = f32("post_quant_conv.weight").view(cfg["z_channels"], 16)
            self.codebook = (self.codebook @ w.t() + f32("post_quant_conv.bias")).contiguous()
            self.post_quant = None
        else:
            self.post_quant = (_conv3_weight(sd["post_quant_conv.weight"], dev), f32("post_quant_conv.bias"))
        self.convs: Dict[str, tuple] = {}

        def conv3(name):
            self.convs[name] = (_conv3_weight(sd[name + ".weight"], dev), f32(name + ".bias"))

        def conv1(name):
            w = sd[name + ".weight"].detach()
            self.convs[name] = (w.view(w.shape[0], w.shape[1]).to(device=dev, dtype=torch.float16).contiguous(), f32(name + ".bias"))

        def norm(name):
            self.convs[name] = (f32(name + ".weight"), f32(name + ".bias"))

        def resblock(pre, cin, cout):
            norm(pre + ".norm1"); conv3(pre + ".conv1"); norm(pre + ".norm2"); conv3(pre + ".conv2")
            if cin != cout:
                conv1(pre + ".nin_shortcut")

        def attn(pre):
            norm(pre + ".norm")
            for n in ("q", "k", "v", "proj_out"):
                conv1(f"{pre}.{n}")

        nres = len(cfg["ch_mult"])
        block_in = cfg["ch"] * cfg["ch_mult"][-1]
        curr_res = cfg["resolution"] // 2 ** (nres - 1)
        conv3("decoder.conv_in")
        resblock("decoder.mid.block_1", block_in, block_in); attn("decoder.mid.attn_1"); resblock("decoder.mid.block_2", block_in, block_in)
        self.plan: List[tuple] = []
        for lvl in reversed(range(nres)):
            block_out = cfg["ch"] * cfg["ch_mult"][lvl]
            for ib in range(cfg["num_res_blocks"] + 1):
                pre = f"decoder.up.{lvl}.block.{ib}"
                resblock(pre, block_in, block_out)
                self.plan.append(("res", pre, block_in, block_out))
                block_in = block_out
                if curr_res in cfg["attn_resolutions"]:
                    attn(f"decoder.up.{lvl}.attn.{ib}")
                    self.plan.append(("attn", f"decoder.up.{lvl}.attn.{ib}", block_in))
            if lvl != 0:
                conv3(f"decoder.up.{lvl}.upsample.conv")
                self.plan.append(("up", f"decoder.up.{lvl}.upsample.conv", block_in))
                curr_res *= 2
        norm("decoder.norm_out")
        w = sd["decoder.conv_out.weight"].detach()
        self.conv_out_w = w.permute(0, 2, 3, 1).reshape(w.shape[0], 9, w.shape[1]).to(device=dev, dtype=torch.float32).contiguous()
        self.conv_out_b = f32("decoder.conv_out.bias")
        # the same convolution for the tensor cores: fp16 [128, 9 * C] / fp32 [128], rows >= out_ch zero (umgen_conv3x3_nchw_f32)
        self.conv_out_w16 = torch.zeros(128, 9 * w.shape[1], dtype=torch.float16, device=dev)
        self.conv_out_w16[:w.shape[0]] = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(dev)
        self.conv_out_b128 = torch.zeros(128, dtype=torch.float32, device=dev)
        self.conv_out_b128[:w.shape[0]] = self.conv_out_b
        self.c_last = block_in
        self.stats = torch.empty(64 * 32 * 2, dtype=torch.float32, device=dev)
        self._bufs: Dict[str, torch.Tensor] = {}
        self._gn_scratch: Dict[tuple, torch.Tensor] = {}
        # implicit_conv: 3x3 convolutions over >= 64 channels fetch their taps by TMA (umgen_conv3x3_f16) instead of through an im2col matrix;
        # conv_out_tc: conv_out on the tensor cores too (fp16 weights, zero-padded to 128 output channels) instead of the warp-per-pixel kernel;
        # slab_groupnorm: coalesced GroupNorm statistics; use_graph: decode_code replays one CUDA graph per batch size (~200 launches of kernels
        # that run for a few microseconds each).  All four on by default; the switches exist for the cross-checks in tests/test_vq_gpu.py.
        self.implicit_conv, self.conv_out_tc, self.slab_groupnorm, self.use_graph = True, True, True, True
        self._graphs: Dict[int, tuple] = {}

    def _buf(self, name: str, numel: int, dtype=torch.float16) -> torch.Tensor:
        t = self._bufs.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(numel, dtype=dtype, device=self.dev)
            self._bufs[name] = t
        return t[:numel]

    # ---- building blocks on channels-last activations x [B*H*W, C] --------------------------------------------
    def _conv3(self, name, x, B, H, W, cin, *, upsample=False, resid=None, weights=None):
        w, b = weights if weights is not None else self.convs[name]
        cout, kp = w.shape
        rows = B * H * W
        out = torch.empty(rows, cout, dtype=torch.float16, device=self.dev)
        epi = ops.EPI_RESID_F16 if resid is not None else ops.EPI_BIAS_F16
        if self.implicit_conv and kp == 9 * cin and ops.conv3x3_supported(H, W, cin, cout):
            if upsample:                      # Upsample.forward (vq_modules.py:34-40): nearest 2x, then the convolution reads the upsampled activation
                x = ops.upsample2x(x, self._buf("upsampled", rows * cin), B, H // 2, W // 2, cin)
            return ops.conv3x3(x, w, b, out, B, H, W, cin, epi, resid)
        a = self._buf("im2col", rows * kp).view(rows, kp)
        ops.im2col3x3(x, a, B, H, W, cin, kp, upsample)
        ops.gemm(a, w, b, out, epi, resid)
        return out

    def _conv1(self, name, x, *, resid=None):
        w, b = self.convs[name]
        out = torch.empty(x.shape[0], w.shape[0], dtype=torch.float16, device=self.dev)
        ops.gemm(x, w, b, out, ops.EPI_RESID_F16 if resid is not None else ops.EPI_BIAS_F16, resid)
        return out

    def _gn(self, name, x, B, HW, Cc, swish):
        g, b = self.convs[name]
        y = torch.empty_like(x)
        if self.slab_groupnorm and Cc in (128, 256, 512):
            sc = self._gn_scratch.get((B, HW))
            if sc is None:                    # mean / rstd, per-slab partial sums, tickets (zero once; the kernel leaves them zero)
                sc = self._gn_scratch[(B, HW)] = torch.zeros(ops.groupnorm_scratch_floats(B, HW), dtype=torch.float32, device=self.dev)
            return ops.groupnorm_slab(x, g, b, y, sc, B, HW, Cc, swish)
        ops.groupnorm(x, g, b, y, self.stats, B, HW, Cc, swish)
        return y

    def _resblock(self, pre, x, B, H, W, cin, cout):      # ResnetBlock.forward (vq_modules.py:108-127)
        h = self._gn(pre + ".norm1", x, B, H * W, cin, True)
        h = self._conv3(pre + ".conv1", h, B, H, W, cin)
        h = self._gn(pre + ".norm2", h, B, H * W, cout, True)
        skip = x if cin == cout else self._conv1(pre + ".nin_shortcut", x)
        return self._conv3(pre + ".conv2", h, B, H, W, cout, resid=skip)

    def _attn(self, pre, x, B, HW, Cc):                    # AttnBlock.forward (vq_modules.py:149-176)
        h = self._gn(pre + ".norm", x, B, HW, Cc, False)
        q, k, v = (self._conv1(f"{pre}.{n}", h) for n in ("q", "k", "v"))
        o = torch.empty(B * HW, Cc, dtype=torch.float16, device=self.dev)
        s = self._buf("scores", HW * HW, torch.float32).view(HW, HW)
        p = self._buf("probs", HW * HW).view(HW, HW)
        vt = self._buf("vt", Cc * HW).view(Cc, HW)
        for b in range(B):
            sl = slice(b * HW, (b + 1) * HW)
            ops.gemm(q[sl], k[sl], None, s, ops.EPI_STORE_F32)             # w_[i, j] = sum_c q[i, c] k[j, c]
            ops.softmax_rows(s, p, Cc ** -0.5)
            ops.transpose_f16(v[sl], vt)
            ops.gemm(p, vt, None, o[sl], ops.EPI_BIAS_F16)                 # h_[i, c] = sum_j p[i, j] v[j, c]
        return self._conv1(pre + ".proj_out", o, resid=x)

    # ---- vq_model.py:87-101 ---------------------------------------------------------------------------------------
    def decode_code(self, code: torch.Tensor) -> torch.Tensor:
        """code: int [B, h, w] token grid -> fp32 [B, out_ch, 8h | 16h, 8w | 16w] (NormVQModel.decode_code)."""
        B, H, W = code.shape
        idx = code.to(device=self.dev, dtype=torch.int32).contiguous()
        with torch.cuda.device(self.dev):
            if not self.use_graph:
                return self._decode(idx)
            ent = self._graphs.get((B, H, W))
            if ent is None:
                # first call at this batch size: one eager pass (loads the kernels, fills the tensor-map cache, sizes the shared buffers), then the capture
                self._decode(idx)
                static_idx = idx.clone()
                torch.cuda.synchronize(self.dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    static_out = self._decode(static_idx)
                ent = self._graphs[(B, H, W)] = (g, static_idx, static_out)
            g, static_idx, static_out = ent
            static_idx.copy_(idx)
            g.replay()
            return static_out.clone()

    def _decode(self, idx: torch.Tensor) -> torch.Tensor:
        """Decoder.forward (vq_modules.py:384-415) on an int32 token grid resident on the device."""
        cfg, dev = self.cfg, self.dev
        B, H, W = idx.shape
        x = torch.empty(B * H * W, 16 if self.post_quant is not None else cfg["z_channels"], dtype=torch.float16, device=dev)
        ops.vq_gather(idx.view(-1), self.codebook, x)
        cin = x.shape[1]
        if self.post_quant is not None:
            x = self._conv3(None, x, B, H, W, 16, weights=self.post_quant)
            cin = cfg["z_channels"]
        x = self._conv3("decoder.conv_in", x, B, H, W, cin)
        c = cfg["ch"] * cfg["ch_mult"][-1]
        x = self._resblock("decoder.mid.block_1", x, B, H, W, c, c)
        x = self._attn("decoder.mid.attn_1", x, B, H * W, c)
        x = self._resblock("decoder.mid.block_2", x, B, H, W, c, c)
        for step in self.plan:
            if step[0] == "res":
                x = self._resblock(step[1], x, B, H, W, step[2], step[3])
            elif step[0] == "attn":
                x = self._attn(step[1], x, B, H * W, step[2])
            else:
                H, W = 2 * H, 2 * W
                x = self._conv3(step[1], x, B, H, W, step[2], upsample=True)
        x = self._gn("decoder.norm_out", x, B, H * W, self.c_last, True)
        out = torch.empty(B, cfg["out_ch"], H, W, dtype=torch.float32, device=dev)
        if self.conv_out_tc and ops.conv3x3_supported(H, W, self.c_last, 128):
            return ops.conv3x3_nchw(x, self.conv_out_w16, self.conv_out_b128, out, B, H, W, self.c_last, cfg["out_ch"])
        ops.conv_out3x3(x, self.conv_out_w, self.conv_out_b, out, B, H, W, self.c_last, cfg["out_ch"])
        return out

    def decode_tokens(self, tokens, chunk: int = 4) -> torch.Tensor:
        """tokens: int [T, h*w] (or [1, T, h*w]) -> fp32 [T, out_ch, H, W], decoded `chunk` frames at a time."""
        t = torch.as_tensor(tokens)
        if t.dim() == 3:
            t = t[0]
        if t.dim() == 1:
            t = t[None]
        h, w = self.cfg["grid"]
        return torch.cat([self.decode_code(t[i:i + chunk].reshape(-1, h, w)) for i in range(0, t.shape[0], chunk)], dim=0)


def rgb_weights(n_in: int = 5, seed: int = 0) -> torch.Tensor:
    """The fixed projection of to_rgb (tools/decode_map.py:25-27): torch.manual_seed(seed); randn(3, C, 1, 1).
    A private generator yields the same numbers without resetting the global RNG."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(3, n_in, 1, 1, generator=g).view(3, n_in)


class Mapdecoder:
    """tools/decode_map.py:110-147."""

    def __init__(self, state_dict, device="cuda:0"):
        self.map_autoencoder = VQDecoder(state_dict, "map", device)
        self._w = rgb_weights(5).to(self.map_autoencoder.dev)
        self._mm = torch.zeros(2, dtype=torch.int32, device=self.map_autoencoder.dev)

    def decode_maps(self, map_tokens, H=32, W=32) -> torch.Tensor:
        t = torch.as_tensor(map_tokens)
        if t.dim() == 3:
            t = t[0]
        outs = []
        for i in range(math.ceil(t.shape[0] / 20)):           # min-max is taken per 20-frame chunk (decode_map.py:137-142)
            rec = self.map_autoencoder.decode_tokens(t[i * 20:(i + 1) * 20])
            B, Cc, Hh, Ww = rec.shape
            rgb = torch.empty(B, 3, Hh, Ww, dtype=torch.float32, device=rec.device)
            ops.to_rgb(rec, self._w, rgb, self._mm, B, Cc, Hh * Ww)
            outs.append(rgb)
        return torch.cat(outs, dim=0)


class Imagedecoder:
    """tools/decode_map.py:150-183."""

    def __init__(self, state_dict, device="cuda:0"):
        self.img_autoencoder = VQDecoder(state_dict, "image", device)

    def decode_images(self, image_tokens, H=16, W=32) -> torch.Tensor:
        return self.img_autoencoder.decode_tokens(image_tokens)
