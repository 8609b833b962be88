"""CPU stand-ins for the GPU-only parts when evaluate.py is driven without a CUDA device (BASELINE configs[0], plumbing only).  TEST INFRASTRUCTURE:
installed by tests/shims/run_evaluate.py, never imported by the product.  The engine stub returns shape- and dtype-correct tokens (the last
conditioning frame repeated), the pixel-decoder stubs return zero images of the reference's shapes; what is exercised is everything around them:
the reference's config / dataset / transforms / Lightning harness / token pickle / value decode / visualiser against this repository's
registry, model class, inference() signature and decoder classes."""
import numpy as np
import torch

CALLS = []


class StubEngine:
    def __init__(self, state_dict, cfg, sample=None, device="cpu"):
        self.cfg, self.sample, self.dev = cfg, sample, torch.device(device)
        n_param = sum(v.numel() for v in state_dict.values())
        CALLS.append(("engine", n_param))

    def inference(self, new_frames, cond_frames=1, input_cond_frames=-1, pred_task="pose_map_bbox3d_image", input_cond_tokens=None,
                  init_tokens=None, cond_on_tar=False, test_map_affine=False, max_objects=100, control_test=False, **kwargs):
        new_frames, cond_frames, input_cond_frames = int(new_frames), int(cond_frames), int(input_cond_frames)     # the harness may pass 1-element tensors
        if input_cond_frames == -1:
            input_cond_frames = cond_frames
        CALLS.append(("inference", new_frames, cond_frames, input_cond_frames, pred_task, sorted(kwargs)))
        CALLS.append(("inference_control", bool(control_test), None if init_tokens is None else {m: list(v.shape) for m, v in sorted(init_tokens.items())}))
        out = {}
        for m, width in (("pose", 3), ("map", 1024), ("bbox3d", 660), ("image", 512)):
            t = input_cond_tokens[m]
            assert t.dtype == torch.int64 and t.shape[0] == 1 and t.shape[2] == width, (m, t.dtype, tuple(t.shape))
            cond = t[0, :input_cond_frames].cpu()
            out[m] = torch.cat([cond, cond[-1:].expand(new_frames, -1)], dim=0)[None].numpy().astype(np.int64)
        return out


class _StubDecoder:
    def __init__(self, state_dict, device="cpu"):
        CALLS.append(("decoder", type(self).__name__, len(state_dict)))


class StubMapdecoder(_StubDecoder):
    def decode_maps(self, map_tokens, H=32, W=32):
        t = torch.as_tensor(map_tokens)
        return torch.zeros(t.reshape(-1, H * W).shape[0], 3, 256, 256)


class StubImagedecoder(_StubDecoder):
    def decode_images(self, image_tokens, H=16, W=32):
        t = torch.as_tensor(image_tokens)
        return torch.zeros(t.reshape(-1, H * W).shape[0], 3, 256, 512)


def install():
    import umgen_b200.engine as engine
    import umgen_b200.vq as vq
    engine.UMGenEngine = StubEngine
    vq.Mapdecoder, vq.Imagedecoder = StubMapdecoder, StubImagedecoder
    # the reference's harness hard-codes CUDA calls (model_pl.py:120-124,364-368,445-447): neutralise them on a CPU-only host
    torch.cuda.current_device = lambda: 0
    torch.Tensor.cuda = lambda self, *a, **k: self
    orig_to = torch.nn.Module.to

    def to(self, *a, **k):
        if a and (isinstance(a[0], int) or (isinstance(a[0], (str, torch.device)) and "cuda" in str(a[0]))):
            return self
        return orig_to(self, *a, **k)

    torch.nn.Module.to = to
