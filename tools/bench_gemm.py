"""TFLOP/s of the tcgen05 GEMM on the four BlockTAR shapes (M = 20 x 2207 rows) and their epilogues, CUDA events, 20 launches each."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umgen_b200 import ops
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 20 * 2207
shapes = [("c_attn  +bias       -> f16", 2304, 768, ops.EPI_BIAS_F16), ("c_proj  +bias+resid -> f32", 768, 768, ops.EPI_RESID_F32),
          ("c_fc    gelu        -> f16", 3072, 768, ops.EPI_GELU_F16), ("mlp proj     +resid -> f32", 768, 3072, ops.EPI_RESID_F32)]
tot_f, tot_t = 0.0, 0.0
for name, N, K, epi in shapes:
    a = (torch.randn(M, K, device=dev) * 0.5).half()
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
    bias = torch.randn(N, device=dev) if epi in (ops.EPI_BIAS_F16, ops.EPI_RESID_F32) and K == 768 else None
    out = torch.zeros(M, N, device=dev, dtype=torch.float16 if epi in (ops.EPI_BIAS_F16, ops.EPI_GELU_F16) else torch.float32)
    for _ in range(3):
        ops.gemm(a, w, bias, out, epi)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.gemm(a, w, bias, out, epi)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    fl = 2.0 * M * N * K
    tot_f += fl; tot_t += ms
    print(f"{name}  M={M} N={N} K={K}: {ms * 1e3:7.1f} us  {fl / ms / 1e9:6.0f} TFLOP/s")
print(f"one sub-block (4 GEMMs): {tot_t * 1e3:.1f} us  {tot_f / tot_t / 1e9:.0f} TFLOP/s")
