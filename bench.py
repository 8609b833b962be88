#!/usr/bin/env python
"""Benchmark of the UMGen next-scene decode hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one generated frame (2207 scene tokens) of the 30-frame free video-infer working point:
UMGen_Large (12/12/24/24/36/36 layers, 2.447 B params, random-init weights of that architecture),
20 conditioning frames in the sliding window, batch 1 per GPU.  The window is full from the first
generated frame on, so every step of the rollout costs the same.  With N > 1 (torchrun) every rank
decodes its own independent scene (weak scaling, no data-path collective; NCCL is used once to
broadcast the packed weights from rank 0).

`value`  scene-tokens/s with the conditioning tokens already resident in HBM (device-timed, max over ranks)
`e2e`    the same through the public `UMGenEngine.frame()` call with HOST token tensors: pinned host -> device
         copy of the conditioning window and device -> host read of the new frame inside the timed region
`roofline` the OAR decode kernel (HBM-bound): algorithmic bytes per frame / its measured time
`cpu_baseline` the oracle port (fp32, all host threads) on a bounded sample, extrapolated to a frame
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from umgen_b200 import synth  # noqa: E402
from umgen_b200.config import MODS, ModelConfig, SampleConfig  # noqa: E402

TOKENS_PER_FRAME = 2207
# SURVEY.md section 8d, per generated frame at UMGen_Large
DECODE_BYTES_PER_FRAME = 1420.4e9      # OAR weights 1122.89 GB + KV read/append 269.70 GB + heads 20.37 GB + GMLP 7.40 GB
ATTN_BYTES_PER_FRAME = 269.70e9
# dram__bytes_read.sum + dram__bytes_write.sum of one full-depth decode_cluster_kernel launch (2206 steps), ncu capture of
# tools/bench_decode.py 36 2206 2 (profiles/r1_traffic_cluster_full.csv, final build of round 1): 1432.12 GB + 2.32 GB
DECODE_TRAFFIC = {"decode_cluster_kernel": 1434.44e9}
TAR_FLOP_PER_FRAME = 187.2e12
STACK_BLOCK_EQUIV = 12 * 1.0 + 24 * (1031 / 2207) + 24 * (1693 / 2207) + 36 * 1.0     # linear-cost blocks in units of S=2207


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1368.0))), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}


def cpu_sample(threads: int):
    """Bounded CPU sample of the oracle port: one BlockTAR of the full pass (20 x 2207 tokens) and 6 full-depth
    decode steps at KV length ~1100; extrapolated to one frame by the linear layer/step counts."""
    from oracle import umgen_oracle as O
    torch.set_num_threads(threads)
    cfg = ModelConfig.large()
    P = synth.LazyParams(cfg, 0)
    with torch.no_grad():
        x = torch.randn(20, 2207, 768)
        t0 = time.time()
        O.block_tar(P, "transformer.TAR.0", x, 16)
        t_block = time.time() - t0
        n_ctx, n_steps = 1100, 6
        caches = [[torch.randn(1, n_ctx, 768), torch.randn(1, n_ctx, 768)] for _ in range(cfg.n_oar_layer)]
        xs = torch.randn(1, 1, 768)
        for i in range(cfg.n_oar_layer):            # materialise the weights outside the timed region
            O.block_oar(P, f"transformer.OAR.{i}", xs, [caches[i][0].clone(), caches[i][1].clone()], 16)
        w_head = P["transformer.head_ar_map.weight"]
        t0 = time.time()
        for _ in range(n_steps):
            h = xs
            for i in range(cfg.n_oar_layer):
                h = O.block_oar(P, f"transformer.OAR.{i}", h, caches[i], 16)
            h = O.layer_norm(h, P["transformer.ln_oar.weight"])
            torch.nn.functional.linear(h[0, -1], w_head).argmax()
        t_step = (time.time() - t0) / n_steps
    t_frame = t_block * STACK_BLOCK_EQUIV + t_step * 2206
    sample = (f"oracle port fp32: 1 BlockTAR on 20x2207 tokens ({t_block:.1f}s) + {n_steps} full-depth OAR steps at KV~{n_ctx} "
              f"({t_step * 1e3:.0f} ms/step); frame time extrapolated as {STACK_BLOCK_EQUIV:.1f} block-equivalents + 2206 steps = {t_frame:.0f}s")
    return TOKENS_PER_FRAME / t_frame, t_frame, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    sample = ""
    for _ in range(max(1, min(args.steps, 2))):
        v, t_frame, sample = cpu_sample(threads)
        vals.append(v)
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "scene-tokens/sec, 30-frame video infer (steady-state frame)", "value": v, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * TOKENS_PER_FRAME / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "UMGen_Large 30-frame free video infer, batch 1 (CPU, reference algorithm via oracle port)",
                   "cond_frames": 20, "tokens_per_frame": TOKENS_PER_FRAME},
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_s": v / TOKENS_PER_FRAME,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layers", type=int, default=0, help="debug: override every stack depth (not a valid benchmark)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="run the box_tar pass before the decode kernel instead of beside it")
    ap.add_argument("--no-lookahead", action="store_true", help="recompute the whole 20-frame window every frame (no TAR work beside the decode kernel)")
    ap.add_argument("--decode-kernel", type=int, default=0, help="0 = default (8-cluster kernel), 1 = L2-exchange kernel, 2 = 8-cluster kernel, 3 = one-cluster kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from umgen_b200 import capi
    from umgen_b200.engine import UMGenEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU path for the engine")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = ModelConfig.tiny(args.layers) if args.layers else ModelConfig.large()
    params = synth.DeviceParams(cfg, seed=0, device=dev)
    eng = UMGenEngine(params, cfg, SampleConfig.greedy(), device=dev)
    eng.check_status = False
    eng.dec.mode = args.decode_kernel
    eng.overlap = not args.no_overlap
    eng.lookahead = not args.no_lookahead
    if world > 1:       # weights come from rank 0 over NCCL (NVLink / NVSwitch); every rank then owns a replica
        from umgen_b200 import dp
        dp.broadcast_tensors(dp.engine_tensors(eng), src=0)
    T = cfg.cond_frame
    scene = synth.make_scene(seed=1 + rank, n_frames=T)
    cond_host = {m: scene[m][0].clone() for m in MODS}
    pinned = {m: cond_host[m].to(torch.int32).pin_memory() for m in MODS}
    tok_dev = {m: pinned[m].to(dev, non_blocking=True) for m in MODS}
    lib = capi.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # A step is one generated frame of a real rollout: the window slides by the frame just generated (on the device for `value`, through the
    # host for `e2e`), so consecutive steps continue each other the way evaluate.py's loop does.
    state = {"win": tok_dev}

    def slide(win, new):
        return {m: torch.cat([win[m][1:], new[m].to(torch.int32)[None]], dim=0).contiguous() for m in MODS}

    def step_device():
        new = eng.frame_device(state["win"], continues=True)
        state["win"] = slide(state["win"], new)

    host_out = torch.empty(TOKENS_PER_FRAME, dtype=torch.int64).pin_memory()

    from umgen_b200.config import CONTENT_LEN, MOD_OFFSET

    def step_e2e():
        tok = {m: pinned[m].to(dev, non_blocking=True) for m in MODS}           # H2D of the conditioning window
        eng.frame_device(tok, continues=True)
        host_out.copy_(eng.dec.out_tokens.to(torch.int64), non_blocking=True)  # D2H of the new frame
        torch.cuda.current_stream().synchronize()
        for m in MODS:                                                          # the host slides its window by the new frame
            pinned[m][:-1] = pinned[m][1:].clone()
            pinned[m][-1] = host_out[MOD_OFFSET[m] + 1: MOD_OFFSET[m] + 1 + CONTENT_LEN[m]].to(torch.int32)

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.umgen_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    dec_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    orig_decode = eng.dec.decode
    k = {"i": 0}

    def timed_decode(*a, **kw):
        e0, e1 = dec_ev[k["i"]]
        e0.record()
        r = orig_decode(*a, **kw)
        e1.record()
        k["i"] += 1
        return r

    eng.dec.decode = timed_decode
    eng.time_lookahead = True
    la_times = []
    ev[0].record()
    for _ in range(args.steps):
        step_device()
        if eng.la_events is not None:
            la_times.append(eng.la_events)
            eng.la_events = None
    ev[1].record()
    barrier()
    eng.time_lookahead = False
    la_ms = [(a.elapsed_time(b), a.elapsed_time(c)) for a, b, c in la_times]
    eng.dec.decode = orig_decode
    launches = lib.umgen_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = ev[0].elapsed_time(ev[1])
    t_decode = sum(a.elapsed_time(b) for a, b in dec_ev) / args.steps / 1e3
    status = eng.dec.status.cpu().tolist()
    if status[0] != 0:
        raise SystemExit(f"decode kernel aborted (code {status[0]})")

    # end to end through host buffers (continuing the same rollout: the host takes over the device's window)
    for m in MODS:
        pinned[m].copy_(state["win"][m])
    torch.cuda.synchronize()
    barrier()
    e2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    t_wall0 = time.time()
    e2[0].record()
    for _ in range(args.steps):
        step_e2e()
    e2[1].record()
    barrier()
    ms_e2e = max(e2[0].elapsed_time(e2[1]), 0.0)

    # TAR side alone (ego net + the three passes), sequential schedule, one extra untimed-for-the-headline frame
    overlapped = eng.overlap and eng.dec.kernel_name == "decode_cluster_kernel"
    lookahead = eng.lookahead and eng.dec.kernel_name != "decode_frame_kernel"
    eng.overlap = False
    eng.lookahead = False
    k["i"] = 0
    eng.dec.decode = timed_decode
    e3 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e3[0].record()
    step_device()
    e3[1].record()
    torch.cuda.synchronize()
    eng.dec.decode = orig_decode
    t_tar_seq = (e3[0].elapsed_time(e3[1]) - dec_ev[0][0].elapsed_time(dec_ev[0][1])) / 1e3
    eng.overlap = not args.no_overlap
    eng.lookahead = not args.no_lookahead

    t = torch.tensor([ms, ms_e2e, t_decode, t_tar_seq], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, t_decode, t_tar_seq = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, tf_peak, peak_src = peaks()
    frames = args.steps * world
    value = TOKENS_PER_FRAME * frames / (ms / 1e3)
    e2e_value = TOKENS_PER_FRAME * frames / (ms_e2e / 1e3)
    t_frame = ms / 1e3 / args.steps
    t_tar = max(t_tar_seq, 1e-9)
    scale = (cfg.n_oar_layer / 36.0)
    achieved = DECODE_BYTES_PER_FRAME * scale / t_decode / 1e9
    h2d = sum(pinned[m].numel() * 4 for m in MODS)
    line = {
        "metric": "scene-tokens/sec, 30-frame video infer (steady-state frame)", "value": value, "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp16 (fp32 accumulate / residual)", "data": "synthetic",
        "config": {"workload": "UMGen_Large 30-frame free video infer, batch 1 per GPU (BASELINE configs[1]); step = one generated frame",
                   "cond_frames": T, "tokens_per_frame": TOKENS_PER_FRAME, "layers": cfg.to_dict(), "sampling": "greedy (top-k 1)",
                   "schedule": ("look-ahead: frames 0..18 of the next window go through the TAR stacks beside the decode kernel (84 free SMs), only the "
                                "window's last frame afterwards" if lookahead else
                                "box_tar pass beside the decode kernel (second stream, 84 free SMs)" if overlapped else "sequential"),
                   "rollout": "each step continues the previous one (window slides by the generated frame)",
                   "lookahead_ms": ({"passes_beside_decode": sum(x for x, _ in la_ms) / len(la_ms), "decode_kernel": sum(y for _, y in la_ms) / len(la_ms)}
                                    if la_ms else None),
                   "l2": "per-step working set (4.9 GB of fp16 weights + 0.5 GB KV) exceeds the 126 MB L2; no explicit flush"},
        "frames_per_s": frames / (ms / 1e3),
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": TOKENS_PER_FRAME * 8},
        "gpu_launches": int(launches),
        "gpu_launches_note": "launches issued through the C ABI in the timed region; the last-frame TAR passes are replayed from CUDA graphs "
                             "(~2100 more kernel launches per frame that this counter does not see)" if lookahead else None,
        "clocks": clocks,
        "roofline": {"kernel": eng.dec.kernel_name + " (OAR decode, 2206 steps/launch)", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                     "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": (DECODE_TRAFFIC.get(eng.dec.kernel_name) if not args.layers else None), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": DECODE_BYTES_PER_FRAME * scale, "seconds_per_launch": t_decode,
                     "attention_path_bytes_per_launch": ATTN_BYTES_PER_FRAME * scale},
        "tar_roofline": {"bound": "tensor", "achieved": TAR_FLOP_PER_FRAME / t_tar / 1e12 if not args.layers else None, "peak": tf_peak,
                         "unit": "TFLOP/s", "frac": (TAR_FLOP_PER_FRAME / t_tar / 1e12 / tf_peak) if not args.layers else None,
                         "seconds_per_frame": t_tar, "note": "ego net + map/box/full TAR passes (~2100 kernel launches per frame), timed in one extra frame with "
                         "the sequential schedule (whole window recomputed); in the headline frames most of it runs beside the decode kernel" if (overlapped or lookahead) else
                         "ego net + map/box/full TAR passes (~2100 kernel launches per frame)"},
    }
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, _, sample = cpu_sample(threads)
        line["cpu_baseline"] = {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
