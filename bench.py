#!/usr/bin/env python
"""Benchmark of the UMGen next-scene decode hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-gpu] [--workload video|control|long] [--scenes-per-gpu B]

A *step* is one generated frame (2207 scene tokens) of the 30-frame free video-infer working point (BASELINE configs[1]):
UMGen_Large (12/12/24/24/36/36 layers, 2.447 B params, random-init weights of that architecture), 20 conditioning frames
in the sliding window, batch 1 per GPU, greedy.  The window is full from the first generated frame on, so every step of
the rollout costs the same.  With N > 1 (torchrun) every rank decodes its own independent scene (weak scaling, no
data-path collective; NCCL is used once to broadcast the packed weights from rank 0).

`value`     scene-tokens/s with the conditioning tokens already resident in HBM (device-timed, max over ranks)
`e2e`       the same through the reference-facing plugin call `projects.models.UMGen.UMGen.inference(...)` (the call
            tools/model_pl.py:237-239 makes): CPU LongTensors in, numpy int64 out, one contiguous K-frame rollout; per
            frame the host -> device copy of the conditioning window, the host-side window check of the look-ahead
            schedule, the decode status read-back and the device -> host copy of the new frame are inside the timed region
`roofline`  the OAR decode kernel (HBM-bound): algorithmic bytes per frame / its measured time; `attention`
            inside it: the KV-cache path alone, timed by the profiling build of the same kernel
`cpu_baseline` / `--impl reference`: ONE real frame of the CPU oracle (the reference algorithm restated in fp32, all host
            threads) at full depth on the same weights and scene -- measured, not extrapolated -- and `parity_fulldepth`:
            the GPU engine teacher-forced on that frame's decode stream, logits and greedy ids compared position by position
"""
from __future__ import annotations

import argparse
import hashlib
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
from umgen_b200.config import CONTENT_LEN, MODS, ModelConfig, SampleConfig  # noqa: E402

TOKENS_PER_FRAME = 2207
# SURVEY.md section 8d, per generated frame at UMGen_Large
DECODE_BYTES_PER_FRAME = 1420.4e9      # OAR weights 1122.89 GB + KV read/append 269.70 GB + heads 20.37 GB + GMLP 7.40 GB
ATTN_BYTES_PER_FRAME = 269.70e9
# dram__bytes_read.sum + dram__bytes_write.sum of one full-depth decode_cluster_kernel launch (2206 steps), ncu capture of
# tools/bench_decode.py 36 2206 2 (profiles/r1_traffic_cluster_full.csv): 1432.12 GB + 2.32 GB
DECODE_TRAFFIC = {"decode_cluster_kernel": 1434.44e9}
TAR_FLOP_PER_FRAME = 187.2e12
VQ_FLOP_PER_FRAME = {"map": 0.460e12, "image": 0.507e12}      # SURVEY.md section 2 (FlopCounterMode on the reference decoders)
LOGIT_ATOL, MARGIN_TOL = 3e-2, 6e-2    # full-depth parity: fp16 tensor-core chain vs the fp32 oracle (same bounds as tests/test_engine_gpu.py)
METRIC = "scene-tokens/sec, 30-frame video infer (steady-state frame)"


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


# ---------------------------------------------------------------------------------------------------------------
# the CPU arm: one real frame of the oracle (test infrastructure; bench.py's cpu_baseline / reference legs may run it)
# ---------------------------------------------------------------------------------------------------------------
def oracle_frame(cfg: ModelConfig, params, scene, threads: int):
    """One new frame through the oracle's UMGen.inference restatement (oracle/umgen_oracle.py) at the depth of `cfg`, greedy, 20 conditioning
    frames.  Returns a dict of plain tensors: what the GPU parity check needs plus the measured seconds."""
    from oracle import umgen_oracle as O
    torch.set_num_threads(threads)
    ocfg = O.ModelCfg(n_tar_layer=cfg.n_tar_layer, n_oar_layer=cfg.n_oar_layer, n_ego_tar_layer=cfg.n_ego_tar_layer,
                      n_ego_ca_layer=cfg.n_ego_ca_layer, n_map_tar_layer=cfg.n_map_tar_layer, n_box_tar_layer=cfg.n_box_tar_layer,
                      cond_frame=cfg.cond_frame, rule_constrain=cfg.rule_constrain, merage_ar_tar=cfg.merage_ar_tar)
    orc = O.UMGenOracle(params, ocfg, O.SampleCfg.greedy())
    orc.keep_trace = True
    T = cfg.cond_frame
    t0 = time.time()
    with torch.no_grad():
        out = orc.inference(1, T, T, {m: scene[m][:, :T] for m in MODS})
    sec = time.time() - t0
    tr = orc.trace[0]
    pos = sorted(p for p in tr.logits if p > 0)
    top = [torch.topk(tr.logits[p], 8) for p in pos]
    return {"seconds": sec, "threads": threads, "positions": torch.tensor(pos), "top_vals": torch.stack([t.values for t in top]),
            "stream": torch.tensor([tr.stream[p] for p in pos]), "tar_feat": tr.tar_feat.clone(), "ego_logits": tr.ego_logits.clone(),
            "pose": torch.from_numpy(out["pose"][0, T]).clone(), "tokens": tr.tokens.clone()}


def oracle_cache_path(cfg: ModelConfig, scene_seed: int) -> str:
    key = hashlib.sha256(json.dumps([cfg.to_dict(), scene_seed, "v2"]).encode()).hexdigest()[:12]
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, f"bench_oracle_frame_{key}.pt")


def get_oracle_frame(cfg, params, scene, scene_seed, threads, reuse=True):
    """The two arms run back to back on one box.  The reference arm (reuse=False) ALWAYS measures its frame inside its own run -- its line never
    quotes a time taken by another process -- and leaves it behind; our arm's `cpu_baseline` / `parity_fulldepth` leg may pick that frame up
    (same box and boot only: the file carries host name + boot id) instead of spending the same minutes of CPU time again, and says so in `sample`."""
    import platform
    try:
        here = platform.node() + ":" + open("/proc/sys/kernel/random/boot_id").read().strip()      # this box, this boot
    except OSError:
        here = platform.node()
    path = oracle_cache_path(cfg, scene_seed)
    if reuse and os.path.exists(path):
        try:
            d = torch.load(path)
            if d.get("host") == here:
                d["cached"] = True
                return d
        except Exception:
            pass
    d = oracle_frame(cfg, params, scene, threads)
    d["cached"] = False
    d["host"] = here
    try:
        torch.save(d, path)
    except Exception:
        pass
    return d


def cpu_line(of, cfg) -> dict:
    v = TOKENS_PER_FRAME / of["seconds"]
    return {"value": v, "unit": "tokens/s", "cores": int(of["threads"]), "kind": "port",
            "sample": f"ONE real generated frame of the CPU oracle (oracle/umgen_oracle.py: the reference algorithm restated in fp32) at full depth "
                      f"{cfg.n_ego_tar_layer}/{cfg.n_ego_ca_layer}/{cfg.n_map_tar_layer}/{cfg.n_box_tar_layer}/{cfg.n_tar_layer}/{cfg.n_oar_layer}, 20 conditioning frames, same "
                      f"weights and scene as the GPU arm: ego net + map/box/full TAR passes + 2202 OAR steps with sampling and the bbox rule path, "
                      f"measured {of['seconds']:.1f} s on {of['threads']} threads" + (" (computed by the other bench arm on this box, reused)" if of.get("cached") else "")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cfg = ModelConfig.tiny(args.layers) if args.layers else ModelConfig.large()
    P = synth.LazyParams(cfg, 0)
    scene = synth.make_scene(seed=1, n_frames=cfg.cond_frame)
    of = get_oracle_frame(cfg, P, scene, 1, threads, reuse=False)      # measured in THIS run, every run
    v = TOKENS_PER_FRAME / of["seconds"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "tokens/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
        "steps_note": "one real full-depth frame is timed whatever --steps / --warmup say (it takes minutes of CPU time)",
        "ms_per_step": 1e3 * of["seconds"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "UMGen_Large 30-frame free video infer, batch 1 (BASELINE configs[1]); step = one generated frame; CPU, reference algorithm via the oracle port",
                   "cond_frames": cfg.cond_frame, "tokens_per_frame": TOKENS_PER_FRAME, "layers": cfg.to_dict(), "sampling": "greedy (top-k 1)"},
        "cpu_baseline": cpu_line(of, cfg),
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_s": v / TOKENS_PER_FRAME,
    }
    print(json.dumps(line), flush=True)


def run_reference_gpu(args):
    """The unmodified reference on the GPU -- its own CUDA path with the installed flash_attn, fp16 autocast, batch 1 (SURVEY.md 8d "O2", the GPU
    baseline to beat) -- through its public `UMGen.inference`, same weights, scene and greedy recipe as the other arms.  Needs the reference tree
    AND a GPU in the same place; the GPU boxes of this project do not mount /root/reference, so there the line says `unavailable`.
    `--layers N` (debug) lowers the depth; UMGEN_REFGPU_ALLOW_CPU=1 runs the same harness on the CPU with the oracle's two import patches
    (plumbing check in the build container)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import ref_import
    if not ref_import.available():
        print(json.dumps({"impl": "reference-gpu", "unavailable": "the reference tree (/root/reference) is not mounted on this box; it needs mmcv / Lightning stand-ins "
                          "besides (tests/shims) -- the native GPU baseline cannot be timed here"}), flush=True)
        return
    cpu_ok = bool(os.environ.get("UMGEN_REFGPU_ALLOW_CPU"))
    if not torch.cuda.is_available() and not cpu_ok:
        print(json.dumps({"impl": "reference-gpu", "unavailable": "reference tree present but no CUDA device here (the build container); set UMGEN_REFGPU_ALLOW_CPU=1 "
                          "with --layers 1 to check the harness on the CPU"}), flush=True)
        return
    native = torch.cuda.is_available()
    cfg = ModelConfig.tiny(args.layers) if args.layers else ModelConfig.large()
    rcfg = ref_import.reference_config(native=native, n_tar_layer=cfg.n_tar_layer, n_oar_layer=cfg.n_oar_layer, n_ego_tar_layer=cfg.n_ego_tar_layer,
                                       n_ego_ca_layer=cfg.n_ego_ca_layer, n_map_tar_layer=cfg.n_map_tar_layer, n_box_tar_layer=cfg.n_box_tar_layer,
                                       cond_frame=cfg.cond_frame)
    params = synth.LazyParams(cfg, 0)
    model = ref_import.build_reference_model(rcfg, state_dict=None, greedy=True)
    sd = model.state_dict()
    with torch.no_grad():
        for k in list(sd):
            sd[k].copy_(params[k].to(sd[k].dtype))            # the synthetic checkpoint of the other arms, key by key (same state_dict ABI)
    if native:
        model = model.cuda()
        rcfg.device_set = torch.device("cuda")
    scene = synth.make_scene(seed=1, n_frames=cfg.cond_frame)
    T = cfg.cond_frame
    cond = {m: scene[m][:, :T] for m in MODS}

    def rollout(n):
        t0 = time.time()
        with torch.no_grad():
            out = model.inference(n, T, T, "pose_map_bbox3d_image", {m: v.clone() for m, v in cond.items()})
        if native:
            torch.cuda.synchronize()
        return time.time() - t0, out

    steps, warm = max(1, args.steps), max(0, args.warmup)
    if warm:
        rollout(min(warm, 1))                                  # kernels loaded, caches warm; the reference recomputes everything per frame anyway
    sec, out = rollout(steps)
    assert out["map"].shape == (1, T + steps, CONTENT_LEN["map"])
    v = TOKENS_PER_FRAME * steps / sec
    print(json.dumps({
        "impl": "reference-gpu", "metric": METRIC, "value": v, "unit": "tokens/s", "n_gpus": 1, "steps": steps, "warmup": min(warm, 1),
        "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16 autocast (the reference's own)" if native else "f32 (CPU harness check)", "data": "synthetic",
        "config": {"workload": "UMGen_Large 30-frame free video infer, batch 1 (BASELINE configs[1]); the unmodified reference through UMGen.inference"
                               + ("" if native else " -- ON THE CPU with patched attention: a harness check, not a baseline"),
                   "cond_frames": T, "tokens_per_frame": TOKENS_PER_FRAME, "layers": cfg.to_dict(), "sampling": "greedy (top-k 1)",
                   "timing": "host wall clock around one K-frame inference() call, device synchronised (the reference syncs every frame itself)"},
        "frames_per_s": steps / sec,
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def build_model(cfg: ModelConfig, dev, real_weights: bool, method: str = "topk"):
    """The drop-in module (projects/models/UMGen.py) with the evaluation Namespace, greedy recipe of SURVEY.md 3.4.  real_weights: parameters are
    drawn on the CPU (bit-identical to what the oracle gets); else the module is a parameter-less shell around an engine whose weights are drawn on the device."""
    from projects.models.UMGen import UMGen
    import contextlib
    ns = synth.evaluation_namespace(cfg, top_k=1, top_k_map=1, sample_method=method, skip_init=not real_weights)
    with contextlib.redirect_stdout(sys.stderr):       # the constructor prints its parameter count like the reference's; stdout carries the one JSON line
        model = UMGen(ns).eval()
    model.sample_param_map = 1 if method == "topk" else model.sample_param_map
    model.topk_image = 1 if method == "topk" else model.topk_image
    if not real_weights:
        from umgen_b200.engine import UMGenEngine
        model._engine = UMGenEngine(synth.DeviceParams(cfg, seed=0, device=dev), cfg, model._sample_config(0), device=dev)
    return model


def attention_path_seconds(cfg: ModelConfig):
    """Seconds per frame the decode kernel spends on its attention path (cache tiles -> scores -> softmax -> P V -> merged partials), from the
    profiling build of the same kernel (-DUMGEN_DECODE_PROFILE=1 accumulates the clock of CTA 0 / thread 0 around that phase): run in a subprocess
    because the library is chosen at load time."""
    from umgen_b200 import build as B
    lib = os.path.join(B.LIBDIR, "libumgen_sm100.prof1.so")
    if not os.path.exists(lib):
        return None
    code = ("import sys, json, torch, dataclasses; sys.path.insert(0, %r); sys.argv=['x'];"
            "sys.path.insert(0, %r);"
            "import bench_decode as bd; from umgen_b200.config import ModelConfig, SampleConfig; from umgen_b200.decoder import FrameDecoder;"
            "L=%d; dev=torch.device('cuda:0'); cfg=dataclasses.replace(ModelConfig.large(), n_oar_layer=L);"
            "dec=FrameDecoder({}, cfg, packed=bd.random_packed(L, dev)); tar=torch.randn(2207,768,device=dev); prev=torch.full((660,),1027);"
            "pose=torch.tensor([5,6,7]);"
            "[dec.decode(tar,pose,prev,SampleConfig.greedy(),check=False) for _ in range(2)]; torch.cuda.synchronize();"
            "e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record();"
            "r=dec.decode(tar,pose,prev,SampleConfig.greedy(),check=False); e1.record(); torch.cuda.synchronize();"
            "st=r.status.cpu().tolist(); print(json.dumps({'ms': e0.elapsed_time(e1), 'total_kc': st[60], 'attn_kc': st[64], 'head_kc': st[65]}))"
            % (ROOT, os.path.join(ROOT, "tools"), cfg.n_oar_layer))
    env = dict(os.environ, UMGEN_LIB=lib)
    try:
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        if d["total_kc"] <= 0:
            return None
        d["attn_seconds"] = d["ms"] / 1e3 * d["attn_kc"] / d["total_kc"]
        d["head_seconds"] = d["ms"] / 1e3 * d["head_kc"] / d["total_kc"]
        return d
    except Exception:
        return None


def scene_tail_line(frames: int = 50) -> dict:
    """Everything that follows a rollout, per scene (same calls as tools/bench_scene_tail.py): token pickle, `decode_tokens`
    (box / pose values on the host, map and image pixels through the GPU decoders in 6-frame pieces, results on the host), scene video (host, OpenCV).
    A 50-frame scene = 20 conditioning + 30 generated frames at evaluate.py's sizes; second of two runs."""
    import tempfile
    import numpy as np
    from umgen_b200 import postprocess, runner
    from umgen_b200.visualize import SceneVideo
    from umgen_b200.vq import Imagedecoder, Mapdecoder
    tmp = tempfile.mkdtemp(prefix="umgen_tail_")
    scene = synth.make_scene(seed=5, n_frames=frames)
    tokens = {m: scene[m][:, :frames].numpy().astype(np.int64) for m in MODS}
    md, idec = Mapdecoder(synth.make_vq_state_dict("map", seed=1)), Imagedecoder(synth.make_vq_state_dict("image", seed=1))
    video = SceneVideo(video_save_path=os.path.join(tmp, "clips/"), video_pretext="UMGen", width=512, height=512, project_name="UMGen_infer",
                       spe_text="bench", addtion_ego=True, cond_frames=20, put_text=True)
    out = {}
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.time()
        postprocess.save_tokens(tokens, os.path.join(tmp, "tokens"), f"scene{rep}")
        t1 = time.time()
        decoded = postprocess.decode_tokens(dict(tokens), {m: tokens[m] for m in ("pose", "bbox3d")}, md, idec)
        torch.cuda.synchronize()
        t2 = time.time()
        runner.write_scene_video(video, decoded, f"scene{rep}")
        t3 = time.time()
        out = {"frames": frames, "token_pickle_ms": 1e3 * (t1 - t0), "decode_tokens_ms": 1e3 * (t2 - t1), "scene_video_ms": 1e3 * (t3 - t2),
               "seconds_per_scene": t3 - t0, "note": "host wall clock, one thread; not part of any timed region of the headline"}
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return out


def vq_line(dev, hbm_peak, tf_peak):
    """The VQ pixel decoders (a11): decode a 6-frame chunk of map and image tokens like tools/model_pl.py:418-442 does."""
    from umgen_b200.vq import Imagedecoder, Mapdecoder
    out = {}
    for kind, cls, n_tok in (("map", Mapdecoder, 1024), ("image", Imagedecoder, 512)):
        dec = cls(synth.make_vq_state_dict(kind, seed=1), dev)
        tok = torch.randint(0, 8192, (6, n_tok), generator=torch.Generator().manual_seed(3))
        fn = dec.decode_maps if kind == "map" else dec.decode_images
        fn(tok)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn(tok)
        e1.record()
        torch.cuda.synchronize()
        s = e0.elapsed_time(e1) / 1e3 / 3 / 6
        out[kind] = {"frames_per_s": 1.0 / s, "tflops": VQ_FLOP_PER_FRAME[kind] / s / 1e12, "frac_of_tensor_peak": VQ_FLOP_PER_FRAME[kind] / s / 1e12 / tf_peak}
    return out


def main():
    import faulthandler
    import signal
    faulthandler.register(signal.SIGUSR1, all_threads=True)      # kill -USR1 <pid> prints where a stuck run is
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--layers", type=int, default=0, help="debug: override every stack depth (not a valid benchmark)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the full-depth oracle frame (cpu_baseline, parity_fulldepth) and the extras")
    ap.add_argument("--no-overlap", action="store_true", help="run the box_tar pass before the decode kernel instead of beside it")
    ap.add_argument("--no-lookahead", action="store_true", help="recompute the whole 20-frame window every frame (no TAR work beside the decode kernel)")
    ap.add_argument("--decode-kernel", type=int, default=0, help="0 = default (8-cluster kernel), 1 = L2-exchange kernel, 2 = 8-cluster kernel")
    ap.add_argument("--workload", default="video", choices=["video", "control", "long"],
                    help="video: BASELINE configs[1] (the headline); control: configs[3], 13 conditioning + 30 new frames with a forced agent slot and ego poses; "
                         "long: configs[4], 120 new frames.  control / long print their own line (one whole rollout through UMGen.inference)")
    ap.add_argument("--scenes-per-gpu", type=int, default=1,
                    help="B > 1: B scenes per GPU decoded in lockstep by one launch per frame (SURVEY.md 8f rank 1; the reference is batch 1): prints its own line, "
                         "the headline (B = 1) is unchanged")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "reference-gpu":
        return run_reference_gpu(args)

    import torch.distributed as dist
    from umgen_b200 import capi

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
    want_cpu = (not args.no_cpu_baseline) and world == 1 and args.workload == "video" and args.scenes_per_gpu == 1
    model = build_model(cfg, dev, real_weights=want_cpu)
    eng = model._get_engine(0)
    eng.check_status = True
    eng.dec.mode = args.decode_kernel
    eng.overlap = not args.no_overlap
    eng.lookahead = not args.no_lookahead
    if world > 1:       # weights come from rank 0 over NCCL (NVLink / NVSwitch); every rank then owns a replica
        from umgen_b200 import dp
        dp.broadcast_tensors(dp.engine_tensors(eng), src=0)
    T = cfg.cond_frame
    lib = capi.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload != "video":
        return run_rollout_workload(args, model, eng, cfg, rank, world, local, barrier)
    if args.scenes_per_gpu > 1:
        return run_batch_workload(args, model, eng, cfg, rank, world, local, barrier)

    scene = synth.make_scene(seed=1 + rank, n_frames=T)
    cond_host = {m: scene[m][0].clone() for m in MODS}
    tok_dev = {m: cond_host[m].to(torch.int32).pin_memory().to(dev, non_blocking=True) for m in MODS}

    # ---- device-resident rollout: `value` -------------------------------------------------------------------------------------------
    # A step is one generated frame of a real rollout: the window slides by the frame just generated, so consecutive steps continue each other
    # the way evaluate.py's loop does.
    state = {"win": tok_dev}
    eng.check_status = False

    def slide(win, new):
        return {m: torch.cat([win[m][1:], new[m].to(torch.int32)[None]], dim=0).contiguous() for m in MODS}

    def step_device():
        new = eng.frame_device(state["win"], continues=True)
        state["win"] = slide(state["win"], new)

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.umgen_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    dec_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps + 1)]
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
    t_decode = sum(a.elapsed_time(b) for a, b in dec_ev[:args.steps]) / args.steps / 1e3
    status = eng.dec.status.cpu().tolist()
    if status[0] != 0:
        raise SystemExit(f"decode kernel aborted (code {status[0]})")

    # ---- end to end through the plugin call: UMGen.inference on HOST tensors, one contiguous rollout of `steps` frames ------------------
    eng.check_status = True
    host_scene = {m: scene[m][:, :T].clone() for m in MODS}                    # int64 [1, 20, S_mod] like the DataLoader hands over
    kw = dict(cond_frames=T, input_cond_frames=T, pred_task="pose_map_bbox3d_image", init_tokens=None, cond_on_par=True, infer_from_gt=False, seed=0)
    model.inference(new_frames=2, input_cond_tokens=host_scene, **kw)          # untimed: the first frame of a rollout computes its whole window
    barrier()
    e2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e2[0].record()
    out = model.inference(new_frames=args.steps + 1, input_cond_tokens=host_scene, **kw)
    e2[1].record()
    barrier()
    ms_e2e_total = e2[0].elapsed_time(e2[1])
    assert out["map"].shape == (1, T + args.steps + 1, 1024) and out["map"].dtype.name == "int64"
    # the rollout's first frame has no look-ahead pass behind it (whole window computed): time it separately and report the steady-state frames
    e3 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e3[0].record()
    model.inference(new_frames=1, input_cond_tokens=host_scene, **kw)
    e3[1].record()
    barrier()
    ms_first = e3[0].elapsed_time(e3[1])
    ms_e2e = ms_e2e_total - ms_first                                           # `steps` steady-state frames

    # ---- TAR side alone (ego net + the three passes), sequential schedule, one extra frame ------------------------------------------
    overlapped = eng.overlap and eng.dec.kernel_name == "decode_cluster_kernel"
    lookahead = eng.lookahead and eng.dec.kernel_name != "decode_frame_kernel"
    eng.overlap = False
    eng.lookahead = False
    eng.check_status = False
    k["i"] = args.steps
    eng.dec.decode = timed_decode
    e4 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e4[0].record()
    step_device()
    e4[1].record()
    torch.cuda.synchronize()
    eng.dec.decode = orig_decode
    t_tar_seq = (e4[0].elapsed_time(e4[1]) - dec_ev[args.steps][0].elapsed_time(dec_ev[args.steps][1])) / 1e3
    eng.overlap = not args.no_overlap
    eng.lookahead = not args.no_lookahead

    t = torch.tensor([ms, ms_e2e, t_decode, t_tar_seq, ms_first], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, t_decode, t_tar_seq, ms_first = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, tf_peak, peak_src = peaks()
    frames = args.steps * world
    value = TOKENS_PER_FRAME * frames / (ms / 1e3)
    e2e_value = TOKENS_PER_FRAME * frames / (ms_e2e / 1e3)
    t_tar = max(t_tar_seq, 1e-9)
    scale = (cfg.n_oar_layer / 36.0)
    achieved = DECODE_BYTES_PER_FRAME * scale / t_decode / 1e9
    h2d = sum(CONTENT_LEN[m] * T * 4 for m in MODS)
    line = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world,
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
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": sum(CONTENT_LEN[m] for m in MODS) * 8 + 96 * 4,
                "api": "projects.models.UMGen.UMGen.inference(new_frames=steps + 1, cond_frames=20, input_cond_tokens={mod: LongTensor[1,20,S_mod] on the host}) -> "
                       "{mod: np.int64[1, 20 + new, S_mod]}; the first frame of the rollout (no look-ahead pass behind it, whole window computed: "
                       f"{ms_first:.0f} ms) is timed by a second one-frame call and subtracted, the remaining `steps` frames are the steady state",
                "first_frame_ms": ms_first},
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
                         "seconds_per_frame": t_tar, "note": "ego net + map/box/full TAR passes, timed in one extra frame with the sequential schedule (whole window "
                         "recomputed); in the headline frames most of it runs beside the decode kernel"},
    }
    if not args.no_cpu_baseline and world == 1:
        # extras (N = 1 only; none of them is inside a timed region above)
        ap_ = attention_path_seconds(cfg)
        if ap_ is not None:
            a_bw = ATTN_BYTES_PER_FRAME * scale / ap_["attn_seconds"] / 1e9
            line["roofline"]["attention"] = {"achieved": a_bw, "frac": a_bw / hbm_peak, "seconds_per_launch": ap_["attn_seconds"],
                                             "head_and_sampling_seconds_per_launch": ap_["head_seconds"],
                                             "how": "profiling build of the same kernel (-DUMGEN_DECODE_PROFILE=1): clock of CTA 0 / thread 0 accumulated over the attention path "
                                                    "(cache tiles -> scores -> softmax -> P V -> partial exchange and merge) of every layer and step, as a share of the launch "
                                                    f"({ap_['attn_kc']} of {ap_['total_kc']} kilo-cycles), applied to that launch's {ap_['ms']:.0f} ms"}
        try:
            line["vq"] = vq_line(dev, hbm_peak, tf_peak)
        except Exception as e:      # the VQ decoders are not on the decode path: report, do not fail the headline
            line["vq"] = {"error": str(e)[:200]}
        try:
            line["scene_tail"] = scene_tail_line()
        except Exception as e:      # what follows a rollout is not on the decode path either
            line["scene_tail"] = {"error": str(e)[:200]}
        # sampling as shipped by the evaluation config is top-p / top-k, not greedy: one more short rollout per sampler
        for name, sc in (("topk5", SampleConfig(method="topk", top_k=5, top_k_map=5, top_k_image=16, seed=1)),
                         ("topp0.4", SampleConfig(method="topp", p=0.4, p_map=0.4, seed=1))):
            eng.sample = sc
            eng.check_status = False
            step_device()
            torch.cuda.synchronize()
            es = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            es[0].record()
            for _ in range(2):
                step_device()
            es[1].record()
            torch.cuda.synchronize()
            line.setdefault("sampling_rows", {})[name] = {"tokens_per_s": TOKENS_PER_FRAME * 2 / (es[0].elapsed_time(es[1]) / 1e3)}
        eng.sample = SampleConfig.greedy()
    if want_cpu:
        threads = os.cpu_count() or 1
        sd = dict(model.state_dict())
        sd.update(model._fixed)
        of = get_oracle_frame(cfg, sd, {m: scene[m] for m in MODS}, 1 + rank, threads)
        line["cpu_baseline"] = cpu_line(of, cfg)
        line["parity_fulldepth"] = parity_fulldepth(eng, of, cond_host)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def parity_fulldepth(eng, of, cond_host) -> dict:
    """The GPU engine on the oracle's frame: same window, the oracle's ego action, decode teacher-forced on the oracle's stream.  Compares the ego
    logits, the conditioning feature, the top-8 AR logits of every sampled position and the greedy pick wherever the oracle's top-2 gap exceeds MARGIN_TOL."""
    from umgen_b200.tar import TarEncoders
    eng.keep_trace, eng.want_logits = True, True
    eng.trace.clear()
    eng.check_status = True
    eng.sample = SampleConfig.greedy()
    cond = {m: cond_host[m].clone() for m in MODS}
    tok = TarEncoders.to_device_tokens(cond, eng.dev)
    eng.tar.ego_action(tok, eng.sample, 0)
    ego_err = float((eng.tar.ego_logits.cpu() - of["ego_logits"]).abs().max())
    pos = of["positions"].tolist()
    teacher = torch.zeros(TOKENS_PER_FRAME, dtype=torch.int64)
    teacher[[p - 1 for p in pos]] = of["stream"].long()
    eng.lookahead = False
    eng.frame(cond, {"pose": of["pose"]}, False, teacher=teacher)
    eng.lookahead = True
    tr = eng.trace[-1]
    feat_err = float((tr.tar_feat.cpu() - of["tar_feat"]).abs().max())
    logits = tr.logits.cpu()
    picks = tr.picks.cpu()[[p - 1 for p in pos]]
    worst = 0.0
    for i, p in enumerate(pos):
        V = 1028 if 1033 <= p <= 1692 else 8192
        worst = max(worst, float((torch.topk(logits[p - 1, :V], 8).values - of["top_vals"][i]).abs().max()))
    margins = of["top_vals"][:, 0] - of["top_vals"][:, 1]
    confident = margins > MARGIN_TOL
    mism = (picks != of["stream"].long()) & confident
    eng.keep_trace, eng.want_logits = False, False
    eng.trace.clear()
    return {"positions_checked": len(pos), "max_logit_err": worst, "logit_tolerance": LOGIT_ATOL, "margin_tolerance": MARGIN_TOL,
            "confident_positions": int(confident.sum()), "confident_mismatches": int(mism.sum()),
            "low_margin_differences": int(((picks != of["stream"].long()) & ~confident).sum()),
            "ego_logit_err": ego_err, "conditioning_feature_err": feat_err,
            "pass": bool(worst < LOGIT_ATOL and int(mism.sum()) == 0 and ego_err < LOGIT_ATOL and feat_err < LOGIT_ATOL),
            "how": "GPU engine (fp16 matrices, fp32 accumulate) vs the fp32 CPU oracle at full depth on the same weights and 20-frame window, decode teacher-forced on "
                   "the oracle's stream; ids must agree wherever the oracle's top-2 logit gap exceeds margin_tolerance"}


def run_batch_workload(args, model, eng, cfg, rank, world, local, barrier):
    """B scenes per GPU: the headline working point (UMGen_Large, 20-frame window, greedy) with B independent scenes advancing in
    lockstep on every GPU -- one decode launch per frame for all of them, the look-ahead passes of the B next windows beside it.  A step = B generated frames."""
    import torch.distributed as dist
    from umgen_b200.decoder import FrameDecoder
    B, T, dev = args.scenes_per_gpu, cfg.cond_frame, eng.dev
    beng = model._get_engine(0, B)
    scenes = [synth.make_scene(seed=1 + rank * B + k, n_frames=T) for k in range(B)]
    wins = [{m: sc[m][0].to(torch.int32).pin_memory().to(dev, non_blocking=True) for m in MODS} for sc in scenes]
    beng.check_status = False

    def step_device():
        new = beng.frames_device(wins, continues=[True] * B)
        for k in range(B):
            wins[k] = {m: torch.cat([wins[k][m][1:], new[k][m].to(torch.int32)[None]], dim=0).contiguous() for m in MODS}

    for _ in range(max(args.warmup, 2)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib = __import__("umgen_b200.capi", fromlist=["lib"]).lib()
    launches0 = lib.umgen_launch_count()
    dec_ev = []
    orig = FrameDecoder.decode_batch

    def timed(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(*a, **kw)
        e1.record()
        dec_ev.append((e0, e1))
        return r

    FrameDecoder.decode_batch = staticmethod(timed)
    beng.time_lookahead = True
    la = []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(args.steps):
        step_device()
        if beng.la_events is not None:
            la.append(beng.la_events)
            beng.la_events = None
    ev[1].record()
    barrier()
    FrameDecoder.decode_batch = staticmethod(orig)
    beng.time_lookahead = False
    launches = lib.umgen_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = ev[0].elapsed_time(ev[1])
    t_decode = sum(a.elapsed_time(b) for a, b in dec_ev) / len(dec_ev) / 1e3
    for e in beng.engines:
        if int(e.dec.status.cpu()[0]) != 0:
            raise SystemExit("decode kernel aborted")
    # end to end through the plugin call with a batch of B scenes on the host (an extension of the reference's batch-1 signature)
    beng.check_status = True
    host = {m: torch.cat([sc[m][:, :T] for sc in scenes], dim=0).clone() for m in MODS}        # int64 [B, 20, S_mod]
    kw = dict(cond_frames=T, input_cond_frames=T, pred_task="pose_map_bbox3d_image", init_tokens=None, cond_on_par=True, infer_from_gt=False, seed=0)
    model.inference(new_frames=2, input_cond_tokens=host, **kw)
    barrier()
    e2 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e2[0].record()
    out = model.inference(new_frames=args.steps + 1, input_cond_tokens=host, **kw)
    e2[1].record()
    barrier()
    assert out["map"].shape == (B, T + args.steps + 1, 1024)
    e2[2].record()
    model.inference(new_frames=1, input_cond_tokens=host, **kw)
    e2[3].record()
    barrier()
    ms_first = e2[2].elapsed_time(e2[3])
    ms_e2e = e2[0].elapsed_time(e2[1]) - ms_first
    t = torch.tensor([ms, ms_e2e, t_decode, ms_first], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, t_decode, ms_first = t.tolist()
    if rank == 0:
        hbm_peak, tf_peak, peak_src = peaks()
        frames = args.steps * B * world
        scale = cfg.n_oar_layer / 36.0
        # SURVEY.md 8d per frame: weights 1122.89 GB + heads 20.37 GB are read once per step whatever B is; KV 269.70 GB and the embedding rows 7.40 GB per scene
        bytes_launch = ((1122.89e9 + 20.37e9) + B * (269.70e9 + 7.40e9)) * scale
        la_ms = [(a.elapsed_time(b), a.elapsed_time(c)) for a, b, c in la]
        print(json.dumps({
            "metric": METRIC, "value": TOKENS_PER_FRAME * frames / (ms / 1e3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 2),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16 (fp32 accumulate / residual)", "data": "synthetic",
            "config": {"workload": f"UMGen_Large 30-frame free video infer, {B} scenes per GPU decoded in lockstep (SURVEY.md 8f rank 1; BASELINE configs[1] is batch 1: "
                                   "see the default line); step = one generated frame of every scene",
                       "scenes_per_gpu": B, "cond_frames": T, "tokens_per_frame": TOKENS_PER_FRAME, "layers": cfg.to_dict(), "sampling": "greedy (top-k 1)",
                       "schedule": "one decode launch per step for all scenes; the look-ahead passes of the B next windows run beside it on the 84 free SMs",
                       "lookahead_ms": ({"passes_beside_decode": sum(x for x, _ in la_ms) / len(la_ms), "decode_kernel": sum(y for _, y in la_ms) / len(la_ms)} if la_ms else None),
                       "l2": "per-step working set (4.9 GB of fp16 weights + KV of B scenes) exceeds the 126 MB L2; no explicit flush"},
            "frames_per_s": frames / (ms / 1e3),
            "e2e": {"value": TOKENS_PER_FRAME * frames / (ms_e2e / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": B * sum(CONTENT_LEN[m] * T * 4 for m in MODS),
                    "d2h_bytes_per_step": B * (sum(CONTENT_LEN[m] for m in MODS) * 8 + 96 * 4),
                    "api": f"projects.models.UMGen.UMGen.inference(new_frames=steps + 1, input_cond_tokens={{mod: LongTensor[{B},20,S_mod] on the host}}) -> "
                           f"{{mod: np.int64[{B}, 20 + new, S_mod]}}; first frame ({ms_first:.0f} ms, whole windows computed) subtracted as in the default line",
                    "first_frame_ms": ms_first},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"kernel": f"decode_cluster_kernel<{B}> (OAR decode of {B} scenes, 2206 steps/launch)", "bound": "hbm", "achieved": bytes_launch / t_decode / 1e9,
                         "peak": hbm_peak, "unit": "GB/s", "frac": bytes_launch / t_decode / 1e9 / hbm_peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_launch, "seconds_per_launch": t_decode,
                         "bytes_note": "weights and heads once per step, KV cache and embedding rows per scene (SURVEY.md 8d)"}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_rollout_workload(args, model, eng, cfg, rank, world, local, barrier):
    """BASELINE configs[3] (control: 13 conditioning + 30 new frames, forced ego poses and one forced agent slot, window growing 13 -> 20 then sliding) and
    configs[4] (long horizon: 120 new frames): one whole rollout through UMGen.inference on host tensors, one scene per rank."""
    import torch.distributed as dist
    control = args.workload == "control"
    n_in, new = (13, 30) if control else (20, 120)
    if args.layers:
        new = min(new, 6)
    scene = synth.make_scene(seed=1 + rank, n_frames=n_in)
    init = synth.make_control(seed=1 + rank, n_frames=new) if control else None
    kw = dict(cond_frames=cfg.cond_frame, input_cond_frames=n_in, pred_task="pose_map_bbox3d_image", cond_on_par=True, infer_from_gt=False, seed=0)
    model.inference(new_frames=2, input_cond_tokens=scene, init_tokens=init, control_test=control, **kw)       # warm-up: kernels, graphs of the first window lengths
    barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    out = model.inference(new_frames=new, input_cond_tokens=scene, init_tokens=init, control_test=control, **kw)
    e[1].record()
    barrier()
    t = torch.tensor([e[0].elapsed_time(e[1])], dtype=torch.float64, device=eng.dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t[0])
        assert out["bbox3d"].shape == (1, n_in + new, 660)
        print(json.dumps({
            "metric": METRIC, "value": TOKENS_PER_FRAME * new * world / (ms / 1e3), "unit": "tokens/s", "n_gpus": world, "steps": new, "warmup": 2,
            "ms_per_step": ms / new, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16 (fp32 accumulate / residual)", "data": "synthetic",
            "config": {"workload": ("BASELINE configs[3]: --infer_task control, 13 conditioning + 30 new frames, forced ego poses + one forced agent slot, 1 scene per GPU"
                                    if control else "BASELINE configs[4]: 120-frame long-horizon rollout, 1 scene per GPU"),
                       "api": "projects.models.UMGen.UMGen.inference on host tensors, whole rollout timed (first frame included)", "layers": cfg.to_dict(),
                       "sampling": "greedy (top-k 1)"},
            "frames_per_s": new * world / (ms / 1e3), "rollout_seconds": ms / 1e3}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
