"""Time the persistent OAR decode kernel on random weights (device-resident), full depth by default.
Usage: python tools/bench_decode.py [layers] [n_steps] [mode] [grid] [scenes per launch]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from umgen_b200.config import ModelConfig, SampleConfig  # noqa: E402
from umgen_b200.decoder import FrameDecoder  # noqa: E402
from umgen_b200.weights import box_value_lut  # noqa: E402
import dataclasses  # noqa: E402


def random_packed(L, dev):
    g = torch.Generator(device=dev).manual_seed(0)

    def h(*shape, scale=0.036):
        return ((torch.rand(*shape, generator=g, device=dev) * 2 - 1) * scale).half()

    def f(*shape):
        return torch.randn(*shape, generator=g, device=dev)

    from umgen_b200.capi import lib
    LH = 2304 * 768 + 768 * 768 + 3072 * 768 + 768 * 3072
    w = {"oar_h": h(L, LH), "oar_f": torch.cat([1 + 0.1 * f(L, 768), 0.03 * f(L, 2304), 0.03 * f(L, 768), 1 + 0.1 * f(L, 768)], 1).contiguous(),
         "ln_oar_f": 1 + 0.1 * f(768), "head_map_h": h(8192, 768), "head_bbox_h": h(1028, 768), "head_img_h": h(8192, 768),
         "head_tar_bbox_h": h(1028, 768), "map_table_f": f(8192, 768), "img_table_f": f(8192, 768),
         "be_f": f(1028, 768), "axe_f": f(8, 768), "tske_f": f(768), "fpe_f": f(1024, 768).clamp(-1, 1),
         "box_lut_d": torch.from_numpy(box_value_lut()).to(dev)}
    return w


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 36
    n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2206
    modes = [int(m) for m in sys.argv[3].split(',')] if len(sys.argv) > 3 else [2, 1]
    dev = torch.device("cuda:0")
    cfg = dataclasses.replace(ModelConfig.large(), n_oar_layer=L)
    dec = FrameDecoder({}, cfg, packed=random_packed(L, dev))
    tar = torch.randn(2207, 768, device=dev)
    pose = torch.tensor([5, 6, 7])
    prev = torch.full((660,), 1027)
    prev[:110] = 500
    dec.debug = torch.zeros(1024, 16, dtype=torch.int64, device=dev)
    dec.grid = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    n_scenes = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    decs = [dec] + [dec.for_scene() for _ in range(n_scenes - 1)]
    frames = [dict(tar_feat=torch.randn(2207, 768, device=dev) if s else tar, pose_tok=pose, prev_bbox=prev) for s in range(n_scenes)]
    for mode in modes:
        dec.mode = mode
        for it in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if n_scenes == 1:
                r = dec.decode(tar, pose, prev, SampleConfig.greedy(), n_steps=n_steps, check=False)
            else:
                r = FrameDecoder.decode_batch(decs, frames, SampleConfig.greedy(), n_steps=n_steps, check=False)[0]
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            st = r.status.cpu().tolist()
            # algorithmic bytes: weights per step + KV read/append + heads
            wbytes = L * 7_082_496 * 2 * n_steps
            kvbytes = n_scenes * sum(L * 2 * 768 * 2 * (n + 1) for n in range(1, n_steps + 1))
            tl = dec.debug.cpu()[:148, :10].double()
            if mode in (2, 3):
                print(f"   kilo-cycles cta0/thread0: total {st[60]} ring-wait {st[61]} dsmem-wait {st[62]} l2-poll {st[63]}")
                print(f"   cluster probes (cycles since layer start, cta0): {st[8:40]}  cta37: {st[40:59]}")
                if st[80]:
                    print(f"   producer of cta0, cycles since the consumers entered that layer: before c_attn issue, after it, after K/V, after c_proj, after c_fc, after MLP c_proj: {[v - st[80] for v in st[81:87]]}")
            if mode == 2 and it == 1 and bool((dec.debug != 0).any()):      # UMGEN_DECODE_PROFILE=3 timeline of the 8-cluster kernel: [cta 64][warp 12][stamp]
                tl2 = dec.debug.cpu()[:768, :16].double().view(64, 12, 16)
                names2 = ['start', 'ln1', 'qkv>', 'attn>', 'merge', 'proj>', 'hop0 L2<', 'hop0 x<', 'ln2', 'fc', 'proj2>', 'rs<', 'hop1 L2<', 'hop1 x<', 'ln2 bar>', 'ln2 bar<']
                t00 = tl2[:, :, 0].min()
                print('   timeline (cycles since the earliest warp entered the layer): min / median / max over all 768 warps; spread inside a CTA (max - min): median, max over CTAs')
                for i, nme in enumerate(names2):
                    col = tl2[:, :, i] - t00
                    sp = col.max(dim=1).values - col.min(dim=1).values
                    print(f'   {nme:9s} min {col.min():7.0f} med {col.median():7.0f} max {col.max():7.0f} | in-CTA spread med {sp.median():6.0f} max {sp.max():6.0f} | CTA-max: min {col.max(dim=1).values.min():7.0f} max {col.max(dim=1).values.max():7.0f}')
                print('   per cluster (globaltimer, ns since the first warp of the chip entered the layer): min..max over the cluster of each stamp')
                for i in (0, 3, 5, 7, 8, 10, 11, 13):
                    blk = (tl2[:, :, i] - t00).view(8, 96)
                    print(f'   {names2[i]:9s} ' + ' '.join(f'{a:6.0f}..{b:<6.0f}' for a, b in zip(blk.min(dim=1).values.tolist(), blk.max(dim=1).values.tolist())))
                for cl in (2,):          # one cluster in detail
                    base = tl2[8 * cl: 8 * cl + 8, :, 0].min()
                    for i in (0, 3, 5, 7, 14, 15, 8, 10, 11, 13):
                        blk = tl2[8 * cl: 8 * cl + 8, :, i] - base
                        print(f'   cluster {cl} {names2[i]:9s} per rank min..max: ' + ' '.join(f'{a:6.0f}..{b:<6.0f}' for a, b in zip(blk.min(dim=1).values.tolist(), blk.max(dim=1).values.tolist())))
                    for r in (0, 3):
                        for i in (7, 14, 15, 8):
                            print(f'   cluster {cl} rank {r} {names2[i]:9s} per warp: ' + ' '.join(f'{v - base:6.0f}' for v in tl2[8 * cl + r, :, i].tolist()))
            if mode == 3 and it == 1 and bool((dec.debug != 0).any()):      # UMGEN_DECODE_PROFILE=3 timeline: [cta 16][warp 12][stamp]
                tl3 = dec.debug.cpu()[:192, :16].double().view(16, 12, 16)
                names3 = ['start', 'ln1', 'qkv', 'attn', 'rs0>', 'rs0<', 'ag0<', 'ln2', 'fc', 'rs1>', 'rs1<', 'ag1<', 'bar0', 'bar1', 'ag0>', 'ag1>']
                t00 = tl3[:, :, 0].min()
                print('   timeline (cycles since the earliest warp entered the layer): per stamp min / median / max over all warps; per-CTA max')
                for i, nme in enumerate(names3):
                    col = tl3[:, :, i] - t00
                    print(f'   {nme:5s} min {col.min():7.0f} med {col.median():7.0f} max {col.max():7.0f} | per-CTA max: ' + ' '.join(f'{v:6.0f}' for v in col.max(dim=1).values.tolist()) + ' | per-CTA min: ' + ' '.join(f'{v:6.0f}' for v in col.min(dim=1).values.tolist()))
                for cta in (0, 4, 9):
                    for i in (4, 5, 12, 6, 9, 10, 13, 11):
                        print(f'   cta {cta:2d} {names3[i]:5s} per warp: ' + ' '.join(f'{v - t00:6.0f}' for v in tl3[cta, :, i].tolist()))
            if n_steps > 1200 and it == 1 and mode == 1:
                base = tl[:, 0].min()
                names = ['start', 'ln1', 'P1', 'attn', 'comb', 'P3', 'ln2', 'P4', 'rdH', 'P5']
                for i, nme in enumerate(names):
                    col = tl[:, i] - base
                    print(f'   {nme:5s} min {col.min():8.0f} med {col.median():8.0f} max {col.max():8.0f} ns  argmax cta {int(col.argmax())}')
            print(f"L={L} steps={n_steps} mode={mode} scenes={n_scenes} iter={it}: {ms:.1f} ms  ({ms * 1e3 / n_steps:.1f} us/step)  "
                  f"~{(wbytes + kvbytes) / ms / 1e6:.0f} GB/s  status={st[:4]} probes cta0 main={st[8:17]} attn={st[18:23]} | cta77 main={st[40:49]} attn={st[50:55]}", flush=True)


if __name__ == "__main__":
    main()

