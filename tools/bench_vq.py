#!/usr/bin/env python
"""Times the VQ pixel decoders (a11) on a 6-frame chunk, like bench.py's `vq` row, with each host-side switch of umgen_b200.vq.VQDecoder
(implicit-GEMM convolutions, slab GroupNorm statistics, CUDA-graph replay) off and on.  python tools/bench_vq.py [frames]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umgen_b200 import synth  # noqa: E402
from umgen_b200.vq import Imagedecoder, Mapdecoder  # noqa: E402

FLOP = {"map": 0.460e12, "image": 0.507e12}


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    dev = torch.device("cuda:0")
    rows = []
    for kind, cls, n_tok in (("map", Mapdecoder, 1024), ("image", Imagedecoder, 512)):
        dec = cls(synth.make_vq_state_dict(kind, seed=1), dev)
        core = dec.map_autoencoder if kind == "map" else dec.img_autoencoder
        fn = dec.decode_maps if kind == "map" else dec.decode_images
        tok = torch.randint(0, 8192, (frames, n_tok), generator=torch.Generator().manual_seed(3))
        for implicit, slab, graph in ((False, False, False), (True, False, False), (True, True, False), (True, True, True)):
            core.implicit_conv, core.slab_groupnorm, core.use_graph = implicit, slab, graph
            fn(tok)
            fn(tok)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn(tok)
            e1.record()
            torch.cuda.synchronize()
            s = e0.elapsed_time(e1) / 1e3 / 5 / frames
            rows.append({"kind": kind, "implicit_conv": implicit, "slab_groupnorm": slab, "graph": graph, "ms_per_frame": 1e3 * s,
                         "frames_per_s": 1 / s, "tflops": FLOP[kind] / s / 1e12})
            print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
