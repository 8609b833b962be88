"""The oracle's token samplers (oracle/umgen_oracle.py sample_rows / SampleCfg.param) against the reference's own `UMGen.topk` and
`UMGen.sample_top_p` (UMGen.py:899-965): identical picks on seeded logits under the same torch seed (tests/golden/samplers.npz, made by
oracle/make_golden.py samplers from the unmodified reference), and the per-modality parameters the model hands to them (:118-126, 1004-1133).
The GPU samplers draw from a counter-based generator instead of torch's and are compared with these definitions distributionally
(tests/test_decode_gpu.py, tests/test_tar_kernels_gpu.py); this file pins the definitions themselves."""
import os

import numpy as np
import pytest
import torch

from oracle import umgen_oracle as O
from tests._cases import SAMPLER_CASES, sampler_logits


@pytest.mark.parametrize("case", SAMPLER_CASES, ids=[c[0] for c in SAMPLER_CASES])
def test_sampler_picks_equal_the_reference(case, golden_dir):
    name, method, param, vocab, rows, scale, seed = case
    g = np.load(os.path.join(golden_dir, "samplers.npz"))
    x = sampler_logits(vocab, rows, scale, seed)
    cfg = O.SampleCfg(method=method, temp=float(g[f"params_{method}"][3]))
    torch.manual_seed(seed)
    got = O.sample_rows(x, cfg, param, None).numpy()
    assert np.array_equal(got, g[name]), (name, got[:8], g[name][:8])
    # every pick lies in the truncation set of its row
    probs = torch.softmax(x / cfg.temp, dim=-1)
    if method == "topk":
        kth = torch.topk(x, min(int(param), vocab)).values[:, -1]
        assert bool((x[torch.arange(rows), torch.from_numpy(got)] >= kth).all())
    else:
        ps, pi = torch.sort(probs, dim=-1, descending=True)
        before = torch.cumsum(ps, dim=-1) - ps
        rank = (pi == torch.from_numpy(got)[:, None]).float().argmax(dim=-1)
        assert bool((before[torch.arange(rows), rank] <= float(param)).all())


def test_per_modality_parameters_equal_the_reference(golden_dir):
    """sample_param / sample_param_map / topk_image as the reference model holds them for either sample_method, and the image-branch quirk:
    under "topp" the image head is sampled with p = topk_image = 16, i.e. from the whole vocabulary (UMGen.py:1133)."""
    g = np.load(os.path.join(golden_dir, "samplers.npz"))
    for method in ("topk", "topp"):
        c = O.SampleCfg(method=method)
        ref_param, ref_param_map, ref_topk_image, temp = g[f"params_{method}"].tolist()
        assert float(c.param("pose")) == float(c.param("bbox3d")) == ref_param
        assert float(c.param("map")) == ref_param_map
        assert float(c.param("image")) == ref_topk_image == 16.0
        assert c.temp == temp


def test_greedy_is_top1_without_a_random_draw():
    x = sampler_logits(1028, 16, 3.0, 1)
    state = torch.random.get_rng_state()
    got = O.sample_rows(x, O.SampleCfg.greedy(), 1, None)
    assert torch.equal(got, x.argmax(dim=-1)) and torch.equal(torch.random.get_rng_state(), state)
