"""Oracle vs the reference's own answers (tests/golden/tables.npz, collision.npz)."""
import hashlib
import os

import numpy as np
import torch

from oracle import umgen_oracle as O
from tests._cases import collision_cases
from umgen_b200 import synth
from umgen_b200.config import ModelConfig


def _sha(t):
    return hashlib.sha256(t.contiguous().view(torch.uint8).numpy().tobytes()).hexdigest()


def test_value_luts_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "tables.npz"))
    assert np.array_equal(O.pose_value_lut(), g["pose_lut"])
    assert np.array_equal(O.box_value_lut(), g["box_lut"])
    assert int(g["pad_token"]) == O.PAD_TOKEN and int(g["bbox_vocab"]) == 1028


def test_fixed_tables_bitexact(golden_dir):
    g = np.load(os.path.join(golden_dir, "tables.npz"))
    want = dict(zip(g["fixed_keys"].tolist(), g["fixed_sha"].tolist()))
    sp = O.sinusoid_table(1030, 768, 1024)
    got = {"fouier_pe": _sha(O.sinusoid_table(1024, 768)), "bbox3d_spatial_posi": _sha(sp),
           "grid_center_posi_embedding": _sha(O.grid_center_embedding(sp))}
    assert got == want
    # the product's generator must produce the same fixed tables
    sd = synth.make_state_dict(ModelConfig.tiny(1), keys=list(want))
    assert {k: _sha(v) for k, v in sd.items()} == want


def test_sequence_maps(golden_dir):
    g = np.load(os.path.join(golden_dir, "tables.npz"))
    assert O.forced_positions() == dict(zip(g["dpos_keys"].tolist(), g["dpos_vals"].tolist()))
    got = np.array([O.MODS.index(O.pos_mod(p)) for p in range(1, 2208)], dtype=np.int8)
    assert np.array_equal(got, g["pos_mod"])


def test_param_count_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "tables.npz"))
    n = sum(int(np.prod(s)) for k, s, kind in synth.param_specs(ModelConfig.tiny(1)) if kind != "scale")
    assert n == int(g["n_params_1layer"])
    # SURVEY.md section 8b counts state_dict entries (the 348 scalar `scale` buffers included)
    n_large = sum(int(np.prod(s)) for k, s, kind in synth.param_specs(ModelConfig.large()))
    assert n_large == 2_447_224_924


def test_collision_matches_reference(golden_dir):
    ans = np.load(os.path.join(golden_dir, "collision.npz"))["answers"]
    cases = collision_cases()
    assert len(cases) == len(ans)
    got = np.array([O.check_collision(c) for c in cases])
    bad = np.nonzero(got != ans)[0]
    assert bad.size == 0, f"{bad.size} mismatches, first {bad[:5]}"
