"""The drop-in ``projects`` surface: registry semantics, constructor from the evaluation
Namespace, state_dict key / shape / dtype compatibility with the reference model, loud failure without CUDA."""
import os
import sys
from argparse import Namespace

import pytest
import torch

from umgen_b200 import capi, synth
from umgen_b200.config import ModelConfig

REF = "/root/reference/projects"


def eval_namespace(layers=1, **over):
    return synth.evaluation_namespace(ModelConfig.tiny(layers), **over)


def test_registry_and_build_from_cfg():
    from projects.registry import MODELS, Registry, build_from_cfg
    from projects.models.UMGen import UMGen
    assert MODELS.get("UMGen") is UMGen
    m = build_from_cfg(dict(type=UMGen, config=eval_namespace()), MODELS)       # class object as type (evaluate.py:193)
    assert isinstance(m, torch.nn.Module)
    m2 = build_from_cfg(dict(type="UMGen", config=eval_namespace()), MODELS)
    assert type(m2) is UMGen
    r = Registry("x")
    r.register_module()(int)
    with pytest.raises(KeyError):
        r.register_module()(int)


def test_state_dict_matches_param_specs_and_roundtrips():
    from projects.models.UMGen import UMGen
    m = UMGen(eval_namespace(layers=1))
    sd = m.state_dict()
    specs = {k: (tuple(s), kind) for k, s, kind in synth.param_specs(ModelConfig.tiny(1))}
    assert set(sd) == set(specs)
    for k, v in sd.items():
        assert tuple(v.shape) == specs[k][0], k
    assert sd["fouier_pe"].dtype == torch.bfloat16 and sd["transformer.OAR.0.temporal_attn.scale"].dtype == torch.float32
    other = synth.make_state_dict(ModelConfig.tiny(1), seed=9)
    res = m.load_state_dict({"module": other}["module"], strict=False)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(m.state_dict()["transformer.be.weight"], other["transformer.be.weight"])
    m.eval(); m.cpu()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_state_dict_is_interchangeable_with_the_reference_model():
    from oracle import ref_import as R
    ref = R.build_reference_model(R.reference_config(layers=1))
    ref_sd = ref.state_dict()
    for k in [k for k in sys.modules if k == "projects" or k.startswith("projects.")]:      # back to our package
        del sys.modules[k]
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from projects.models.UMGen import UMGen
    ours = UMGen(eval_namespace(layers=1))
    sd = ours.state_dict()
    assert set(sd) == set(ref_sd)
    for k in sd:
        assert sd[k].shape == ref_sd[k].shape and sd[k].dtype == ref_sd[k].dtype, k
    assert not ours.load_state_dict(ref_sd, strict=True).missing_keys            # reference checkpoint -> our module
    ref.load_state_dict(sd, strict=True)                                          # and back


def test_inference_without_cuda_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from projects.models.UMGen import UMGen
    m = UMGen(eval_namespace(layers=1))
    scene = synth.make_scene(seed=1, n_frames=3)
    with pytest.raises(capi.UmgenError):
        m.inference(new_frames=1, cond_frames=2, input_cond_frames=2, pred_task="pose_map_bbox3d_image", input_cond_tokens=scene)
