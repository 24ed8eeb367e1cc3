"""Drop-in boundary: same ctor, same 81-key state_dict, reference loaders run unchanged."""
import copy
import os
import pickle
import sys

import numpy as np
import pytest
import torch

from helpers import REFERENCE, build_model, load_contract, torch_state_dict
from oracle import synth


def test_state_dict_layout_matches_reference_contract():
    c, hp, tasks = load_contract()
    m = build_model(hp, tasks)
    sd = m.state_dict()
    assert list(sd.keys()) == c["loaded_keys"]
    for k, shape, dtype in c["state_dict"]:
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dtype, k
    assert sum(p.numel() for p in m.parameters()) == c["n_params"] == 2_431_284


def test_strict_load_and_attributes():
    _, hp, tasks = load_contract()
    m = build_model(hp, tasks)
    missing, unexpected = m.load_state_dict(torch_state_dict(synth.pmg_state_dict(0)), strict=True)
    assert not missing and not unexpected
    assert m.patch_div == 3 and m.gram_channels == 96 and m.C == 128 and m.Cr == 96
    assert m.num_classes_per_task == tasks
    assert isinstance(m.se, torch.nn.Identity)
    convs = [l for l in m.feature_extractor if isinstance(l, torch.nn.Conv2d)]
    assert [c.out_channels for c in convs] == [32, 64, 128]
    assert list(m.task_pool.query.keys()) == sorted(tasks) or set(m.task_pool.query.keys()) == set(tasks)
    assert list(m.classifiers.keys()) == list(tasks)
    # kernel-side buffers never leak into the checkpoint
    assert not any(k.startswith("_") for k in m.state_dict())


def test_small_config_layout():
    _, hp, tasks = load_contract("pmg_small_contract.json")
    m = build_model(hp, tasks)
    assert list(m.state_dict().keys()) == list(synth.pmg_state_dict(2, hp, tasks).keys())


def test_error_behaviour_without_gpu():
    _, hp, tasks = load_contract()
    m = build_model(hp, tasks).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 224, 224))
    m.train()
    with pytest.raises(RuntimeError, match="eval"):
        m(torch.zeros(1, 3, 224, 224))
    with pytest.raises(TypeError):                       # the reference cannot construct norm="group" either (models_PMG.py:102-103,116)
        build_model({**hp, "norm": "group"}, tasks)
    with pytest.raises(ValueError):
        build_model(hp, tasks).eval()(torch.zeros(3, 224, 224))


def test_pickle_and_deepcopy_drop_kernel_state():
    _, hp, tasks = load_contract()
    m = build_model(hp, tasks)
    m2 = copy.deepcopy(m)
    m3 = pickle.loads(pickle.dumps(m))
    for other in (m2, m3):
        assert list(other.state_dict().keys()) == list(m.state_dict().keys())
        assert other._cfg.d_model == 192 and other._packed is None


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout only exists in the build container")
def test_reference_loaders_run_unchanged(tmp_path, capsys):
    """utils/PMG/PMG_model_utils.py (build_model_from_hparams, load_best_model, print_model_parameters)
    imported from the reference, bound to the B200 module through install_as_reference_models()."""
    import pmg_b200
    c, hp, tasks = load_contract()
    saved = {k: sys.modules.get(k) for k in ("Models", "Models.models_PMG", "utils", "utils.PMG", "utils.PMG.PMG_model_utils")}
    sys.path.insert(0, REFERENCE)
    try:
        for k in saved:
            sys.modules.pop(k, None)
        pmg_b200.install_as_reference_models()
        from utils.PMG.PMG_model_utils import build_model_from_hparams, load_best_model, print_model_parameters
        model = build_model_from_hparams(c["tasks"], hp, torch.device("cpu"))
        assert type(model) is pmg_b200.MultiTaskPatchGANGramModelNonOverlapV2
        ckpt = tmp_path / "best_model_PMG.pth"
        torch.save(torch_state_dict(synth.pmg_state_dict(0)), ckpt)
        model = load_best_model(model, str(ckpt), torch.device("cpu"))
        assert "strict=True" in capsys.readouterr().out
        assert not model.training
        print_model_parameters(model)
        out = capsys.readouterr().out
        assert "2,431,284" in out and "gram_channels:         96" in out
        w = model.token_proj.weight.detach().numpy()
        assert np.array_equal(w, synth.pmg_state_dict(0)["token_proj.weight"])
    finally:
        sys.path.remove(REFERENCE)
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def test_host_pipeline_piece_schedule():
    """HostPipeline cuts a host batch into H2D pieces: uniform `chunk`-sized pieces, the last one halving down to
    `tail_chunk` (the forward of the LAST piece is what nothing can overlap).  Pure host logic."""
    from pmg_b200.runner import HostPipeline
    hp = HostPipeline.__new__(HostPipeline)
    for chunk, tail, B, want in [(512, 32, 1024, [512, 256, 128, 64, 32, 32]), (128, 32, 1024, [128] * 7 + [64, 32, 32]),
                                 (8, 8, 21, [8, 8, 5]), (512, 32, 5, [5]), (512, 32, 33, [1, 32]), (512, 32, 0, [])]:
        hp.chunk, hp.tail_chunk = chunk, min(tail, chunk)
        edges = hp._schedule(B)
        assert [b - a for a, b in edges] == want, (chunk, tail, B, edges)
        assert all(edges[i][1] == edges[i + 1][0] for i in range(len(edges) - 1))
        assert (edges[0][0], edges[-1][1]) == (0, B) if B else edges == []
