"""Constructor variants outside the committed config (Models/models_PMG.py:15-30,100-105,133): norm = "instance",
use_channel_attention = True (SEBlock); norm = "group" raises TypeError exactly as the reference does.  Fixtures come from
the unmodified reference (tests/golden/make_variant_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, REFERENCE, argmax_table, rel_fro, split_packed, torch_state_dict
from oracle import pmg_oracle, synth

with open(os.path.join(GOLDEN, "pmg_variants.json")) as f:
    META = json.load(f)
TASKS = {t: len(c) for t, c in META["tasks"].items()}
NAMES = list(META["variants"])


def _hp(name):
    return dict(META["base"], **META["variants"][name])


def _build(name):
    import pmg_b200
    hp = _hp(name)
    return pmg_b200.MultiTaskPatchGANGramModelNonOverlapV2(
        input_nc=3, ndf=hp["ndf"], norm=hp["norm"], patch_size=hp["patch_size"], patch_div=hp["patch_div"],
        num_classes_per_task=TASKS, use_channel_attention=hp["use_channel_attention"], use_token_attention=True,
        gram_channels=hp["gram_channels"], d_model=hp["d_model"], transformer_layers=hp["transformer_layers"],
        transformer_heads=hp["transformer_heads"])


@pytest.mark.parametrize("name", NAMES)
def test_state_dict_layout_matches_reference(name):
    m = _build(name)
    assert [[k, list(v.shape)] for k, v in m.state_dict().items()] == META[f"{name}_keys"]
    missing, unexpected = m.load_state_dict(torch_state_dict(synth.pmg_state_dict(META["seed"], _hp(name), TASKS)), strict=True)
    assert not missing and not unexpected


def test_group_norm_raises_like_the_reference():
    import pmg_b200
    with pytest.raises(TypeError, match="num_groups"):
        pmg_b200.MultiTaskPatchGANGramModelNonOverlapV2(norm="group", num_classes_per_task=TASKS)
    if os.path.isdir(REFERENCE):
        import sys
        sys.path.insert(0, REFERENCE)
        try:
            import importlib
            ref = importlib.import_module("Models.models_PMG")
            if "b200" not in ref.__name__ and "pmg_b200" not in getattr(ref, "__file__", ""):
                with pytest.raises(TypeError, match="num_groups"):
                    ref.MultiTaskPatchGANGramModelNonOverlapV2(norm="group", num_classes_per_task=TASKS)
        finally:
            sys.path.remove(REFERENCE)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_fixture(name):
    g = np.load(os.path.join(GOLDEN, "pmg_variants.npz"))
    hp = _hp(name)
    sd = synth.pmg_state_dict(META["seed"], hp, TASKS)
    x = synth.images(META["seed"], 3, META["input"][1], META["input"][2])
    out, inter = pmg_oracle.pmg_forward(x, sd, TASKS, hp["patch_div"], hp["transformer_heads"], return_intermediates=True, norm=hp["norm"])
    assert rel_fro(pmg_oracle.pack_logits(out, TASKS), g[f"{name}_logits"]) <= 2e-5
    assert rel_fro(inter["gram"], g[f"{name}_gram"]) <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("in_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", NAMES)
def test_cuda_matches_reference_fixture(name, in_dtype):
    g = np.load(os.path.join(GOLDEN, "pmg_variants.npz"))
    hp = _hp(name)
    m = _build(name)
    m.load_state_dict(torch_state_dict(synth.pmg_state_dict(META["seed"], hp, TASKS)), strict=True)
    m.cuda().eval()
    x = torch.from_numpy(synth.images(META["seed"], 3, META["input"][1], META["input"][2])).cuda().to(in_dtype)
    with torch.no_grad():
        out, gram = m.forward_with_gram(x)
        packed = torch.cat([out[t] for t in TASKS], 1).cpu().numpy()
    ref = g[f"{name}_logits"]
    assert np.array_equal(argmax_table(packed, TASKS), argmax_table(ref, TASKS))
    for t, v in split_packed(packed, TASKS).items():
        assert rel_fro(v, split_packed(ref, TASKS)[t]) <= 1e-3, t
    assert rel_fro(gram.cpu().numpy(), g[f"{name}_gram"]) <= 1e-3
    sg = m.style_patch_grams(x).cpu().numpy()             # the style-transfer export takes feature_extractor only (no SE)
    want = pmg_oracle.style_patch_grams(x.float().cpu().numpy(), synth.pmg_state_dict(META["seed"], hp, TASKS), hp["patch_div"], norm=hp["norm"])
    assert rel_fro(sg, want) <= 1e-4
