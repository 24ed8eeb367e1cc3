"""Style-transfer Gram export: oracle vs the reference's own extract_patch_grams_for_image
(golden, CPU), CUDA path vs golden and oracle (GPU).  Reference: utils/PMG/PMG_style_transfer.py:11-48."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, build_model, load_contract, rel_fro, torch_state_dict
from oracle import pmg_oracle, synth

CASES = {"c224": "pmg_contract.json", "c160x208": "pmg_contract.json", "small60x68": "pmg_small_contract.json"}


def _golden(name):
    g = np.load(os.path.join(GOLDEN, "pmg_style_grams.npz"))
    seed, h, w, npatch, c = (int(v) for v in g[f"{name}_meta"])
    return g, seed, h, w, npatch, c


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_fixture(name):
    _, hp, tasks = load_contract(CASES[name])
    g, seed, h, w, npatch, c = _golden(name)
    x = synth.images(seed, 2, h, w)[1:2]
    got = pmg_oracle.style_patch_grams(x, synth.pmg_state_dict(seed, hp, tasks), hp["patch_div"])[0]
    assert got.shape == (npatch, c, c)
    assert rel_fro(got[list(g[f"{name}_keep"])], g[f"{name}_gram"]) <= 2e-5
    assert np.allclose(np.sqrt((got.astype(np.float64) ** 2).sum(axis=(1, 2))), g[f"{name}_fro"], rtol=2e-5)
    assert np.allclose(np.trace(got.astype(np.float64), axis1=1, axis2=2), g[f"{name}_trace"], rtol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("in_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_style_grams_vs_reference_fixture_and_oracle(name, in_dtype):
    _, hp, tasks = load_contract(CASES[name])
    g, seed, h, w, npatch, c = _golden(name)
    m = build_model(hp, tasks)
    sd = synth.pmg_state_dict(seed, hp, tasks)
    m.load_state_dict(torch_state_dict(sd), strict=True)
    m.cuda().eval()
    x = synth.images(seed, 2, h, w)
    got = m.style_patch_grams(torch.from_numpy(x).cuda().to(in_dtype)).cpu().numpy()
    assert got.shape == (2, npatch, c, c)
    assert rel_fro(got[1][list(g[f"{name}_keep"])], g[f"{name}_gram"]) <= 1e-3           # north_star gate; measured ~1e-6 (fp32 kernels)
    assert np.allclose(np.sqrt((got[1].astype(np.float64) ** 2).sum(axis=(1, 2))), g[f"{name}_fro"], rtol=1e-4)
    ref = pmg_oracle.style_patch_grams(x, sd, hp["patch_div"])
    assert rel_fro(got, ref) <= 1e-4
    assert (np.abs(got - got.transpose(0, 1, 3, 2)).max() <= 1e-6 * np.abs(got).max())   # exactly symmetric up to summation order
    assert m.style_patch_grams(torch.from_numpy(x[:0]).cuda()).shape == (0, npatch, c, c)
