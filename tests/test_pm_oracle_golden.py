"""numpy PM oracle (oracle/pm_oracle.py) pinned against the unmodified reference (tests/golden/pm_*.npz)."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, rel_fro
from oracle import pm_oracle, synth

CASES = [("pm_b2_224", "pm_contract.json"), ("pm_small_b3_40x52", "pm_small_contract.json")]


def load_pm_case(name, contract):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    with open(os.path.join(GOLDEN, contract)) as f:
        c = json.load(f)
    tasks = {t: len(v) for t, v in c["tasks"].items()}
    seed, batch, h, w = (int(v) for v in g["meta"])
    return g, c["hparams"], tasks, seed, batch, h, w


@pytest.mark.parametrize("name,contract", CASES)
def test_pm_oracle_matches_reference_golden(name, contract):
    g, hp, tasks, seed, batch, h, w = load_pm_case(name, contract)
    sd = synth.pm_state_dict(seed, hp, tasks)
    x = synth.images(seed, batch, h, w)
    out, feats = pm_oracle.pm_forward(x, sd, tasks, tau=hp["attn_tau"], return_full=True)
    logits = np.concatenate([out[t]["logits"] for t in tasks], 1)
    attn = np.stack([out[t]["attn"][:, 0] for t in tasks], 1)
    assert rel_fro(logits, g["logits"]) < 2e-5
    assert rel_fro(attn, g["attn"]) < 2e-5
    assert rel_fro(feats[0, 0], g["feats_img0_ch0"]) < 2e-5
    assert np.allclose(np.sqrt((feats.astype(np.float64) ** 2).sum(axis=(1, 2, 3))), g["feats_fro"], rtol=1e-5)


def test_pm_contract_matches_synth():
    with open(os.path.join(GOLDEN, "pm_contract.json")) as f:
        c = json.load(f)
    sd = synth.pm_state_dict(0)
    assert [k for k, _, _ in c["state_dict"]] == list(sd.keys())
    for k, shape, _ in c["state_dict"]:
        assert list(sd[k].shape) == shape, k
    assert c["n_params"] == 11_215_692
