"""The numpy oracle (oracle/pmg_oracle.py) pinned against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import pmg_oracle, synth


def rel_fro(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _load_case(golden_dir, name, contract):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    with open(os.path.join(golden_dir, contract)) as f:
        c = json.load(f)
    tasks = {t: len(v) for t, v in c["tasks"].items()}
    seed, batch, h, w = (int(v) for v in g["meta"])
    return g, c["hparams"], tasks, seed, batch, h, w


CASES = [("pmg_b8_224", "pmg_contract.json"), ("pmg_b3_224", "pmg_contract.json"),
         ("pmg_b3_160x208", "pmg_contract.json"), ("pmg_small_b2_60x68", "pmg_small_contract.json")]


@pytest.mark.parametrize("name,contract", CASES)
def test_oracle_matches_reference_golden(golden_dir, name, contract):
    g, hp, tasks, seed, batch, h, w = _load_case(golden_dir, name, contract)
    sd = synth.pmg_state_dict(seed, hp, tasks)
    x = synth.images(seed, batch, h, w)
    logits, inter = pmg_oracle.pmg_forward(x, sd, tasks, hp["patch_div"], hp["transformer_heads"],
                                           return_intermediates=True)
    packed = pmg_oracle.pack_logits(logits, tasks)
    # fp32 reference (MKL-DNN) vs fp32 numpy: summation order differs, nothing else
    assert rel_fro(packed, g["logits"]) < 2e-5
    assert np.array_equal(_argmax(packed, tasks), _argmax(g["logits"], tasks))
    assert rel_fro(inter["gram"][list(g["gram_images"])], g["gram"]) < 1e-5
    gf = np.sqrt((inter["gram"].astype(np.float64) ** 2).sum(axis=(1, 2, 3)))
    assert np.allclose(gf, g["gram_fro"], rtol=1e-5)
    assert np.allclose(np.trace(inter["gram"].astype(np.float64), axis1=2, axis2=3), g["gram_trace"], rtol=1e-5)
    assert rel_fro(inter["tokens_in"], g["tokens_in"]) < 1e-5
    assert rel_fro(inter["chan"][0], g["chan0"]) < 1e-5
    if "tokens_out" in g:
        assert rel_fro(inter["tokens_out"], g["tokens_out"]) < 2e-5
    for k in [k for k in g.files if k.endswith("_img0_ch0")]:
        assert rel_fro(inter[k[:-9]][0, 0], g[k]) < 1e-5


def _argmax(packed, tasks):
    out, o = [], 0
    for t, n in tasks.items():
        out.append(packed[:, o:o + n].argmax(1))
        o += n
    return np.stack(out, 1)


def test_oracle_fp64_witness_agrees(golden_dir):
    """fp64 oracle vs fp32 reference: bounds the reference's own rounding noise (≪ the 1e-3 gate)."""
    g, hp, tasks, seed, batch, h, w = _load_case(golden_dir, "pmg_b3_224", "pmg_contract.json")
    sd = synth.pmg_state_dict(seed, hp, tasks)
    x = synth.images(seed, batch, h, w)
    logits = pmg_oracle.pmg_forward(x, sd, tasks, hp["patch_div"], hp["transformer_heads"], dtype=np.float64)
    assert rel_fro(pmg_oracle.pack_logits(logits, tasks), g["logits"]) < 2e-5


def test_synth_is_bf16_representable_and_stable():
    sd = synth.pmg_state_dict(0)
    w = sd["token_proj.weight"]
    assert np.array_equal(synth.round_bf16(w), w)
    x = synth.images(0, 1)
    assert np.array_equal(synth.round_bf16(x), x)
    # pinned values: the generator must never drift (golden fixtures depend on it)
    assert float(x[0, 0, 0, 0]) == pytest.approx(float(synth.images(0, 1)[0, 0, 0, 0]))
    assert len(sd) == 81


def test_contract_matches_synth(golden_dir):
    with open(os.path.join(golden_dir, "pmg_contract.json")) as f:
        c = json.load(f)
    sd = synth.pmg_state_dict(0)
    assert [k for k, _, _ in c["state_dict"]] == list(sd.keys()) == c["loaded_keys"]
    for k, shape, _ in c["state_dict"]:
        assert list(sd[k].shape) == shape, k
    assert c["hparams"]["num_classes_per_task"] == synth.PMG_TASKS
    assert {k: c["hparams"][k] for k in synth.PMG_HPARAMS} == synth.PMG_HPARAMS


def test_split_patches_replicate_and_order():
    x = np.arange(2 * 3 * 14 * 14, dtype=np.float32).reshape(2, 3, 14, 14)
    p, n = pmg_oracle.split_patches(x, 3)
    assert p.shape == (2, 9, 3, 25) and n == 25
    # patch 8 = bottom-right; its last row/col duplicate row/col 13 (replicate pad)
    last = p[0, 8, 0].reshape(5, 5)
    assert np.array_equal(last[4], last[3]) and np.array_equal(last[:, 4], last[:, 3])
    assert p[0, 1, 0, 0] == x[0, 0, 0, 5]       # patch k=1 ↔ (row 0, col 1)


@pytest.mark.parametrize("name,contract", CASES)
def test_torch_port_matches_reference_golden(golden_dir, name, contract):
    """oracle/pmg_torch_port.py (the timed CPU baseline) against the unmodified reference's outputs."""
    import torch
    from oracle.pmg_torch_port import PMGTorchPort
    g, hp, tasks, seed, batch, h, w = _load_case(golden_dir, name, contract)
    port = PMGTorchPort(synth.pmg_state_dict(seed, hp, tasks), tasks, hp["patch_div"], hp["transformer_heads"])
    out = port(torch.from_numpy(synth.images(seed, batch, h, w)))
    packed = torch.cat([out[t] for t in tasks], 1).numpy()
    assert rel_fro(packed, g["logits"]) < 1e-5
