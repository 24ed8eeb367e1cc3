#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``Models.models_PMG`` / ``utils.PMG.PMG_model_utils`` from /root/reference, loads the
seeded synthetic checkpoint of ``oracle/synth.py`` through the reference's own
``build_model_from_hparams`` + ``load_best_model`` path, runs ``model(x)`` in fp32 on CPU and stores
logits, Gram matrices and a few intermediates (captured with forward hooks) as small ``.npz`` files,
plus ``pmg_contract.json`` (hparams block, Tasks.json, the 81-key state_dict layout with shapes).
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PMG_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import synth  # noqa: E402
from utils.PMG.PMG_model_utils import build_model_from_hparams, load_best_model  # noqa: E402  (reference)

torch.set_num_threads(8)
torch.manual_seed(0)


def run_case(name, hparams, tasks_lists, seed, batch, h, w, keep_gram_images=(0,)):
    tasks_n = {t: len(v) for t, v in tasks_lists.items()}
    model = build_model_from_hparams(tasks_lists, hparams, torch.device("cpu"))
    sd_np = synth.pmg_state_dict(seed, hparams, tasks_n)
    assert list(model.state_dict().keys()) == list(sd_np.keys()), "synthetic key order != reference"
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "best_model_PMG.pth")
        torch.save({k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()}, p)
        model = load_best_model(model, p, torch.device("cpu"))
    x = torch.from_numpy(synth.images(seed, batch, h, w))
    cap = {}
    hooks = []
    n_fe = len(model.feature_extractor)
    for j in range(2, n_fe, 3):
        hooks.append(model.feature_extractor[j].register_forward_hook(
            lambda m, i, o, j=j: cap.__setitem__(f"act{j // 3}", o.detach().clone()) and None))
    hooks.append(model.chan_proj.register_forward_hook(lambda m, i, o: cap.__setitem__("chan", o.detach().clone()) and None))

    def _tp_hook(m, i, o):
        cap["gram_flat"] = i[0].detach().clone()
        cap["tokens_in"] = o.detach().clone()

    hooks.append(model.token_proj.register_forward_hook(_tp_hook))
    if not isinstance(model.transformer, torch.nn.Identity):
        hooks.append(model.transformer.register_forward_hook(lambda m, i, o: cap.__setitem__("tokens_out", o.detach().clone()) and None))
    with torch.no_grad():
        out = model(x)
    for hk in hooks:
        hk.remove()
    assert list(out.keys()) == list(tasks_lists.keys())
    logits = torch.cat([out[t] for t in tasks_lists], dim=1).numpy()
    Cr = hparams["gram_channels"]
    G = cap["gram_flat"].reshape(batch, -1, Cr, Cr).numpy()
    rec = {
        "logits": logits.astype(np.float32),
        "gram_images": np.array(keep_gram_images, dtype=np.int64),
        "gram": G[list(keep_gram_images)].astype(np.float32),
        "gram_fro": np.sqrt((G.astype(np.float64) ** 2).sum(axis=(1, 2, 3))),
        "gram_trace": np.trace(G.astype(np.float64), axis1=2, axis2=3),
        "tokens_in": cap["tokens_in"].numpy().astype(np.float32),
        "chan0": cap["chan"][0].numpy().astype(np.float32),
        "meta": np.array([seed, batch, h, w], dtype=np.int64),
    }
    if "tokens_out" in cap:
        rec["tokens_out"] = cap["tokens_out"].numpy().astype(np.float32)
    for k in sorted(c for c in cap if c.startswith("act")):
        a = cap[k].numpy().astype(np.float64)
        rec[f"{k}_stats"] = np.array([a.mean(), np.abs(a).mean(), np.sqrt((a ** 2).mean())])
        rec[f"{k}_img0_ch0"] = cap[k][0, 0].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **rec)
    print(f"{name}: logits {logits.shape} |logits|max {np.abs(logits).max():.3f} gram fro {rec['gram_fro'][:2]}")
    return model


def main():
    with open(os.path.join(REF, "Model_weights/PMG/hyperparameters_PMG.json")) as f:
        cfg = json.load(f)
    with open(os.path.join(REF, "build_classifiers/Tasks.json")) as f:
        tasks_lists = json.load(f)
    hp = cfg["hparams"]
    model = run_case("pmg_b8_224", hp, tasks_lists, seed=0, batch=8, h=224, w=224, keep_gram_images=(0, 7))
    contract = {
        "source": "generated from /root/reference Model_weights/PMG/hyperparameters_PMG.json:8-33,53-135 "
                  "and build_classifiers/Tasks.json by tests/golden/make_golden.py",
        "hparams": hp,
        "tasks": tasks_lists,
        "loaded_keys": cfg["pretrained_load_report"]["loaded"],
        "state_dict": [[k, list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()],
        "n_params": int(sum(p.numel() for p in model.parameters())),
    }
    assert contract["loaded_keys"] == [k for k, _, _ in contract["state_dict"]]
    with open(os.path.join(HERE, "pmg_contract.json"), "w") as f:
        json.dump(contract, f, indent=1)

    # ragged batch, same config
    run_case("pmg_b3_224", hp, tasks_lists, seed=3, batch=3, h=224, w=224)
    # non-224 input: 160×208 → 40×52 → 10×13 → replicate-pad to 12×15 → 3×3 patches of 4×5
    run_case("pmg_b3_160x208", hp, tasks_lists, seed=1, batch=3, h=160, w=208)
    # a different (small) configuration: one strided layer, patch_div 2, one encoder layer
    small_hp = {"patch_size": 16, "patch_div": 2, "ndf": 16, "norm": "batch", "gram_channels": 32,
                "d_model": 64, "transformer_layers": 1, "transformer_heads": 2,
                "use_token_attention": True, "use_channel_attention": False}
    small_tasks = {"alpha": ["a", "b", "c"], "beta": ["x", "y"]}
    run_case("pmg_small_b2_60x68", small_hp, small_tasks, seed=2, batch=2, h=60, w=68)
    with open(os.path.join(HERE, "pmg_small_contract.json"), "w") as f:
        json.dump({"hparams": small_hp, "tasks": small_tasks}, f, indent=1)


def run_pm_case(name, hp, tasks_n, seed, batch, h, w):
    from Models.models_PM import MultiTaskPatchGAN, load_model_weights   # reference
    model = MultiTaskPatchGAN(tasks_dict=tasks_n, input_nc=3, ndf=hp["ndf"], norm=hp["norm"], patch_size=hp["patch_size"],
                              device="cpu", attn_tau=hp["attn_tau"], attn_use_se=hp["attn_use_se"],
                              attn_softmax_spatial=hp["attn_softmax_spatial"])
    sd_np = synth.pm_state_dict(seed, hp, tasks_n)
    assert list(model.state_dict().keys()) == list(sd_np.keys()), "synthetic PM key order != reference"
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "best_model_PM.pth")
        torch.save({k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()}, p)
        load_model_weights(model, p, torch.device("cpu"), strict=True)
    model.eval()
    x = torch.from_numpy(synth.images(seed, batch, h, w))
    with torch.no_grad():
        full = model(x, return_full=True)
        feats = model.trunk(x)
    logits = torch.cat([full[t]["logits"] for t in tasks_n], dim=1).numpy()
    attn = torch.stack([full[t]["attn"][:, 0] for t in tasks_n], dim=1).numpy()        # (B, T, H, W)
    f = feats.numpy().astype(np.float64)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), logits=logits.astype(np.float32), attn=attn.astype(np.float32),
                        feats_stats=np.array([f.mean(), np.abs(f).mean(), np.sqrt((f ** 2).mean())]),
                        feats_img0_ch0=feats[0, 0].numpy().astype(np.float32),
                        feats_fro=np.sqrt((f ** 2).sum(axis=(1, 2, 3))), meta=np.array([seed, batch, h, w], dtype=np.int64))
    print(f"{name}: logits {logits.shape} |max| {np.abs(logits).max():.3f} feats {tuple(feats.shape)}")
    return model


def main_pm():
    with open(os.path.join(REF, "build_classifiers/Tasks.json")) as f:
        tasks_lists = json.load(f)
    tasks_n = {t: len(v) for t, v in tasks_lists.items()}
    with open(os.path.join(REF, "Model_weights/PM/hyperparameters_PM.json")) as f:
        cfg = json.load(f)
    hp = {"ndf": 64, "norm": "instance", "patch_size": cfg["patch_size"], "attn_tau": cfg["attn_tau"],
          "attn_use_se": cfg["attn_use_se"], "attn_softmax_spatial": cfg["attn_softmax_spatial"]}   # test_PM.py:146-164
    model = run_pm_case("pm_b2_224", hp, tasks_n, seed=4, batch=2, h=224, w=224)
    contract = {"source": "generated from /root/reference Model_weights/PM/hyperparameters_PM.json + test_PM.py:153-164 + Tasks.json",
                "hparams": hp, "tasks": tasks_lists,
                "state_dict": [[k, list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()],
                "n_params": int(sum(p.numel() for p in model.parameters()))}
    with open(os.path.join(HERE, "pm_contract.json"), "w") as f:
        json.dump(contract, f, indent=1)
    small = {"ndf": 32, "norm": "instance", "patch_size": 16, "attn_tau": 0.5, "attn_use_se": False, "attn_softmax_spatial": True}
    run_pm_case("pm_small_b3_40x52", small, {"alpha": 3, "beta": 2}, seed=6, batch=3, h=40, w=52)
    with open(os.path.join(HERE, "pm_small_contract.json"), "w") as f:
        json.dump({"hparams": small, "tasks": {"alpha": ["a", "b", "c"], "beta": ["x", "y"]}}, f, indent=1)


if __name__ == "__main__":
    if "--pm-only" not in sys.argv:
        main()
    main_pm()
    main()
