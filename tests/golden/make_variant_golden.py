#!/usr/bin/env python
"""Golden fixtures for the constructor variants the committed config does not use (norm = instance,
use_channel_attention = SEBlock — Models/models_PMG.py:15-30,100-101,133; norm = group raises TypeError in the reference), produced by the UNMODIFIED reference module
(CPU, fp32) on a small configuration with the seeded synthetic weights.  Build container only:

    python tests/golden/make_variant_golden.py        -> tests/golden/pmg_variants.npz (+ pmg_variants.json)"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PMG_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import synth  # noqa: E402
from utils.PMG.PMG_model_utils import build_model_from_hparams  # noqa: E402  (reference)

torch.set_num_threads(8)
BASE = {"patch_size": 64, "patch_div": 2, "ndf": 32, "gram_channels": 48, "d_model": 64, "transformer_layers": 1,
        "transformer_heads": 2, "use_token_attention": True}
VARIANTS = {"instance": {"norm": "instance", "use_channel_attention": False},
            "batch_se": {"norm": "batch", "use_channel_attention": True},
            "instance_se": {"norm": "instance", "use_channel_attention": True}}
TASKS = {"alpha": ["a", "b", "c"], "beta": ["x", "y"], "gamma": ["p", "q", "r", "s"]}


def main():
    rec, meta = {}, {"base": BASE, "variants": VARIANTS, "tasks": TASKS, "input": [3, 96, 112], "seed": 4}
    for name, v in VARIANTS.items():
        hp = dict(BASE, **v)
        tasks_n = {t: len(c) for t, c in TASKS.items()}
        model = build_model_from_hparams(TASKS, hp, torch.device("cpu"))
        sd_np = synth.pmg_state_dict(4, hp, tasks_n)
        assert list(model.state_dict().keys()) == list(sd_np.keys()), (name, "synthetic key order != reference")
        model.load_state_dict({k: torch.from_numpy(np.array(val)) for k, val in sd_np.items()}, strict=True)
        model.eval()
        x = torch.from_numpy(synth.images(4, 3, 96, 112))
        cap = {}
        h = model.token_proj.register_forward_hook(lambda m, i, o: cap.__setitem__("gram", i[0].detach().clone()) and None)
        with torch.no_grad():
            out = model(x)
        h.remove()
        rec[f"{name}_logits"] = torch.cat([out[t] for t in TASKS], 1).numpy().astype(np.float32)
        rec[f"{name}_gram"] = cap["gram"].reshape(3, 4, 48, 48).numpy().astype(np.float32)
        meta[f"{name}_keys"] = [[k, list(val.shape)] for k, val in model.state_dict().items()]
        print(name, rec[f"{name}_logits"][0, :5])
    np.savez_compressed(os.path.join(HERE, "pmg_variants.npz"), **rec)
    with open(os.path.join(HERE, "pmg_variants.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
