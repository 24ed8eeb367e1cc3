#!/usr/bin/env python
"""Golden fixtures for the style-transfer Gram export, produced by the UNMODIFIED reference:
``utils/PMG/PMG_style_transfer.py:extract_patch_grams_for_image`` on the reference model (CPU, fp32) with the seeded
synthetic checkpoint.  Build container only (needs /root/reference):

    python tests/golden/make_style_golden.py

Writes tests/golden/pmg_style_grams.npz: per case the full Gram of two patches plus Frobenius norm and trace of all."""
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
from utils.PMG.PMG_style_transfer import extract_patch_grams_for_image  # noqa: E402  (reference)

torch.set_num_threads(8)


def case(hp, tasks_lists, seed, h, w):
    tasks_n = {t: len(v) for t, v in tasks_lists.items()}
    model = build_model_from_hparams(tasks_lists, hp, torch.device("cpu"))
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.pmg_state_dict(seed, hp, tasks_n).items()}, strict=True)
    model.eval()
    x = torch.from_numpy(synth.images(seed, 2, h, w))
    grams = torch.stack(extract_patch_grams_for_image(model, x[1:2], detach=True)).numpy()      # image 1 of the seeded pair
    return grams


def main():
    with open(os.path.join(HERE, "pmg_contract.json")) as f:
        c = json.load(f)
    with open(os.path.join(HERE, "pmg_small_contract.json")) as f:
        cs = json.load(f)
    rec = {}
    for name, (hp, tl, seed, h, w) in {"c224": (c["hparams"], c["tasks"], 0, 224, 224), "c160x208": (c["hparams"], c["tasks"], 1, 160, 208),
                                       "small60x68": (cs["hparams"], cs["tasks"], 2, 60, 68)}.items():
        g = case(hp, tl, seed, h, w)
        keep = [0, g.shape[0] - 1]
        rec[f"{name}_meta"] = np.array([seed, h, w, g.shape[0], g.shape[1]], dtype=np.int64)
        rec[f"{name}_keep"] = np.array(keep, dtype=np.int64)
        rec[f"{name}_gram"] = g[keep].astype(np.float32)
        rec[f"{name}_fro"] = np.sqrt((g.astype(np.float64) ** 2).sum(axis=(1, 2)))
        rec[f"{name}_trace"] = np.trace(g.astype(np.float64), axis1=1, axis2=2)
        print(name, g.shape, rec[f"{name}_fro"][:3])
    np.savez_compressed(os.path.join(HERE, "pmg_style_grams.npz"), **rec)


if __name__ == "__main__":
    main()
