import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def rel_fro(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def load_contract(name="pmg_contract.json"):
    with open(os.path.join(GOLDEN, name)) as f:
        c = json.load(f)
    tasks = {t: len(v) for t, v in c["tasks"].items()}
    return c, c["hparams"], tasks


def build_model(hp, tasks, cls=None):
    import pmg_b200
    cls = cls or pmg_b200.MultiTaskPatchGANGramModelNonOverlapV2
    return cls(input_nc=3, ndf=hp["ndf"], norm=hp["norm"], patch_size=hp["patch_size"], patch_div=hp["patch_div"],
               num_classes_per_task=tasks, use_channel_attention=hp["use_channel_attention"],
               use_token_attention=hp["use_token_attention"], gram_channels=hp["gram_channels"],
               d_model=hp["d_model"], transformer_layers=hp["transformer_layers"],
               transformer_heads=hp["transformer_heads"])


def torch_state_dict(sd_np):
    import torch
    return {k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()}


def split_packed(packed, tasks):
    out, o = {}, 0
    for t, n in tasks.items():
        out[t] = packed[:, o:o + n]
        o += n
    return out


def argmax_table(packed, tasks):
    return np.stack([v.argmax(1) for v in split_packed(packed, tasks).values()], 1)
