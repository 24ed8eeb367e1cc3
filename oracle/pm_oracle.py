"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference PM eval forward
(/root/reference/Models/models_PM.py: MultiTaskPatchGAN.forward :159-194, TaskHeadImproved.forward :54-76).

As for PMG the arithmetic lives in PyTorch (third-party, torch==2.8.0 pinned); the reference has no tests or
golden vectors for this path, so this oracle is pinned against outputs of the unmodified reference
(tests/golden/pm_*.npz, written by tests/golden/make_golden.py)."""
from __future__ import annotations

from typing import Dict

import numpy as np

IN_EPS = 1e-5      # nn.InstanceNorm2d default eps (models_PM.py:109-110)
LEAKY = 0.2        # models_PM.py:128,139


def conv_k4(x: np.ndarray, w: np.ndarray, b: np.ndarray, stride: int) -> np.ndarray:
    """nn.Conv2d(cin, cout, 4, stride, padding=1) with bias — models_PM.py:125,135."""
    B, C, H, W = x.shape
    Ho, Wo = (H + 2 - 4) // stride + 1, (W + 2 - 4) // stride + 1
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    cols = np.empty((B, Ho, Wo, C, 4, 4), dtype=x.dtype)
    for ky in range(4):
        for kx in range(4):
            cols[:, :, :, :, ky, kx] = xp[:, :, ky:ky + stride * Ho:stride, kx:kx + stride * Wo:stride].transpose(0, 2, 3, 1)
    o = cols.reshape(B * Ho * Wo, C * 16) @ w.reshape(w.shape[0], -1).T + b
    return o.reshape(B, Ho, Wo, w.shape[0]).transpose(0, 3, 1, 2)


def instance_norm(x, gamma, beta, eps=IN_EPS):
    """nn.InstanceNorm2d(affine=True), no running stats: per-(sample, channel) biased statistics over H×W."""
    mu = x.mean(axis=(2, 3), keepdims=True)
    var = ((x - mu) ** 2).mean(axis=(2, 3), keepdims=True)
    return (x - mu) / np.sqrt(var + x.dtype.type(eps)) * gamma.reshape(1, -1, 1, 1) + beta.reshape(1, -1, 1, 1)


def pm_trunk(x, sd):
    i = 0
    n_convs = sum(1 for k in sd if k.startswith("trunk.") and k.endswith(".weight") and sd[k].ndim == 4)
    for li in range(n_convs):
        stride = 2 if li < n_convs - 1 else 1                                          # models_PM.py:119,135
        x = conv_k4(x, sd[f"trunk.{i}.weight"], sd[f"trunk.{i}.bias"], stride)
        x = instance_norm(x, sd[f"trunk.{i + 1}.weight"], sd[f"trunk.{i + 1}.bias"])
        x = np.where(x >= 0, x, x * x.dtype.type(LEAKY))
        i += 3
    return x


def pm_forward(x: np.ndarray, sd: Dict[str, np.ndarray], tasks: Dict[str, int], tau: float = 0.7,
               ablate: bool = False, dtype=np.float32, return_full: bool = False):
    """MultiTaskPatchGAN.forward(x) / forward(x, return_full=True) with attn_softmax_spatial=True, attn_use_se=False."""
    sd = {k: v.astype(dtype) for k, v in sd.items()}
    feats = pm_trunk(x.astype(dtype), sd)                                              # :167
    N, C, H, W = feats.shape
    f = feats.reshape(N, C, H * W)
    out = {}
    for t in tasks:                                                                    # :191-193 → TaskHeadImproved :54-76
        wc = sd[f"task_heads.{t}.cls_conv.weight"].reshape(-1, C)
        M = np.einsum("kc,ncp->nkp", wc, f)                                            # cls_conv :72
        if ablate:
            A = np.full((N, 1, H * W), 1.0 / (H * W), dtype=dtype)                    # :60-63
            logits = (M * A).sum(-1)
        else:
            a = np.einsum("c,ncp->np", sd[f"task_heads.{t}.attn_conv.weight"].reshape(C), f) + sd[f"task_heads.{t}.attn_conv.bias"]
            a = a / dtype(tau)                                                         # :67
            e = np.exp(a - a.max(axis=-1, keepdims=True))
            A = (e / e.sum(-1, keepdims=True))[:, None, :]
            logits = (M * A).sum(-1) / (A.sum(-1) + dtype(1e-6))                       # :73-75
        out[t] = {"logits": logits, "attn": A.reshape(N, 1, H, W)} if return_full else logits
    return (out, feats) if return_full else out
