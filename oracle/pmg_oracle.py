"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference PMG eval forward.

Reference: /root/reference/Models/models_PMG.py (class MultiTaskPatchGANGramModelNonOverlapV2).
The arithmetic itself lives in the third-party dependency PyTorch (reference pins torch==2.8.0,
requirements.txt:55; container has 2.11.0), not in the reference tree, so every function below
restates the published semantics of the torch op at the cited call site.

Parity pinning: the reference holds NO tests / golden vectors for this path.
This oracle is therefore pinned against outputs of the reference itself, produced by importing
the unmodified reference module in the build container: tests/golden/make_golden.py writes
tests/golden/*.npz, and tests/test_oracle_golden.py checks this file against them.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

BN_EPS = 1e-5      # nn.BatchNorm2d default, models_PMG.py:105
LN_EPS = 1e-5      # nn.TransformerEncoderLayer default layer_norm_eps, models_PMG.py:64-68
LEAKY = 0.2        # models_PMG.py:117,127


def _plan_from_state_dict(sd) -> List[Tuple[int, str]]:
    """feature_extractor conv indices in order: [(idx, 'strided'|'1x1')] (models_PMG.py:113-128)."""
    out = []
    i = 0
    while f"feature_extractor.{i}.weight" in sd:
        w = sd[f"feature_extractor.{i}.weight"]
        out.append((i, "strided" if w.shape[-1] == 4 else "1x1"))
        i += 3
    return out


def conv_k4s4(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """nn.Conv2d(cin, cout, kernel_size=4, stride=4, padding=0, bias=False) — models_PMG.py:115.

    Non-overlapping ⇒ a space-to-depth reshape followed by one GEMM; trailing rows/cols that do
    not fill a 4×4 window are dropped (floor), as torch does.
    """
    B, C, H, W = x.shape
    Ho, Wo = H // 4, W // 4
    xs = x[:, :, :Ho * 4, :Wo * 4].reshape(B, C, Ho, 4, Wo, 4)
    a = xs.transpose(0, 2, 4, 1, 3, 5).reshape(B * Ho * Wo, C * 16)          # (pixels, c·ky·kx)
    o = a @ w.reshape(w.shape[0], -1).T
    return o.reshape(B, Ho, Wo, w.shape[0]).transpose(0, 3, 1, 2)


def conv_1x1(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """nn.Conv2d(cin, cout, kernel_size=1, bias=False) — models_PMG.py:125,136."""
    B, C, H, W = x.shape
    o = x.transpose(0, 2, 3, 1).reshape(-1, C) @ w.reshape(w.shape[0], C).T
    return o.reshape(B, H, W, w.shape[0]).transpose(0, 3, 1, 2)


def batchnorm_eval(x, gamma, beta, mean, var, eps=BN_EPS):
    """nn.BatchNorm2d in eval(): running stats (models_PMG.py:116; eval set by PMG_model_utils.py:199)."""
    s = (1.0 / np.sqrt(var.astype(x.dtype) + x.dtype.type(eps))).reshape(1, -1, 1, 1)
    return (x - mean.reshape(1, -1, 1, 1)) * s * gamma.reshape(1, -1, 1, 1) + beta.reshape(1, -1, 1, 1)


def instance_norm(x, gamma, beta, eps=1e-5):
    """nn.InstanceNorm2d(C, affine=True) (models_PMG.py:101; no running stats: per-sample, per-channel statistics in
    train AND eval), biased variance."""
    mu = x.mean(axis=(2, 3), keepdims=True)
    var = ((x - mu) ** 2).mean(axis=(2, 3), keepdims=True)
    return (x - mu) / np.sqrt(var + x.dtype.type(eps)) * gamma.reshape(1, -1, 1, 1) + beta.reshape(1, -1, 1, 1)


def se_block(x, w1, w2):
    """SEBlock.forward (models_PMG.py:26-30): w = sigmoid(W2·relu(W1·avgpool(x))); x·w."""
    pooled = x.mean(axis=(2, 3))
    gate = 1.0 / (1.0 + np.exp(-(np.maximum(pooled @ w1.T, 0) @ w2.T)))
    return x * gate[:, :, None, None].astype(x.dtype)


def _norm(x, sd, b, norm):
    """norm_layer of the ctor (models_PMG.py:100-105; "group" cannot be constructed in the reference: TypeError at :116)."""
    if norm == "instance":
        return instance_norm(x, sd[f"{b}.weight"], sd[f"{b}.bias"])
    return batchnorm_eval(x, sd[f"{b}.weight"], sd[f"{b}.bias"], sd[f"{b}.running_mean"], sd[f"{b}.running_var"])


def leaky_relu(x, slope=LEAKY):
    return np.where(x >= 0, x, x * x.dtype.type(slope))


def split_patches(x: np.ndarray, patch_div: int) -> Tuple[np.ndarray, int]:
    """_pad_to_divisible + _split_patches — models_PMG.py:157-174.

    Replicate-pad bottom/right up to a multiple of patch_div, then cut patch_div² disjoint
    windows; patch k ↔ (k // patch_div, k % patch_div); inside a patch pixels are row-major.
    Returns (B, Np, C, N), N.
    """
    B, C, H, W = x.shape
    ph_pad, pw_pad = (-H) % patch_div, (-W) % patch_div
    if ph_pad or pw_pad:
        x = np.pad(x, ((0, 0), (0, 0), (0, ph_pad), (0, pw_pad)), mode="edge")
    H, W = x.shape[-2:]
    ph, pw = H // patch_div, W // patch_div
    p = x.reshape(B, C, patch_div, ph, patch_div, pw).transpose(0, 2, 4, 1, 3, 5)
    return np.ascontiguousarray(p.reshape(B, patch_div * patch_div, C, ph * pw)), ph * pw


def gram(patches: np.ndarray, n: int) -> np.ndarray:
    """G = P·Pᵀ / N — models_PMG.py:188."""
    return np.matmul(patches, patches.transpose(0, 1, 3, 2)) / patches.dtype.type(n)


def layer_norm(x, w, b, eps=LN_EPS):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + x.dtype.type(eps)) * w + b


def softmax(x, axis):
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)


def encoder_layer(x, sd, prefix: str, nhead: int):
    """nn.TransformerEncoderLayer(batch_first=True, norm_first=True, activation=relu) in eval —
    models_PMG.py:62-69.  x ← x + MHA(LN1(x));  x ← x + W2·relu(W1·LN2(x)+b1)+b2."""
    B, N, d = x.shape
    hd = d // nhead
    h = layer_norm(x, sd[f"{prefix}.norm1.weight"], sd[f"{prefix}.norm1.bias"])
    qkv = h @ sd[f"{prefix}.self_attn.in_proj_weight"].T + sd[f"{prefix}.self_attn.in_proj_bias"]
    q, k, v = (qkv[..., i * d:(i + 1) * d].reshape(B, N, nhead, hd).transpose(0, 2, 1, 3) for i in range(3))
    att = softmax(np.matmul(q, k.transpose(0, 1, 3, 2)) * x.dtype.type(1.0 / np.sqrt(hd)), axis=-1)
    o = np.matmul(att, v).transpose(0, 2, 1, 3).reshape(B, N, d)
    x = x + (o @ sd[f"{prefix}.self_attn.out_proj.weight"].T + sd[f"{prefix}.self_attn.out_proj.bias"])
    h = layer_norm(x, sd[f"{prefix}.norm2.weight"], sd[f"{prefix}.norm2.bias"])
    f = np.maximum(h @ sd[f"{prefix}.linear1.weight"].T + sd[f"{prefix}.linear1.bias"], 0)
    return x + (f @ sd[f"{prefix}.linear2.weight"].T + sd[f"{prefix}.linear2.bias"])


def task_pool_and_heads(tokens, sd, tasks: Dict[str, int]) -> Dict[str, np.ndarray]:
    """TaskAttentionPooling.forward (models_PMG.py:47-57, NO 1/sqrt(d) scale) + classifiers (:198)."""
    out = {}
    for t in tasks:
        q = sd[f"task_pool.query.{t}"]
        w = softmax((tokens * q).sum(-1), axis=1)[..., None]
        pooled = (w * tokens).sum(1)
        out[t] = pooled @ sd[f"classifiers.{t}.weight"].T + sd[f"classifiers.{t}.bias"]
    return out


def pmg_forward(x: np.ndarray, sd: Dict[str, np.ndarray], tasks: Dict[str, int], patch_div: int,
                nhead: int, dtype=np.float32, return_intermediates: bool = False, norm: str = "batch"):
    """MultiTaskPatchGANGramModelNonOverlapV2.forward — models_PMG.py:176-199.

    ``sd`` is a numpy state_dict in the reference's key layout. ``dtype=np.float64`` gives a
    higher-precision witness used to rank fp32 reference vs CUDA errors.
    """
    sd = {k: (v.astype(dtype) if v.dtype.kind == "f" else v) for k, v in sd.items()}
    x = x.astype(dtype)
    inter = {}
    for (i, kind) in _plan_from_state_dict(sd):                                   # :178
        w = sd[f"feature_extractor.{i}.weight"]
        x = conv_k4s4(x, w) if kind == "strided" else conv_1x1(x, w)
        b = f"feature_extractor.{i + 1}"
        x = leaky_relu(_norm(x, sd, b, norm))
        inter[f"act{i // 3}"] = x
    if "se.fc.0.weight" in sd:                                                     # :133,179 (use_channel_attention)
        x = se_block(x, sd["se.fc.0.weight"], sd["se.fc.2.weight"])
    x = conv_1x1(x, sd["chan_proj.weight"])                                        # :182
    inter["chan"] = x
    patches, n = split_patches(x, patch_div)                                       # :185
    G = gram(patches, n)                                                           # :188
    inter["gram"] = G
    B, Np, Cr, _ = G.shape
    tokens = G.reshape(B, Np, Cr * Cr) @ sd["token_proj.weight"].T + sd["token_proj.bias"]   # :192-193
    inter["tokens_in"] = tokens
    l = 0
    while f"transformer.encoder.layers.{l}.norm1.weight" in sd:                    # :194
        tokens = encoder_layer(tokens, sd, f"transformer.encoder.layers.{l}", nhead)
        l += 1
    inter["tokens_out"] = tokens
    logits = task_pool_and_heads(tokens, sd, tasks)                                # :197-198
    if return_intermediates:
        return logits, inter
    return logits


def style_patch_grams(x: np.ndarray, sd: Dict[str, np.ndarray], patch_div: int, dtype=np.float32, norm: str = "batch") -> np.ndarray:
    """extract_patch_grams_for_image — utils/PMG/PMG_style_transfer.py:11-48, for every image of the batch.

    Features = ``model.feature_extractor(x)`` (:21; the conv/BN/LeakyReLU stack only, C = last conv width — NOT chan_proj);
    ``patch_h = H // patch_div`` (:29-30); ``feats.unfold(2, ph, ph).unfold(3, pw, pw)`` (:32) keeps ``nH = H // ph`` x
    ``nW = W // pw`` windows WITHOUT padding (trailing rows / cols that do not fill a window are dropped).  Then (:37)
    ``permute(0, 1, 2, 4, 3, 5)`` puts the axes back into IMAGE order (nH, ph, nW, pw) and ``reshape(1, C, nb, ph, pw)``
    re-chunks that row-major image: what the reference calls patch k is the run of ``ph·pw`` CONSECUTIVE pixels
    ``[k·N, (k+1)·N)`` of the cropped (nH·ph) x (nW·pw) map flattened row-major — not the k-th spatial window.  The
    reference's results are the contract, so that is what is restated (and what the CUDA export computes).
    ``G_k = F_k @ F_k.T / (ph·pw)`` (:44-45).  Returns (B, nH·nW, C, C)."""
    sd = {k: (v.astype(dtype) if v.dtype.kind == "f" else v) for k, v in sd.items()}
    x = x.astype(dtype)
    for (i, kind) in _plan_from_state_dict(sd):
        w = sd[f"feature_extractor.{i}.weight"]
        x = conv_k4s4(x, w) if kind == "strided" else conv_1x1(x, w)
        b = f"feature_extractor.{i + 1}"
        x = leaky_relu(_norm(x, sd, b, norm))
    B, C, H, W = x.shape
    ph, pw = H // patch_div, W // patch_div
    ny, nx = H // ph, W // pw
    flat = np.ascontiguousarray(x[:, :, :ny * ph, :nx * pw]).reshape(B, C, ny * nx, ph * pw)
    p = flat.transpose(0, 2, 1, 3)
    return np.matmul(p, p.transpose(0, 1, 3, 2)) / dtype(ph * pw)


def pack_logits(logits: Dict[str, np.ndarray], tasks: Dict[str, int]) -> np.ndarray:
    """(B, Σncls) in Tasks.json order — the layout the C-ABI writes."""
    return np.concatenate([logits[t] for t in tasks], axis=1)
