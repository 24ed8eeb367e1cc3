"""TEST INFRASTRUCTURE ONLY — deterministic synthetic weights / images for the PMG + PM paths.

The committed checkpoints (``best_model_PMG.pth`` / ``best_model_PM.pth``) are absent from the
reference mount (``/root/reference/.MISSING_LARGE_BLOBS:1-2``), so parity is proven on seeded
synthetic weights.  Everything here is generated with a counter-based integer hash (splitmix64)
and exact float arithmetic, so the very same numbers come out on every machine / numpy version —
the golden fixtures under ``tests/golden/`` depend on that.

All floating values are rounded to bf16-representable fp32: the CUDA path holds
GEMM weights and the image in bf16, the oracle runs fp32 on *identical* values.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on a uint64 array (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        z = z ^ (z >> np.uint64(31))
    return z


def _stream_id(name: str, seed: int) -> int:
    """FNV-1a of the tensor name mixed with the seed (pure Python ints: no platform dependence)."""
    h = 0xCBF29CE484222325
    for b in name.encode("utf-8"):
        h ^= b
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return (h ^ (seed * 0xD6E8FEB86659FD93)) & 0xFFFFFFFFFFFFFFFF


def uniform01(name: str, seed: int, n: int) -> np.ndarray:
    """n doubles in [0,1) with 32 random bits each — exact in float64."""
    with np.errstate(over="ignore"):
        ctr = (np.arange(n, dtype=np.uint64) + np.uint64(_stream_id(name, seed))) & _M64
    bits = _splitmix64(_splitmix64(ctr)) >> np.uint64(32)
    return bits.astype(np.float64) / 4294967296.0


def normal(name: str, seed: int, n: int) -> np.ndarray:
    """Approximate N(0,1): Irwin–Hall sum of 12 uniforms − 6 (exact float64 arithmetic, no libm)."""
    acc = np.zeros(n, dtype=np.float64)
    for j in range(12):
        acc += uniform01(f"{name}#{j}", seed, n)
    return acc - 6.0


def round_bf16(a: np.ndarray) -> np.ndarray:
    """Round fp32 → nearest-even bf16 → back to fp32 (bit arithmetic; matches torch's cast)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) >> np.uint64(16)) << np.uint64(16)
    return (r & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.float32).reshape(a.shape)


def _t(name, seed, shape, std=1.0, kind="normal", lo=0.0, hi=1.0):
    n = int(np.prod(shape)) if len(shape) else 1
    if kind == "normal":
        v = normal(name, seed, n) * std
    else:
        v = lo + (hi - lo) * uniform01(name, seed, n)
    return round_bf16(v.astype(np.float32).reshape(shape))


# --------------------------------------------------------------------------------------
# PMG contract (reference: Model_weights/PMG/hyperparameters_PMG.json:8-33, build_classifiers/Tasks.json)
# --------------------------------------------------------------------------------------
PMG_HPARAMS = {
    "patch_size": 64, "patch_div": 3, "ndf": 32, "norm": "batch", "gram_channels": 96,
    "d_model": 192, "transformer_layers": 2, "transformer_heads": 4,
    "use_token_attention": True, "use_channel_attention": False,
}
PMG_TASKS = {
    "Weather Type": 8, "Weather Intensity": 4, "Visibility": 4, "Sky Condition": 6,
    "Precipitation Presence": 4, "Precipitation Intensity": 4, "Ground Condition": 7,
    "Glare or Reflections": 2, "Light Conditions": 5, "Road Spray": 2,
    "Water On Window": 3, "Snow On Window": 3,
}


def pmg_backbone_plan(input_nc: int, ndf: int, patch_size: int):
    """Layer-count loop of the reference ctor (Models/models_PMG.py:107-128).

    Returns [(cin, cout, kernel)] — strided k=s=4 convs followed by the final 1×1.
    """
    plan, cin, nf, rf = [], input_nc, ndf, float(patch_size)
    while rf > 4 and nf <= 512:
        plan.append((cin, nf, 4))
        cin, nf, rf = nf, nf * 2, rf / 4
    plan.append((cin, nf, 1))
    return plan


def pmg_state_dict(seed: int = 0, hparams: dict | None = None, tasks: dict | None = None,
                   input_nc: int = 3) -> "dict[str, np.ndarray]":
    """Synthetic PMG state_dict, key set / order / shapes of hyperparameters_PMG.json:53-135.

    BN running stats and affine are randomised (default init would make BN ≈ identity and hide
    folding bugs — SURVEY.md §7 step 0).
    """
    hp = dict(PMG_HPARAMS if hparams is None else hparams)
    tasks = dict(PMG_TASKS if tasks is None else tasks)
    d, cr, nl = hp["d_model"], hp["gram_channels"], hp["transformer_layers"]
    sd: dict[str, np.ndarray] = {}
    idx = 0
    for (cin, cout, k) in pmg_backbone_plan(input_nc, hp["ndf"], hp["patch_size"]):
        fan = cin * k * k
        sd[f"feature_extractor.{idx}.weight"] = _t(f"fe{idx}.w", seed, (cout, cin, k, k), std=1.4 / np.sqrt(fan))
        b = f"feature_extractor.{idx + 1}"
        sd[f"{b}.weight"] = _t(f"{b}.g", seed, (cout,), kind="uniform", lo=0.6, hi=1.4)
        sd[f"{b}.bias"] = _t(f"{b}.b", seed, (cout,), std=0.15)
        if hp.get("norm", "batch") != "instance":                       # InstanceNorm2d(affine=True) keeps no statistics
            sd[f"{b}.running_mean"] = _t(f"{b}.m", seed, (cout,), std=0.25)
            sd[f"{b}.running_var"] = _t(f"{b}.v", seed, (cout,), kind="uniform", lo=0.5, hi=1.5)
            sd[f"{b}.num_batches_tracked"] = np.array(1234, dtype=np.int64)
        idx += 3
        c_last = cout
    if hp.get("use_channel_attention", False):                          # SEBlock, models_PMG.py:15-30
        hid = max(1, c_last // 16)
        sd["se.fc.0.weight"] = _t("se.w1", seed, (hid, c_last), std=1.5 / np.sqrt(c_last))
        sd["se.fc.2.weight"] = _t("se.w2", seed, (c_last, hid), std=1.5 / np.sqrt(hid))
    sd["chan_proj.weight"] = _t("chan_proj.w", seed, (cr, c_last, 1, 1), std=1.0 / np.sqrt(c_last))
    sd["token_proj.weight"] = _t("token_proj.w", seed, (d, cr * cr), std=1.0 / np.sqrt(cr * cr) * 4.0)
    sd["token_proj.bias"] = _t("token_proj.b", seed, (d,), std=0.1)
    if hp.get("use_token_attention", True) and nl > 0:
        for l in range(nl):
            p = f"transformer.encoder.layers.{l}"
            sd[f"{p}.self_attn.in_proj_weight"] = _t(f"{p}.inw", seed, (3 * d, d), std=1.0 / np.sqrt(d))
            sd[f"{p}.self_attn.in_proj_bias"] = _t(f"{p}.inb", seed, (3 * d,), std=0.05)
            sd[f"{p}.self_attn.out_proj.weight"] = _t(f"{p}.ow", seed, (d, d), std=1.0 / np.sqrt(d))
            sd[f"{p}.self_attn.out_proj.bias"] = _t(f"{p}.ob", seed, (d,), std=0.05)
            sd[f"{p}.linear1.weight"] = _t(f"{p}.l1w", seed, (2 * d, d), std=1.0 / np.sqrt(d))
            sd[f"{p}.linear1.bias"] = _t(f"{p}.l1b", seed, (2 * d,), std=0.05)
            sd[f"{p}.linear2.weight"] = _t(f"{p}.l2w", seed, (d, 2 * d), std=1.0 / np.sqrt(2 * d))
            sd[f"{p}.linear2.bias"] = _t(f"{p}.l2b", seed, (d,), std=0.05)
            sd[f"{p}.norm1.weight"] = _t(f"{p}.n1w", seed, (d,), kind="uniform", lo=0.7, hi=1.3)
            sd[f"{p}.norm1.bias"] = _t(f"{p}.n1b", seed, (d,), std=0.1)
            sd[f"{p}.norm2.weight"] = _t(f"{p}.n2w", seed, (d,), kind="uniform", lo=0.7, hi=1.3)
            sd[f"{p}.norm2.bias"] = _t(f"{p}.n2b", seed, (d,), std=0.1)
    # nn.ParameterDict keeps insertion order in recent torch but the committed key list
    # (hyperparameters_PMG.json:97-108) is alphabetical; the module decides the order, values
    # are keyed by name here so either order loads.
    for t in sorted(tasks):
        sd[f"task_pool.query.{t}"] = _t(f"q.{t}", seed, (d,), std=0.35)
    for t, ncls in tasks.items():
        sd[f"classifiers.{t}.weight"] = _t(f"cls.{t}.w", seed, (ncls, d), std=2.0 / np.sqrt(d))
        sd[f"classifiers.{t}.bias"] = _t(f"cls.{t}.b", seed, (ncls,), std=0.2)
    return sd


def images(seed: int, batch: int, h: int = 224, w: int = 224, c: int = 3) -> np.ndarray:
    """ImageNet-normalised-looking synthetic images, bf16-representable fp32, NCHW.

    A smooth low-frequency field plus pixel noise, so neighbouring pixels correlate like a
    photograph does (pure noise would make every Gram matrix look alike).
    """
    n = batch * c * h * w
    noise = normal("img.noise", seed, n).reshape(batch, c, h, w)
    gh, gw = max(2, h // 16), max(2, w // 16)
    coarse = normal("img.coarse", seed, batch * c * gh * gw).reshape(batch, c, gh, gw)
    yy = (np.arange(h) * gh // h)
    xx = (np.arange(w) * gw // w)
    field = coarse[:, :, yy][:, :, :, xx]
    return round_bf16((1.1 * field + 0.6 * noise).astype(np.float32))


# --------------------------------------------------------------------------------------
# PM contract (reference: Models/models_PM.py:91-157, test_PM.py:153-164, hyperparameters_PM.json)
# --------------------------------------------------------------------------------------
PM_HPARAMS = {"ndf": 64, "norm": "instance", "patch_size": 73, "attn_tau": 0.7, "attn_use_se": False,
              "attn_softmax_spatial": True}


def pm_trunk_plan(input_nc: int, ndf: int, patch_size: float):
    """[(cin, cout, stride)] of the k=4, pad=1 trunk convs — layer-count loop of models_PM.py:121-139."""
    plan, cin, nf, rf = [], input_nc, ndf, float(patch_size)
    while rf > 4 and nf <= 512:
        plan.append((cin, nf, 2))
        cin, nf, rf = nf, nf * 2, rf / 2
    plan.append((cin, nf, 1))
    return plan


def pm_state_dict(seed: int = 0, hparams: dict | None = None, tasks: dict | None = None, input_nc: int = 3):
    """Synthetic PM state_dict in the reference's key layout (InstanceNorm2d(affine=True) has no running stats)."""
    hp = dict(PM_HPARAMS if hparams is None else hparams)
    tasks = dict(PMG_TASKS if tasks is None else tasks)
    sd = {}
    idx = 0
    for (cin, cout, _s) in pm_trunk_plan(input_nc, hp["ndf"], hp["patch_size"]):
        fan = cin * 16
        sd[f"trunk.{idx}.weight"] = _t(f"pm.t{idx}.w", seed, (cout, cin, 4, 4), std=1.4 / np.sqrt(fan))
        sd[f"trunk.{idx}.bias"] = _t(f"pm.t{idx}.b", seed, (cout,), std=0.1)
        sd[f"trunk.{idx + 1}.weight"] = _t(f"pm.t{idx}.g", seed, (cout,), kind="uniform", lo=0.6, hi=1.4)
        sd[f"trunk.{idx + 1}.bias"] = _t(f"pm.t{idx}.be", seed, (cout,), std=0.2)
        idx += 3
        c_last = cout
    for t, n in tasks.items():
        sd[f"task_heads.{t}.attn_conv.weight"] = _t(f"pm.{t}.aw", seed, (1, c_last, 1, 1), std=2.0 / np.sqrt(c_last))
        sd[f"task_heads.{t}.attn_conv.bias"] = _t(f"pm.{t}.ab", seed, (1,), std=0.1)
        sd[f"task_heads.{t}.cls_conv.weight"] = _t(f"pm.{t}.cw", seed, (n, c_last, 1, 1), std=1.0 / np.sqrt(c_last))
    return sd
