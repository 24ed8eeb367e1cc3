"""GPU parity at the BENCH configuration and over every switchable code path of the fast forward.

* B = 1024 (BASELINE.json configs[2], what bench.py times) bf16 AND device-resident fp32 input, compared image by
  image with the CPU port of the reference's op sequence (oracle/pmg_torch_port.py, itself pinned to the golden
  fixtures of the unmodified reference in tests/test_oracle_golden.py): per-task argmax identical on EVERY image,
  per-task logits within 1e-3 relative (Frobenius), every single image's 52-logit row within 1e-3 — so the
  persistent-CTA tails of the front / midgram / token-stage kernels (2048 / 3072 / 147 work items on 148 SMs) are
  covered, not only the first tile.
* every PMG_FLAG_NO_FUSED_* path (the kernels behind models_PMG.fused_front / fused_mid / fused_midgram /
  fused_encoder) with bf16 input against the golden fixtures of the reference.
Reference lines: Models/models_PMG.py:176-199."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, argmax_table, build_model, load_contract, rel_fro, split_packed, torch_state_dict
from oracle import synth
from oracle.pmg_torch_port import PMGTorchPort

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _model(seed=0):
    _, hp, tasks = load_contract()
    m = build_model(hp, tasks)
    m.load_state_dict(torch_state_dict(synth.pmg_state_dict(seed, hp, tasks)), strict=True)
    m.to("cuda").eval()
    return m, hp, tasks


def _photo_like(batch, seed):
    """Low-frequency field + pixel noise (the recipe of synth.images, drawn with torch's CPU generator so that 1024
    images take a second instead of minutes), rounded to bf16-representable fp32 (north_star: bf16 inputs)."""
    g = torch.Generator().manual_seed(seed)
    noise = torch.randn(batch, 3, 224, 224, generator=g)
    coarse = torch.randn(batch, 3, 14, 14, generator=g)
    x = 1.1 * torch.nn.functional.interpolate(coarse, scale_factor=16, mode="nearest") + 0.6 * noise
    return x.to(torch.bfloat16).float()


@pytest.mark.parametrize("in_dtype", [torch.bfloat16, torch.float32])
def test_bench_batch_every_image_vs_reference_port(in_dtype):
    m, hp, tasks = _model(0)
    B = 1024
    x = _photo_like(B, 1234)
    port = PMGTorchPort(synth.pmg_state_dict(0, hp, tasks), tasks, hp["patch_div"], hp["transformer_heads"])
    ref = torch.cat([port(x)[t] for t in tasks], 1).numpy()
    with torch.no_grad():
        got = m.forward_packed(x.cuda().to(in_dtype)).cpu().numpy()
    # front (+ the hi/lo plane split for fp32 input), conv2+chan_proj+Gram, token_proj x2, token stage
    assert m.launches_per_forward(B, dtype=in_dtype) == (4 if in_dtype == torch.bfloat16 else 5)
    assert np.isfinite(got).all()
    # (1) per-task argmax identical on every one of the 1024 x 12 decisions
    bad = np.argwhere(argmax_table(got, tasks) != argmax_table(ref, tasks))
    assert bad.size == 0, f"argmax differs at (image, task) {bad[:8].tolist()}"
    # (2) per-task relative error
    worst = max(rel_fro(v, split_packed(ref, tasks)[t]) for t, v in split_packed(got, tasks).items())
    assert worst <= TOL, worst
    # (3) every image on its own (a wrong tail tile cannot hide inside a Frobenius norm over 1024 images)
    row = np.linalg.norm(got.astype(np.float64) - ref, axis=1) / np.linalg.norm(ref.astype(np.float64), axis=1)
    assert row.max() <= TOL, (int(row.argmax()), float(row.max()))
    # margin the DESIGN quotes: the hi/lo planes sit two orders below the gate
    assert worst <= 2e-4


FLAG_SETS = [
    dict(fused_front=False),
    dict(fused_mid=False, fused_midgram=False),
    dict(fused_midgram=False),
    dict(fused_midgram="always"),
    dict(fused_encoder=False),
    dict(fused_front=False, fused_mid=False, fused_midgram=False, fused_encoder=False),
]


@pytest.mark.parametrize("flags", FLAG_SETS, ids=lambda f: "+".join(f"{k}={v}" for k, v in f.items()))
def test_every_flag_path_against_golden_and_default(flags):
    """Each unfused kernel chain (im2col + conv0/conv1 GEMMs, conv2 / chan_proj GEMMs, mid kernel + Gram kernel, per-op
    transformer kernels + heads GEMM + pool_finish) with bf16 input: golden fixtures of the reference, then a ragged
    150-image batch (tails of every persistent kernel) against the default path."""
    m, hp, tasks = _model(0)
    g = np.load(os.path.join(GOLDEN, "pmg_b8_224.npz"))
    seed, batch, h, w = (int(v) for v in g["meta"])
    x = torch.from_numpy(synth.images(seed, batch, h, w)).cuda().to(torch.bfloat16)
    for k, v in flags.items():
        setattr(m, k, v)
    with torch.no_grad():
        packed, gram, taps = m.forward_debug(x)
    packed, gram = packed.cpu().numpy(), gram.cpu().numpy()
    assert np.array_equal(argmax_table(packed, tasks), argmax_table(g["logits"], tasks))
    for t, v in split_packed(packed, tasks).items():
        assert rel_fro(v, split_packed(g["logits"], tasks)[t]) <= TOL, t
    assert rel_fro(gram[list(g["gram_images"])], g["gram"]) <= TOL
    assert rel_fro(taps["tokens_in"].cpu().numpy(), g["tokens_in"]) <= TOL
    if "tokens_out" in g:
        assert rel_fro(taps["tokens_out"].cpu().numpy(), g["tokens_out"]) <= TOL
    assert rel_fro(taps["chan"][0].permute(2, 0, 1).cpu().numpy(), g["chan0"]) <= TOL
    xb = _photo_like(150, 77).cuda().to(torch.bfloat16)
    with torch.no_grad():
        alt = m.forward_packed(xb).cpu().numpy()
        for k in flags:
            setattr(m, k, True)
        dflt = m.forward_packed(xb).cpu().numpy()
    assert np.array_equal(argmax_table(alt, tasks), argmax_table(dflt, tasks))
    assert rel_fro(alt, dflt) <= 1e-4
    row = np.linalg.norm(alt.astype(np.float64) - dflt, axis=1) / np.linalg.norm(dflt.astype(np.float64), axis=1)
    assert row.max() <= 1e-3


def test_fp32_input_matches_bf16_input_on_bf16_exact_images():
    """fp32 NCHW is what the reference's DataLoader delivers (Functions/PMG_evaluation.py:126); on bf16-representable
    pixels the fp32 and bf16 entry formats must give the same logits up to summation order."""
    m, hp, tasks = _model(3)
    x = _photo_like(257, 5).cuda()
    with torch.no_grad():
        a = m.forward_packed(x).cpu().numpy()
        b = m.forward_packed(x.to(torch.bfloat16)).cpu().numpy()
    assert np.array_equal(argmax_table(a, tasks), argmax_table(b, tasks))
    assert rel_fro(a, b) <= 1e-4


def test_fp32_fused_front_matches_fp32_im2col_path_on_general_fp32_images():
    """Pixels that are NOT bf16-representable (what ToTensor + Normalize really produce): the fused front kernel reads the
    fp32 image as bf16 hi + lo planes (6 "channels" against the same conv0 weights), the unfused path as im2col hi/lo
    planes — same 16-bit operand precision, so they agree to fp32 summation order, and both agree with the fp32 port."""
    m, hp, tasks = _model(0)
    g = torch.Generator().manual_seed(99)
    x = (torch.rand(150, 3, 224, 224, generator=g) - 0.45) / 0.226           # fp32 with full mantissas
    port = PMGTorchPort(synth.pmg_state_dict(0, hp, tasks), tasks, hp["patch_div"], hp["transformer_heads"])
    ref = torch.cat([port(x)[t] for t in tasks], 1).numpy()
    with torch.no_grad():
        a = m.forward_packed(x.cuda()).cpu().numpy()
        m.fused_front = False
        b = m.forward_packed(x.cuda()).cpu().numpy()
    assert rel_fro(a, b) <= 1e-4
    for got in (a, b):
        assert np.array_equal(argmax_table(got, tasks), argmax_table(ref, tasks))
        assert max(rel_fro(v, split_packed(ref, tasks)[t]) for t, v in split_packed(got, tasks).items()) <= 2e-4


def test_unaligned_input_view_is_accepted():
    """A contiguous view at an odd storage offset: the boundary clones it instead of faulting in a 16-byte load."""
    m, hp, tasks = _model(0)
    base = torch.zeros(4 * 3 * 224 * 224 + 1, device="cuda")
    x = _photo_like(4, 9).cuda()
    view = base[1:].view(4, 3, 224, 224)
    view.copy_(x)
    assert view.data_ptr() % 16 != 0 and view.is_contiguous()
    with torch.no_grad():
        assert torch.equal(m.forward_packed(view), m.forward_packed(x))
