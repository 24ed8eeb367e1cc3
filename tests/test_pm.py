"""PM model (BASELINE.json configs[3]): boundary on CPU, parity on GPU."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from helpers import GOLDEN, REFERENCE, rel_fro, torch_state_dict
from oracle import pm_oracle, synth
from test_pm_oracle_golden import CASES, load_pm_case

TOL = 1e-3


def build_pm(hp, tasks):
    import pmg_b200
    return pmg_b200.MultiTaskPatchGAN(tasks_dict=tasks, input_nc=3, ndf=hp["ndf"], norm=hp["norm"], patch_size=hp["patch_size"],
                                      device="cpu", attn_tau=hp["attn_tau"], attn_use_se=hp["attn_use_se"],
                                      attn_softmax_spatial=hp["attn_softmax_spatial"])


def test_pm_state_dict_layout_matches_reference_contract():
    with open(os.path.join(GOLDEN, "pm_contract.json")) as f:
        c = json.load(f)
    tasks = {t: len(v) for t, v in c["tasks"].items()}
    m = build_pm(c["hparams"], tasks)
    sd = m.state_dict()
    assert [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()] == c["state_dict"]
    assert sum(p.numel() for p in m.parameters()) == c["n_params"]
    missing, unexpected = m.load_state_dict(torch_state_dict(synth.pm_state_dict(0)), strict=True)
    assert not missing and not unexpected


def test_pm_errors_without_gpu():
    import pmg_b200
    hp, tasks = synth.PM_HPARAMS, synth.PMG_TASKS
    m = build_pm(hp, tasks)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 224, 224))
    with pytest.raises(NotImplementedError):
        pmg_b200.MultiTaskPatchGAN(tasks, norm="batch")
    with pytest.raises(NotImplementedError):
        pmg_b200.MultiTaskPatchGAN(tasks, attn_use_se=True)
    with pytest.raises(NotImplementedError):
        pmg_b200.MultiTaskPatchGAN(tasks, attn_use_se=False, attn_softmax_spatial=False)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout only exists in the build container")
def test_pm_reference_helpers_bind_to_b200_module(tmp_path, capsys):
    """`from Models.models_PM import MultiTaskPatchGAN, load_model_weights, …` (test_PM.py:16-21): the model class resolves to
    this package, the four host-side helpers to the reference's OWN definitions (bound by compat, not restated here)."""
    import pmg_b200
    saved = {k: sys.modules.get(k) for k in ("Models", "Models.models_PM", "Models.models_PMG")}
    sys.path.insert(0, REFERENCE)                      # what `python /root/reference/test_PM.py` does
    try:
        for k in saved:
            sys.modules.pop(k, None)
        pmg_b200.install_as_reference_models()
        from Models.models_PM import MultiTaskPatchGAN, checkpoint_has_se, load_model_weights, print_model_parameters
        m = MultiTaskPatchGAN(tasks_dict=synth.PMG_TASKS, input_nc=3, ndf=64, norm="instance", patch_size=73, device="cpu",
                              attn_tau=0.7, attn_use_se=False, attn_softmax_spatial=True)
        assert type(m) is pmg_b200.MultiTaskPatchGAN
        ck = tmp_path / "best_model_PM.pth"
        torch.save({"state_dict": {"module." + k: v for k, v in torch_state_dict(synth.pm_state_dict(0)).items()}}, ck)
        assert not checkpoint_has_se(str(ck))
        load_model_weights(m, str(ck), torch.device("cpu"), strict=True)
        print_model_parameters(m)
        assert "11215692" in capsys.readouterr().out
        assert load_model_weights.__module__ == "_pmg_b200_reference_models_PM"
    finally:
        sys.path.remove(REFERENCE)
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


@pytest.mark.gpu
@pytest.mark.parametrize("in_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name,contract", CASES)
def test_pm_golden_gpu(name, contract, in_dtype):
    """CUDA PM forward vs outputs of the unmodified reference: argmax identical, logits / attention within 1e-3."""
    g, hp, tasks, seed, batch, h, w = load_pm_case(name, contract)
    m = build_pm(hp, tasks)
    m.load_state_dict(torch_state_dict(synth.pm_state_dict(seed, hp, tasks)))
    m.to("cuda").eval()
    x = torch.from_numpy(synth.images(seed, batch, h, w)).to(in_dtype)
    with torch.no_grad():
        full = m(x.cuda(), return_full=True)
        plain = m(x)                                   # CPU tensor is moved to the model's device like the reference does
    assert list(full.keys()) == list(tasks)
    logits = torch.cat([full[t]["logits"] for t in tasks], 1).cpu().numpy()
    attn = torch.stack([full[t]["attn"][:, 0] for t in tasks], 1).cpu().numpy()
    o = 0
    for t, n in tasks.items():
        assert np.array_equal(logits[:, o:o + n].argmax(1), g["logits"][:, o:o + n].argmax(1)), t
        assert rel_fro(logits[:, o:o + n], g["logits"][:, o:o + n]) <= TOL, t
        assert torch.equal(plain[t], full[t]["logits"])
        o += n
    assert rel_fro(attn, g["attn"]) <= TOL
    with torch.no_grad():
        emb = m(x.cuda(), return_embeddings=True)
    assert emb.shape == (batch, int(np.prod(attn.shape[2:])) * m.trunk[-3].out_channels) and emb.device.type == "cpu"
    assert np.allclose(np.sqrt((emb.double().numpy() ** 2).sum(1)), g["feats_fro"], rtol=TOL)


@pytest.mark.gpu
def test_pm_against_oracle_ragged_batch():
    hp, tasks = synth.PM_HPARAMS, synth.PMG_TASKS
    sd = synth.pm_state_dict(9, hp, tasks)
    x = synth.images(9, 5)
    ref = pm_oracle.pm_forward(x, sd, tasks, tau=hp["attn_tau"])
    m = build_pm(hp, tasks)
    m.load_state_dict(torch_state_dict(sd))
    m.to("cuda").eval()
    with torch.no_grad():
        out = m(torch.from_numpy(x).cuda().to(torch.bfloat16))
    for t in tasks:
        got = out[t].cpu().numpy()
        assert np.array_equal(got.argmax(1), ref[t].argmax(1)), t
        assert rel_fro(got, ref[t]) <= TOL, t


@pytest.mark.gpu
@pytest.mark.parametrize("batch,h,w", [(3, 224, 224), (2, 160, 208), (258, 224, 224)])
def test_pm_implicit_gemm_matches_im2col_path(batch, h, w):
    """Stride-2 trunk layers as implicit GEMM (TMA boxes with element strides straight from the NHWC planes, tc_igemm.cuh)
    vs the im2col + GEMM path: the same products in the same K order.  160×208 gives tiles whose last M tile of an image
    is short; 258 images cross the internal 128-image passes (two in flight on two streams)."""
    hp, tasks = synth.PM_HPARAMS, synth.PMG_TASKS
    m = build_pm(hp, tasks)
    m.load_state_dict(torch_state_dict(synth.pm_state_dict(4, hp, tasks)))
    m.to("cuda").eval()
    x = torch.from_numpy(synth.images(4, batch, h, w)).cuda().to(torch.bfloat16)
    with torch.no_grad():
        a = m(x, return_full=True)                     # 256-wide N tiles where 256 divides the output channels
        m.igemm_n128 = True
        c = m(x, return_full=True)                     # 128-wide N tiles everywhere
        m.igemm_n128 = False
        m.fused_stats = False
        d = m(x, return_full=True)                     # InstanceNorm statistics from a separate pass over the conv output
        m.fused_stats = True
        m.fused_conv0 = False
        e = m(x, return_full=True)                     # layer 0 as im2col + GEMM + InstanceNorm kernels
        ef = m(x.float(), return_full=True)
        m.fused_conv0 = True
        af = m(x.float(), return_full=True)            # fp32 image: hi/lo operand planes built in the layer-0 loader
        m.implicit_gemm = False
        b = m(x, return_full=True)
    for t in tasks:
        # same products in the same K order; the InstanceNorm statistics come from the conv epilogue (Chan-merged two-pass
        # partials) on one side and from the single-pass shifted kernel on the other: rounding-level differences
        assert rel_fro(a[t]["logits"].cpu().numpy(), b[t]["logits"].cpu().numpy()) <= 1e-4, t
        assert rel_fro(a[t]["attn"].cpu().numpy(), b[t]["attn"].cpu().numpy()) <= 1e-4, t
        assert torch.equal(a[t]["logits"], c[t]["logits"]) and torch.equal(a[t]["attn"], c[t]["attn"]), t   # same products, same K order
        assert rel_fro(a[t]["logits"].cpu().numpy(), d[t]["logits"].cpu().numpy()) <= 1e-4, t   # two statistics algorithms: rounding-level differences
        assert rel_fro(a[t]["attn"].cpu().numpy(), d[t]["attn"].cpu().numpy()) <= 1e-4, t
        for u, v in ((a, e), (af, ef), (a, af)):       # the bf16 image is exact in fp32: all four must agree
            assert rel_fro(u[t]["logits"].cpu().numpy(), v[t]["logits"].cpu().numpy()) <= 1e-4, t
            assert rel_fro(u[t]["attn"].cpu().numpy(), v[t]["attn"].cpu().numpy()) <= 1e-4, t


@pytest.mark.gpu
def test_pm_forward_is_graph_capturable_with_two_passes_in_flight():
    """pm_forward runs its 128-image passes two at a time, the second on a library-owned side stream forked from / joined to
    the caller's stream with events: the whole call must still be ordered on the caller's stream — in particular it must be
    capturable into a CUDA graph and replay to the same bits."""
    hp, tasks = synth.PM_HPARAMS, synth.PMG_TASKS
    m = build_pm(hp, tasks)
    m.load_state_dict(torch_state_dict(synth.pm_state_dict(4, hp, tasks)))
    m.to("cuda").eval()
    x = torch.from_numpy(synth.images(6, 130, 224, 224)).cuda().to(torch.bfloat16)       # 128 + 2 images: both lanes
    with torch.no_grad():
        want = m(x)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            m(x)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            got = m(x)
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
    for t in tasks:
        assert torch.equal(got[t], want[t]), t
