"""runner.GraphedPMG — the CUDA-graph latency path that stands in for the per-image loops of the reference's callers
(utils/benchmark_utils.py:67-75: B = 1, one ``.cpu()`` per task; utils/camera_utils.py:119-123; the ragged last
DataLoader batch of Functions/PMG_evaluation.py:124-129).  Replay must be bit-identical to the eager module call."""
import numpy as np
import pytest
import torch

from helpers import build_model, load_contract, torch_state_dict
from oracle import prepost_oracle, synth

pytestmark = pytest.mark.gpu


def _model(seed=0):
    _, hp, tasks = load_contract()
    m = build_model(hp, tasks)
    m.load_state_dict(torch_state_dict(synth.pmg_state_dict(seed, hp, tasks)), strict=True)
    m.to("cuda").eval()
    return m, hp, tasks


def _x(seed, n, dtype=torch.float32):
    return torch.from_numpy(synth.images(seed, n)).cuda().to(dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_replay_equals_eager_b1_and_ragged(dtype):
    from pmg_b200.runner import GraphedPMG
    m, hp, tasks = _model(0)
    g = GraphedPMG(m)
    xs = [_x(40 + i, n, dtype) for i, n in enumerate([1, 1, 5, 1, 37, 5])]       # shapes alternate: graphs are cached per shape
    with torch.no_grad():
        want = [m.forward_packed(x).clone() for x in xs]
    for x, w in zip(xs, want):
        out = g(x)
        assert list(out.keys()) == list(tasks)
        assert torch.equal(torch.cat([out[t] for t in tasks], 1), w)
    assert len(g._graphs) == 3
    e = g(xs[0][:0])                                                             # empty last batch: no launch, empty dict
    assert [tuple(v.shape) for v in e.values()] == [(0, n) for n in tasks.values()]


def test_small_graph_survives_larger_capture_and_eager_calls():
    """a graph captured for B = 1 must stay valid after a larger shape is captured and after eager calls that
    re-allocate the module's shared workspace — every graph owns its workspace."""
    from pmg_b200.runner import GraphedPMG
    m, hp, tasks = _model(1)
    g = GraphedPMG(m)
    x1, x8 = _x(50, 1), _x(51, 8)
    with torch.no_grad():
        w1 = m.forward_packed(x1).clone()
    a = g.forward_packed(x1).clone()
    ws_small = g._graphs[(tuple(x1.shape), x1.dtype, x1.device)]["ws"]
    g.forward_packed(x8)
    with torch.no_grad():
        m.forward_packed(_x(52, 64))                 # grows model._ws (frees the old buffer)
        junk = [torch.full((1 << 20,), float("nan"), device="cuda") for _ in range(8)]   # recycle freed blocks
    b = g.forward_packed(x1).clone()
    del junk
    assert torch.equal(a, w1) and torch.equal(b, w1)
    assert g._graphs[(tuple(x1.shape), x1.dtype, x1.device)]["ws"] is ws_small
    assert ws_small.data_ptr() != m._ws.data_ptr()


def test_recapture_after_load_state_dict():
    from pmg_b200.runner import GraphedPMG
    m, hp, tasks = _model(0)
    g = GraphedPMG(m)
    x = _x(60, 2)
    a = g.forward_packed(x).clone()
    m.load_state_dict(torch_state_dict(synth.pmg_state_dict(1, hp, tasks)))
    b = g.forward_packed(x).clone()
    with torch.no_grad():
        want = m.forward_packed(x)
    assert not torch.equal(a, b) and torch.equal(b, want)


@pytest.mark.parametrize("host_outputs", [False, True])
def test_postprocess_and_host_outputs(host_outputs):
    """postprocess=True adds the per-task softmax / arg-max / threshold (PMG_evaluation.py:144-146) to the graph;
    host_outputs=True lands the packed logits in pinned host memory with ONE sync (benchmark_utils.py:73 does 12)."""
    from pmg_b200.runner import GraphedPMG
    m, hp, tasks = _model(2)
    g = GraphedPMG(m, postprocess=True, prob_threshold=0.4, host_outputs=host_outputs)
    for seed, n in [(70, 1), (71, 3), (72, 1)]:
        x = _x(seed, n)
        with torch.no_grad():
            want = m.forward_packed(x).cpu()
        lg = g.forward_packed(x)
        assert (not lg.is_cuda and lg.is_pinned()) if host_outputs else lg.is_cuda
        assert torch.equal(lg.cpu(), want)
        probs, preds, maxp = g.predictions(x)
        rp, rpred, rmax = prepost_oracle.postprocess(want.numpy(), list(tasks.values()), 0.4)
        near = np.abs(rmax - 0.4) < 1e-5                    # ties with the threshold may round either way
        assert np.array_equal(preds.cpu().numpy()[~near], rpred[~near])
        assert np.allclose(probs.cpu().numpy(), rp, atol=1e-6)
        assert np.allclose(maxp.cpu().numpy(), rmax, atol=1e-6)
