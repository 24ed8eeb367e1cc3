"""Pre/post-processing either side of the forward.

CPU: the numpy oracle is pinned against the libraries the reference calls (torchvision ToTensor + Normalize,
torch.softmax / torch.max — utils/datasets_utils.py:212-214, Functions/PMG_evaluation.py:144-146).
GPU: the CUDA kernels (through the C ABI) against the oracle: bit-exact fp32 / nearest-even bf16 normalisation for both
layouts and ragged widths; softmax within 1e-6, arg-max / threshold rule identical; uint8 end-to-end pipeline."""
import numpy as np
import pytest
import torch

from helpers import load_contract, rel_fro
from oracle import prepost_oracle as po


def _images(seed, b, h, w):
    return np.random.default_rng(seed).integers(0, 256, size=(b, h, w, 3), dtype=np.uint8)


def test_oracle_matches_torchvision_transforms():
    tv = pytest.importorskip("torchvision.transforms")
    from PIL import Image
    x = _images(0, 3, 37, 53)
    tf = tv.Compose([tv.ToTensor(), tv.Normalize(list(po.IMAGENET_MEAN), list(po.IMAGENET_STD))])
    ref = torch.stack([tf(Image.fromarray(im)) for im in x]).numpy()
    assert np.array_equal(po.preprocess_u8(x), ref)          # same fp32 operation order → bit identical


def test_oracle_matches_torch_softmax_rule():
    _, _, tasks = load_contract()
    ncls = list(tasks.values())
    lg = (np.random.default_rng(1).standard_normal((64, sum(ncls))) * 3).astype(np.float32)
    probs, preds, maxp = po.postprocess(lg, ncls, 0.5)
    o = 0
    for t, n in enumerate(ncls):
        p = torch.softmax(torch.from_numpy(lg[:, o:o + n]), dim=1)
        mp, pr = torch.max(p, dim=1)
        pr[mp < 0.5] = -1
        assert np.allclose(probs[:, o:o + n], p.numpy(), atol=1e-6)
        assert np.array_equal(preds[:, t], pr.numpy())
        o += n


def test_bf16_rounding_helper():
    x = np.random.default_rng(2).standard_normal(4096).astype(np.float32)
    assert np.array_equal(po.to_bf16_rne(x), torch.from_numpy(x).to(torch.bfloat16).float().numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 224, 224), (3, 37, 53), (1, 16, 18)])
def test_preprocess_kernel_bit_exact(shape):
    from pmg_b200 import prepost
    b, h, w = shape
    x = _images(3, b, h, w)
    ref = po.preprocess_u8(x)
    xh = torch.from_numpy(x).cuda()
    xc = torch.from_numpy(np.ascontiguousarray(np.transpose(x, (0, 3, 1, 2)))).cuda()
    for layout, src in (("NHWC", xh), ("NCHW", xc)):
        got = prepost.preprocess_u8(src, layout, torch.float32).cpu().numpy()
        assert np.array_equal(got, ref), layout
        got16 = prepost.preprocess_u8(src, layout, torch.bfloat16).float().cpu().numpy()
        assert np.array_equal(got16, po.to_bf16_rne(ref)), layout


RESIZE_SIZES = [(300, 340), (256, 256), (480, 640), (1080, 1920), (257, 511), (1000, 300), (224, 224), (260, 900), (333, 257)]


@pytest.mark.parametrize("h,w", RESIZE_SIZES)
def test_resize_crop_oracle_is_bit_exact_vs_pillow_and_torchvision(h, w):
    """The oracle restates Pillow's Resample.c + torchvision's size / crop rules: pin it to the libraries themselves."""
    from PIL import Image
    from torchvision import transforms
    x = np.random.default_rng(h * 7 + w).integers(0, 256, (1, h, w, 3), dtype=np.uint8)
    ref = np.asarray(transforms.Compose([transforms.Resize(256), transforms.CenterCrop(224)])(Image.fromarray(x[0])))
    assert np.array_equal(po.resize_center_crop_u8(x)[0], ref)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", RESIZE_SIZES)
def test_resize_crop_kernel_is_bit_exact(h, w):
    """CUDA Resize(256) + CenterCrop(224) == oracle (== Pillow) byte for byte; fused with ToTensor + Normalize it equals
    build_default_transform (utils/datasets_utils.py:205-215) bit for bit in fp32."""
    from pmg_b200 import prepost
    x = np.random.default_rng(h * 7 + w).integers(0, 256, (3, h, w, 3), dtype=np.uint8)
    ref = po.resize_center_crop_u8(x)
    xd = torch.from_numpy(x).cuda()
    assert np.array_equal(prepost.resize_center_crop_u8(xd).cpu().numpy(), ref)
    f32 = prepost.resize_center_crop_u8(xd, out_dtype=torch.float32).cpu().numpy()
    assert np.array_equal(f32, po.preprocess_u8(ref))
    b16 = prepost.resize_center_crop_u8(xd, out_dtype=torch.bfloat16).float().cpu().numpy()
    assert np.array_equal(b16, po.to_bf16_rne(po.preprocess_u8(ref)))


@pytest.mark.gpu
def test_postprocess_kernel():
    from pmg_b200 import prepost
    _, _, tasks = load_contract()
    ncls = list(tasks.values())
    lg = (np.random.default_rng(4).standard_normal((257, sum(ncls))) * 3).astype(np.float32)
    for thr in (0.0, 0.6):
        probs, preds, maxp = prepost.postprocess(torch.from_numpy(lg).cuda(), tasks, thr)
        rp, rd, rm = po.postprocess(lg, ncls, thr)
        assert np.allclose(probs.cpu().numpy(), rp, atol=1e-6)
        assert np.allclose(maxp.cpu().numpy(), rm, atol=1e-6)
        near = np.abs(rm - thr) < 1e-5                      # ties with the threshold may round either way
        assert np.array_equal(preds.cpu().numpy()[~near], rd[~near])


@pytest.mark.gpu
def test_uint8_host_pipeline_matches_float_path():
    """HostPipeline on decoded uint8 images == forward on the oracle-normalised bf16 batch (same kernels after the
    pre-processing kernel), ragged last chunk included."""
    from helpers import build_model, torch_state_dict
    from oracle import synth
    from pmg_b200.runner import HostPipeline
    _, hp, tasks = load_contract()
    m = build_model(hp, tasks)
    m.load_state_dict(torch_state_dict(synth.pmg_state_dict(0, hp, tasks)), strict=True)
    m.cuda().eval()
    x = _images(5, 21, 224, 224)
    xf = torch.from_numpy(po.to_bf16_rne(po.preprocess_u8(x))).cuda().to(torch.bfloat16)
    with torch.no_grad():
        # same chunking as the pipeline: the batch size picks token_proj's split-K factor (summation order)
        ref = torch.cat([m.forward_packed(xf[i:i + 8]) for i in range(0, 21, 8)]).cpu().numpy()
        pipe = HostPipeline(m, chunk=8)
        got = pipe.infer(torch.from_numpy(x).pin_memory()).clone().numpy()
        pipe_c = HostPipeline(m, chunk=8, u8_layout="NCHW")
        got_c = pipe_c.infer(torch.from_numpy(np.ascontiguousarray(np.transpose(x, (0, 3, 1, 2)))).pin_memory()).clone().numpy()
    torch.cuda.synchronize()
    assert rel_fro(got, ref) <= 1e-6 and rel_fro(got_c, ref) <= 1e-6


@pytest.mark.gpu
def test_host_pipeline_streaming_matches_blocking():
    """submit()/collect() with two batches outstanding delivers, per ticket, exactly what a blocking infer() of the same
    batch returns (tail-chunk schedule, alternating device / pinned buffers across calls, ragged batch sizes)."""
    from helpers import build_model, torch_state_dict
    from oracle import synth
    from pmg_b200.runner import HostPipeline
    _, hp, tasks = load_contract()
    m = build_model(hp, tasks)
    m.load_state_dict(torch_state_dict(synth.pmg_state_dict(0, hp, tasks)), strict=True)
    m.cuda().eval()
    batches = [torch.from_numpy(_images(10 + i, n, 224, 224)).pin_memory() for i, n in enumerate((37, 64, 5, 70))]
    pipe = HostPipeline(m, chunk=32, tail_chunk=8)
    assert sum(b - a for a, b in pipe._schedule(70)) == 70 and pipe._schedule(70)[-1][1] - pipe._schedule(70)[-1][0] <= 8
    want = [pipe.infer(b).clone() for b in batches]
    got = []
    prev = pipe.submit(batches[0])
    for b in batches[1:]:
        nxt = pipe.submit(b)
        got.append(pipe.collect(prev).clone())
        prev = nxt
    got.append(pipe.collect(prev).clone())
    for g, w in zip(got, want):
        assert g.shape == w.shape and torch.equal(g, w)
    with pytest.raises(AssertionError):
        pipe.submit(batches[0]); pipe.submit(batches[1]); pipe.submit(batches[2])
