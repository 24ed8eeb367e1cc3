"""TEST INFRASTRUCTURE — CPU restatement of the pre/post-processing around the reference forward.

preprocess_u8: transforms.ToTensor() (uint8 HWC → float32 CHW, value / 255) then transforms.Normalize(mean, std)
               ((x - mean) / std), /root/reference/utils/datasets_utils.py:212-214.  The arithmetic lives in
               torchvision (reference pins 0.23.0, requirements.txt:56): ToTensor = ``img.permute(2,0,1).float().div(255)``,
               Normalize = ``tensor.sub_(mean).div_(std)`` in float32.  Pinned against torchvision itself in
               tests/test_prepost.py.
postprocess:   torch.softmax(logits, dim=1); torch.max(probs, dim=1); preds[max_probs < prob_threshold] = -1,
               /root/reference/Functions/PMG_evaluation.py:144-146.
Only tests/ may import this module."""
from __future__ import annotations

import numpy as np

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def preprocess_u8(x_hwc: np.ndarray, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> np.ndarray:
    """(B,H,W,3) uint8 → (B,3,H,W) float32."""
    assert x_hwc.dtype == np.uint8 and x_hwc.ndim == 4 and x_hwc.shape[3] == 3
    t = np.transpose(x_hwc, (0, 3, 1, 2)).astype(np.float32) / np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32).reshape(1, 3, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(1, 3, 1, 1)
    return ((t - m) / s).astype(np.float32)


PRECISION_BITS = 32 - 8 - 2       # Pillow src/libImaging/Resample.c


def resize_output_size(h: int, w: int, size: int = 256):
    """transforms.Resize(int): the SHORTER edge becomes ``size``, the longer ``int(size * long / short)`` — torchvision
    transforms/functional.py ``_compute_resized_output_size`` (reference: utils/datasets_utils.py:209).  Returns (new_h, new_w)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    return new_h, new_w


def pil_bilinear_coeffs(in_size: int, out_size: int):
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` (Resample.c) for the BILINEAR filter (support 1.0) over the
    whole axis: per output index the first input index, the tap count and the taps as 22-bit fixed point."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
            ww += w[x]
        if ww != 0.0:
            w[:xmax] /= ww
        for x in range(ksize):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(img: np.ndarray, bounds, kk, axis: int) -> np.ndarray:
    """One 8-bit pass of ImagingResample{Horizontal,Vertical}_8bpc: Σ pixel·tap in int32 from 1 << 21, >> 22, clamp."""
    img = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + img.shape[1:], dtype=np.uint8)
    for o in range(bounds.shape[0]):
        x0, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for j in range(n):
            acc += img[x0 + j] * int(kk[o, j])
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_center_crop_u8(x_hwc: np.ndarray, size: int = 256, crop: int = 224) -> np.ndarray:
    """transforms.Resize(size) + transforms.CenterCrop(crop) on decoded uint8 images (B,H,W,3) → (B,crop,crop,3)
    (utils/datasets_utils.py:209-210).  The arithmetic lives in Pillow (``Image.resize(..., BILINEAR)``: horizontal 8-bit pass,
    then vertical 8-bit pass, each rounded to uint8 — ImagingResample) and torchvision (output size, crop offsets
    ``int(round((H' - crop) / 2.0))``); pinned bit-exactly against both in tests/test_prepost.py."""
    B, H, W, _ = x_hwc.shape
    nh, nw = resize_output_size(H, W, size)
    if (nh, nw) != (H, W):
        bh, kh = pil_bilinear_coeffs(W, nw)
        bv, kv = pil_bilinear_coeffs(H, nh)
        if nw != W:
            x_hwc = _resample_axis(x_hwc, bh, kh, axis=2)
        if nh != H:
            x_hwc = _resample_axis(x_hwc, bv, kv, axis=1)
    top, left = int(round((nh - crop) / 2.0)), int(round((nw - crop) / 2.0))
    assert nh >= crop and nw >= crop, "CenterCrop padding of images smaller than the crop is not restated"
    return np.ascontiguousarray(x_hwc[:, top:top + crop, left:left + crop])


def to_bf16_rne(x: np.ndarray) -> np.ndarray:
    """float32 → nearest-even bfloat16, returned as float32."""
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def postprocess(logits: np.ndarray, ncls, prob_threshold: float = 0.0):
    """(B, Σncls) float32 → probs (B, Σncls), preds (B, ntask) int32, max_probs (B, ntask)."""
    B = logits.shape[0]
    probs = np.zeros_like(logits, dtype=np.float32)
    preds = np.zeros((B, len(ncls)), dtype=np.int32)
    maxp = np.zeros((B, len(ncls)), dtype=np.float32)
    o = 0
    for t, n in enumerate(ncls):
        z = logits[:, o:o + n].astype(np.float32)
        e = np.exp(z - z.max(axis=1, keepdims=True))
        p = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
        probs[:, o:o + n] = p
        maxp[:, t] = p.max(axis=1)
        a = p.argmax(axis=1).astype(np.int32)
        a[maxp[:, t] < np.float32(prob_threshold)] = -1
        preds[:, t] = a
        o += n
    return probs, preds, maxp
