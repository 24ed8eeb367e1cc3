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
