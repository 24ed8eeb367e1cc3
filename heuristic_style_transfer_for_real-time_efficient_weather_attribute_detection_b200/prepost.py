"""GPU pre/post-processing either side of ``model(x)`` in the reference's evaluation loops.

``preprocess_u8``  = ``transforms.ToTensor()`` + ``transforms.Normalize(mean, std)`` of
utils/datasets_utils.py:212-214 on decoded uint8 images; ``resize_center_crop_u8`` = the ``Resize(256)`` + ``CenterCrop(224)``
in front of them (:209-210), bit-identical to Pillow's 8-bit bilinear resample, optionally fused with the former;
``postprocess``    = the per-task softmax / max / arg-max / ``prob_threshold`` → -1 block of
Functions/PMG_evaluation.py:144-146 on the packed logits.  Both call the CUDA library; no CPU fallback."""
from __future__ import annotations

import ctypes
from typing import Dict, Sequence, Tuple

import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)     # utils/datasets_utils.py:213
IMAGENET_STD = (0.229, 0.224, 0.225)      # utils/datasets_utils.py:214


def preprocess_u8(x: torch.Tensor, layout: str = "NHWC", out_dtype: torch.dtype = torch.bfloat16,
                  mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD,
                  out: torch.Tensor | None = None) -> torch.Tensor:
    """uint8 (B,H,W,3) [layout="NHWC"] or (B,3,H,W) [layout="NCHW"] on CUDA → normalised (B,3,H,W)."""
    if not x.is_cuda:
        raise RuntimeError("preprocess_u8 needs a CUDA tensor: there is no CPU fallback")
    if x.dtype != torch.uint8 or x.dim() != 4:
        raise TypeError("expected a 4-D uint8 tensor")
    if layout == "NHWC":
        B, H, W, C = x.shape
        lay = 1
    elif layout == "NCHW":
        B, C, H, W = x.shape
        lay = 0
    else:
        raise ValueError("layout must be 'NHWC' or 'NCHW'")
    if C != 3:
        raise ValueError("3-channel images only")
    dt = {torch.bfloat16: _lib.PMG_BF16, torch.float32: _lib.PMG_F32}.get(out_dtype)
    if dt is None:
        raise TypeError("out_dtype must be bfloat16 or float32")
    x = x.contiguous()
    if out is None:
        out = torch.empty((B, 3, H, W), dtype=out_dtype, device=x.device)
    elif out.shape != (B, 3, H, W) or out.dtype != out_dtype or not out.is_contiguous():
        raise ValueError("out has the wrong shape / dtype / layout")
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.lib.pmg_preprocess_u8(x.data_ptr(), lay, B, H, W, m, s, dt, out.data_ptr(), st), "pmg_preprocess_u8")
    return out


def resize_center_crop_u8(x: torch.Tensor, size: int = 256, crop: int = 224, out_dtype: torch.dtype = torch.uint8,
                          mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD) -> torch.Tensor:
    """``transforms.Resize(size)`` + ``transforms.CenterCrop(crop)`` (utils/datasets_utils.py:209-210) on decoded uint8 HWC
    images (B,H,W,3) on CUDA, bit-identical to the Pillow / torchvision CPU transform.  ``out_dtype=torch.uint8`` →
    (B,crop,crop,3) uint8; ``torch.bfloat16`` / ``torch.float32`` → ToTensor + Normalize fused, (B,3,crop,crop) — the whole
    ``build_default_transform`` in one kernel."""
    if not x.is_cuda:
        raise RuntimeError("resize_center_crop_u8 needs a CUDA tensor: there is no CPU fallback")
    if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[3] != 3:
        raise TypeError("expected a (B, H, W, 3) uint8 tensor")
    mode = {torch.uint8: 0, torch.bfloat16: 1, torch.float32: 2}.get(out_dtype)
    if mode is None:
        raise TypeError("out_dtype must be uint8, bfloat16 or float32")
    x = x.contiguous()
    B, H, W, _ = x.shape
    out = torch.empty((B, crop, crop, 3) if mode == 0 else (B, 3, crop, crop), dtype=out_dtype, device=x.device)
    if B == 0:
        return out
    need = _lib.lib.pmg_resize_scratch_bytes(H, W, size, crop)
    if need == 0:
        raise ValueError(f"a {H}x{W} image resized to shorter edge {size} is smaller than the {crop} crop")
    scratch = torch.empty(need, dtype=torch.uint8, device=x.device)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.lib.pmg_resize_crop_u8(x.data_ptr(), B, H, W, size, crop, m, s, mode, out.data_ptr(), scratch.data_ptr(),
                                               need, st), "pmg_resize_crop_u8")
    scratch.record_stream(torch.cuda.current_stream(x.device))
    return out


def postprocess(logits: torch.Tensor, num_classes_per_task: Dict[str, int], prob_threshold: float = 0.0
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Packed logits (B, Σncls) → (probs (B, Σncls), preds (B, n_tasks) int32 with -1 below the threshold,
    max_probs (B, n_tasks)); tasks in ``num_classes_per_task`` (Tasks.json) order."""
    if not logits.is_cuda:
        raise RuntimeError("postprocess needs a CUDA tensor: there is no CPU fallback")
    ncls = [int(v) for v in num_classes_per_task.values()]
    B, L = logits.shape
    if L != sum(ncls) or logits.dtype != torch.float32:
        raise ValueError("logits must be fp32 (B, sum(ncls))")
    logits = logits.contiguous()
    probs = torch.empty_like(logits)
    preds = torch.empty((B, len(ncls)), dtype=torch.int32, device=logits.device)
    maxp = torch.empty((B, len(ncls)), dtype=torch.float32, device=logits.device)
    arr = (ctypes.c_int32 * len(ncls))(*ncls)
    with torch.cuda.device(logits.device):
        st = torch.cuda.current_stream(logits.device).cuda_stream
        _lib.check(_lib.lib.pmg_postprocess(logits.data_ptr(), B, len(ncls), arr, float(prob_threshold), probs.data_ptr(),
                                            preds.data_ptr(), maxp.data_ptr(), st), "pmg_postprocess")
    return probs, preds, maxp
