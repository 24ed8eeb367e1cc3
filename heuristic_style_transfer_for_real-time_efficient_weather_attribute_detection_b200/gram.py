"""Batched Gram matrices on the B200 library (BASELINE.json configs[1]).

``gram_batched(F, scale)`` = ``scale * F @ F.mT`` for F (B, C, HW) — the operation of
Models/models_PMG.py:188 and utils/PMG/PMG_style_transfer.py:44-45, as one kernel."""
from __future__ import annotations

import torch

from . import _lib


def gram_batched(F: torch.Tensor, scale: float | None = None) -> torch.Tensor:
    if not F.is_cuda:
        raise RuntimeError("gram_batched needs a CUDA tensor: there is no CPU fallback")
    if F.dim() != 3:
        raise ValueError("expected (B, C, HW)")
    if F.dtype == torch.bfloat16:
        dt = _lib.PMG_BF16
    elif F.dtype == torch.float32:
        dt = _lib.PMG_F32
    else:
        raise TypeError("float32 or bfloat16 only")
    F = F.contiguous()
    B, C, HW = F.shape
    if scale is None:
        scale = 1.0 / HW
    G = torch.empty((B, C, C), dtype=torch.float32, device=F.device)
    with torch.cuda.device(F.device):
        st = torch.cuda.current_stream(F.device).cuda_stream
        _lib.check(_lib.lib.pmg_gram_batched(F.data_ptr(), dt, B, C, HW, float(scale), G.data_ptr(), st),
                   "pmg_gram_batched")
    return G
