"""Drop-in replacement for the reference module ``Models/models_PM.py`` (PM = PatchGAN multi-task model).

``MultiTaskPatchGAN`` keeps the reference constructor (models_PM.py:91-103), the same state_dict keys
(``trunk.{0,1,3,4,…}``, ``task_heads.<task>.attn_conv / cls_conv``) and the same call forms
(``model(x)``, ``return_full``, ``return_embeddings``, ``return_task_embeddings`` — :159-194).  The arithmetic runs
in the CUDA library behind ``pm_forward`` of ``include/pmg_b200.h`` (tcgen05 GEMMs over im2col planes,
InstanceNorm with per-sample statistics, task heads as one GEMM + a warp-shuffle softmax); there is no CPU path.
One module instance holds ONE workspace: concurrent calls from several Python threads (PM_evaluation.py:845-865) are
serialised by a lock around the enqueue, so all of them must target the same CUDA stream.

Supported configuration = what test_PM.py builds (:153-164) with hyperparameters_PM.json: ``norm="instance"``,
``attn_use_se=False``, ``attn_softmax_spatial=True``; other values raise at construction.
"""
from __future__ import annotations

import ctypes
import threading
from typing import Any, Dict

import torch
import torch.nn as nn

from . import _lib

__all__ = ["MultiTaskPatchGAN", "TaskHeadImproved", "SE"]


class SE(nn.Module):
    def __init__(self, c: int, r: int = 16):
        super().__init__()
        raise NotImplementedError("attn_use_se=True (SE block, models_PM.py:10-23) is not implemented by the B200 kernels; "
                                  "hyperparameters_PM.json sets attn_use_se=false")


class TaskHeadImproved(nn.Module):
    """Parameter holder of one task head (reference models_PM.py:27-76); the math is fused into pm_forward."""

    def __init__(self, in_channels: int, out_channels: int, use_se: bool = True, tau: float = 0.7,
                 use_softmax_spatial: bool = True, ablate_attention: bool = False):
        super().__init__()
        if use_se:
            SE(in_channels)
        if not use_softmax_spatial:
            raise NotImplementedError("attn_softmax_spatial=False (sigmoid attention) is not implemented by the B200 kernels")
        self.use_se, self.tau = use_se, tau
        self.use_softmax_spatial, self.ablate_attention = use_softmax_spatial, ablate_attention
        self.se = nn.Identity()
        self.attn_conv = nn.Conv2d(in_channels, 1, kernel_size=1, bias=True)
        self.cls_conv = nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=False)

    def forward(self, feat):  # pragma: no cover - parameter holder
        raise RuntimeError("TaskHeadImproved is fused into the PM CUDA forward; call the model instead")


def _trunk_plan(input_nc: int, ndf: int, patch_size: float):
    plan, cin, width, rf = [], input_nc, ndf, float(patch_size)
    while rf > 4 and width <= 512:
        plan.append((cin, width, 2))
        cin, width, rf = width, width * 2, rf / 2
    plan.append((cin, width, 1))
    return plan


class MultiTaskPatchGAN(nn.Module):
    def __init__(self, tasks_dict: Dict[str, int], input_nc: int = 3, ndf: int = 64, norm: str = "instance",
                 patch_size: int = 70, device: torch.device | str = "cpu", attn_tau: float = 0.7, attn_use_se: bool = True,
                 attn_softmax_spatial: bool = True, ablate_attention: bool = False):
        super().__init__()
        if norm != "instance":
            raise NotImplementedError(f"norm={norm!r}: the PM B200 kernels implement InstanceNorm2d only (test_PM.py:157)")
        self.device = torch.device(device)
        self.tasks_dict = tasks_dict
        self.ablate_attention = ablate_attention
        layers = []
        for cin, cout, stride in _trunk_plan(input_nc, ndf, patch_size):
            layers += [nn.Conv2d(cin, cout, 4, stride, 1), nn.InstanceNorm2d(cout, affine=True), nn.LeakyReLU(0.2, True)]
        self.trunk = nn.Sequential(*layers).to(self.device)
        final_c = layers[-3].out_channels
        self.task_heads = nn.ModuleDict({
            t: TaskHeadImproved(final_c, n, use_se=attn_use_se, tau=attn_tau, use_softmax_spatial=attn_softmax_spatial,
                                ablate_attention=ablate_attention).to(self.device)
            for t, n in tasks_dict.items()})
        self._cfg = _lib.make_pm_config(input_nc, ndf, patch_size, list(tasks_dict.values()), attn_tau, ablate_attention)
        if _lib.lib.pm_param_count(ctypes.byref(self._cfg)) < 0:
            raise ValueError("configuration rejected by the PM CUDA library (ndf must be a multiple of 32, ≤ 8 classes "
                             "per task, n_tasks + total classes ≤ 64)")
        self._n_logits = sum(tasks_dict.values())
        self._packed = self._packed_key = self._ws = None
        self._lock = threading.Lock()

    def __getstate__(self):
        st = dict(self.__dict__)
        for k in ("_packed", "_packed_key", "_ws", "_lock", "_staged"):
            st.pop(k, None)
        c = st.pop("_cfg")
        st["_cfg_tuple"] = (c.input_nc, c.ndf, c.patch_size, [c.ncls[i] for i in range(c.n_tasks)], c.attn_tau, c.ablate_attention)
        return st

    def __setstate__(self, st):
        st = dict(st)
        cfg = st.pop("_cfg_tuple")
        super().__setstate__(st)
        self._cfg = _lib.make_pm_config(*cfg)
        self._packed = self._packed_key = self._ws = None
        self._lock = threading.Lock()

    def _abi_params(self):
        out = []
        mods = list(self.trunk)
        for i in range(0, len(mods), 3):
            out += [mods[i].weight, mods[i].bias, mods[i + 1].weight, mods[i + 1].bias]
        for t in self.tasks_dict:
            h = self.task_heads[t]
            out += [h.attn_conv.weight, h.attn_conv.bias, h.cls_conv.weight]
        return out

    def _ensure_packed(self, device, stream_ptr):
        params = self._abi_params()
        key = (device, tuple((p.data_ptr(), p._version) for p in params))
        if self._packed is not None and key == self._packed_key:
            return self._packed
        staged = []
        for p in params:
            if p.device != device:
                raise RuntimeError(f"model parameters live on {p.device} but the input is on {device}; call model.to(device)")
            t = p.detach()
            staged.append(t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous())
        n = len(staged)
        ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in staged])
        nbytes = _lib.lib.pm_packed_bytes(ctypes.byref(self._cfg))
        packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _lib.check(_lib.lib.pm_pack_weights(ctypes.byref(self._cfg), ptrs, n, packed.data_ptr(), nbytes, stream_ptr), "pm_pack_weights")
        self._packed, self._packed_key, self._staged = packed, key, staged
        return packed

    def _run(self, x: torch.Tensor, want_attn=False, want_feats=False):
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise ValueError("expected a (B, C, H, W) tensor")
        pdev = next(self.parameters()).device
        x = x.to(pdev)                                         # reference: x = x.to(self.device), models_PM.py:166
        if not x.is_cuda:
            raise RuntimeError("PM B200 forward needs a CUDA tensor / CUDA model: there is no CPU fallback")
        if torch.is_grad_enabled() and (x.requires_grad or self.training):
            # InstanceNorm has no running statistics, so train() and eval() compute the same forward; what the kernels cannot
            # give is an autograd graph — refuse the calls that would expect one
            raise RuntimeError("the PM B200 forward is inference-only: no autograd graph is recorded. Call model.eval() "
                               "and/or wrap the call in torch.no_grad()")
        if x.dtype in (torch.float16, torch.float64):
            x = x.float()                      # the reference module takes any float tensor; the kernels read fp32 or bf16
        if x.dtype == torch.bfloat16:
            dt = _lib.PMG_BF16
        elif x.dtype == torch.float32:
            dt = _lib.PMG_F32
        else:
            raise TypeError(f"input dtype {x.dtype} not supported (a floating-point tensor is expected)")
        B, Cin, H, W = x.shape
        if Cin != self._cfg.input_nc:
            raise ValueError(f"expected {self._cfg.input_nc} input channels, got {Cin}")
        x = x.contiguous()
        if x.data_ptr() % 16:                  # view at an odd storage offset: the kernels use 16-byte loads
            x = x.clone(memory_format=torch.contiguous_format)
        cfgp = ctypes.byref(self._cfg)
        hf, wf, cf = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        if _lib.lib.pm_feature_shape(cfgp, H, W, ctypes.byref(hf), ctypes.byref(wf), ctypes.byref(cf)) != 0:
            raise ValueError(f"input {tuple(x.shape)} is too small for this trunk")
        if B == 0:                             # empty batch: empty outputs, no launch (plain torch ops accept it)
            e = lambda *shape: torch.empty(shape, dtype=torch.float32, device=x.device)
            return (e(0, self._n_logits), e(0, len(self.tasks_dict), hf.value, wf.value) if want_attn else None,
                    e(0, hf.value, wf.value, cf.value) if want_feats else None)
        with self._lock, torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            packed = self._ensure_packed(x.device, stream)
            need = _lib.lib.pm_workspace_bytes(cfgp, B, H, W)
            if self._ws is None or self._ws.device != x.device or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            logits = torch.empty((B, self._n_logits), dtype=torch.float32, device=x.device)
            attn = torch.empty((B, len(self.tasks_dict), hf.value, wf.value), dtype=torch.float32, device=x.device) if want_attn else None
            feats = torch.empty((B, hf.value, wf.value, cf.value), dtype=torch.float32, device=x.device) if want_feats else None
            rc = _lib.lib.pm_forward(cfgp, packed.data_ptr(), x.data_ptr(), dt, B, H, W, logits.data_ptr(),
                                     attn.data_ptr() if attn is not None else None,
                                     feats.data_ptr() if feats is not None else None,
                                     self._ws.data_ptr(), self._ws.numel(),
                                     (0 if getattr(self, "implicit_gemm", True) else _lib.PM_FLAG_NO_IGEMM)
                                     | (_lib.PM_FLAG_IGEMM_N128 if getattr(self, "igemm_n128", False) else 0)
                                     | (0 if getattr(self, "fused_stats", True) else _lib.PM_FLAG_NO_FUSED_STATS)
                                     | (0 if getattr(self, "fused_conv0", True) else _lib.PM_FLAG_NO_FUSED_CONV0), stream)
            _lib.check(rc, "pm_forward")
        return logits, attn, feats

    def _split(self, logits):
        out, o = {}, 0
        for t, n in self.tasks_dict.items():
            out[t] = logits[:, o:o + n]
            o += n
        return out

    def forward_packed(self, x) -> torch.Tensor:
        return self._run(x)[0]

    def forward(self, x: torch.Tensor, return_embeddings: bool = False, return_task_embeddings: bool = False,
                return_full: bool = False) -> Any:
        if return_task_embeddings:                                             # models_PM.py:169-176
            logits, _, feats = self._run(x, want_feats=True)
            emb = feats.mean(dim=(1, 2)).cpu()
            return self._split(logits), {t: emb for t in self.tasks_dict}
        if return_embeddings:                                                  # :178-181 (NCHW flatten order)
            _, _, feats = self._run(x, want_feats=True)
            return feats.permute(0, 3, 1, 2).reshape(feats.shape[0], -1).cpu()
        if return_full:                                                        # :183-188
            logits, attn, _ = self._run(x, want_attn=True)
            sp = self._split(logits)
            return {t: {"logits": sp[t], "attn": attn[:, i:i + 1]} for i, t in enumerate(self.tasks_dict)}
        return self._split(self._run(x)[0])


_REFERENCE_HELPERS = ("TaskSpecificModel", "load_model_weights", "checkpoint_has_se", "print_model_parameters")


def __getattr__(name):
    """``TaskSpecificModel`` / ``load_model_weights`` / ``checkpoint_has_se`` / ``print_model_parameters`` (reference
    models_PM.py:197-253, imported by test_PM.py:16-21) are host-side helpers that do not touch the forward path.  They
    are not restated here: ``compat.install_as_reference_models()`` binds the reference's own definitions onto this
    module when its checkout is importable."""
    if name in _REFERENCE_HELPERS:
        from . import compat
        if compat.bind_reference_pm_helpers():
            return globals()[name]
        raise AttributeError(f"{name} lives in the reference's Models/models_PM.py; put the reference checkout on sys.path "
                             "(python -m pmg_b200.compat <reference>/test_PM.py ...) so it can be bound")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
