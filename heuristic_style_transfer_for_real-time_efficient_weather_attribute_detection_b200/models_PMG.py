"""Drop-in replacement for the reference module ``Models/models_PMG.py``.

``MultiTaskPatchGANGramModelNonOverlapV2`` keeps the reference constructor signature
(Models/models_PMG.py:78-92), registers exactly the same parameters / buffers under the same names
(the 81 keys of Model_weights/PMG/hyperparameters_PMG.json:53-135, so ``load_state_dict(strict=True)``
from utils/PMG/PMG_model_utils.py:187 works), exposes the attributes the reference's callers read
(``feature_extractor``, ``se``, ``chan_proj``, ``token_proj``, ``transformer``, ``task_pool.query``,
``classifiers``, ``patch_div``, ``num_classes_per_task`` …) and returns the same
``Dict[task, Tensor(B, ncls)]`` from ``forward``.

What differs is the inside of ``forward`` (reference :176-199): the sub-modules are parameter
holders only; the arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/pmg_b200.h`` (``libpmg_b200.so``).  There is no CPU path and no eager-PyTorch fallback: a
CPU tensor, train mode, or a missing library raises.
"""
from __future__ import annotations

import ctypes
import threading
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib

__all__ = ["MultiTaskPatchGANGramModelNonOverlapV2", "TaskAttentionPooling", "TransformerBlock", "SEBlock"]


class SEBlock(nn.Module):
    """Parameter holder of the reference's channel attention (models_PMG.py:15-30: global average pool → Linear(C, C/16,
    no bias) → ReLU → Linear(C/16, C, no bias) → Sigmoid → x·w); same state_dict keys (``se.fc.0.weight``,
    ``se.fc.2.weight``).  The arithmetic runs in the CUDA library (shape-generic fp32 kernels; off in the committed config)."""

    def __init__(self, C: int, r: int = 16):
        super().__init__()
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(C, max(1, C // r), bias=False), nn.ReLU(True),
                                nn.Linear(max(1, C // r), C, bias=False), nn.Sigmoid())

    def forward(self, x):  # pragma: no cover - parameter holder
        raise RuntimeError("SEBlock is fused into the PMG CUDA forward; call the model instead")


class TaskAttentionPooling(nn.Module):
    """Holder of the per-task query vectors (reference models_PMG.py:36-57).

    The pooling itself (s = tok·q_t, softmax over tokens, weighted sum — no 1/sqrt(d)) is fused with
    the classifier heads inside the CUDA library."""

    def __init__(self, d_model: int, tasks: Dict[str, int]):
        super().__init__()
        self.tasks = list(tasks.keys())
        self.query = nn.ParameterDict({name: nn.Parameter(torch.randn(d_model)) for name in self.tasks})

    def forward(self, tokens):  # pragma: no cover - parameter holder
        raise RuntimeError("TaskAttentionPooling is fused into the PMG CUDA forward; call the model instead")


class TransformerBlock(nn.Module):
    """Holder of the shared pre-LN encoder stack (reference models_PMG.py:60-72)."""

    def __init__(self, d_model=256, nhead=4, dim_feedforward=512, dropout=0.1, num_layers=1):
        super().__init__()
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.encoder = nn.TransformerEncoder(
                nn.TransformerEncoderLayer(d_model=d_model, nhead=nhead, dim_feedforward=dim_feedforward,
                                           dropout=dropout, batch_first=True, norm_first=True),
                num_layers=num_layers)

    def forward(self, x):  # pragma: no cover - parameter holder
        raise RuntimeError("TransformerBlock is fused into the PMG CUDA forward; call the model instead")


def _backbone_plan(input_nc: int, ndf: int, patch_size: int):
    """(cin, cout, k) per conv — the reference's layer-count loop, models_PMG.py:107-128."""
    plan, cin, width, rf = [], input_nc, ndf, patch_size
    while rf > 4 and width <= 512:
        plan.append((cin, width, 4))
        cin, width, rf = width, width * 2, rf / 4
    plan.append((cin, width, 1))
    return plan


def _epoch_hook(module, incompatible_keys):   # module-level (picklable) load_state_dict post hook
    module._bump_epoch()


class MultiTaskPatchGANGramModelNonOverlapV2(nn.Module):
    def __init__(
        self,
        input_nc: int = 3,
        ndf: int = 64,
        norm: str = "batch",
        patch_size: int = 70,
        patch_div: int = 4,
        num_classes_per_task: Dict[str, int] | None = None,
        use_channel_attention: bool = False,
        use_token_attention: bool = True,
        gram_channels: int = 64,
        d_model: int = 256,
        transformer_layers: int = 1,
        transformer_heads: int = 4,
    ):
        super().__init__()
        self.num_classes_per_task = num_classes_per_task or {}
        if not self.num_classes_per_task:
            raise ValueError("num_classes_per_task must name at least one task")
        self.patch_div = patch_div
        self.use_channel_attention = use_channel_attention
        self.use_token_attention = use_token_attention
        self.gram_channels = gram_channels

        # models_PMG.py:100-105: "instance" → InstanceNorm2d(affine=True); "group" → partial(nn.GroupNorm, num_groups=32) called
        # as norm_layer(num_filters), which binds num_groups twice: the reference raises TypeError at construction, and so does
        # this module; anything else means BatchNorm2d
        if norm == "group":
            raise TypeError("GroupNorm.__init__() got multiple values for argument 'num_groups' (norm='group' cannot be "
                            "constructed in the reference either: Models/models_PMG.py:102-103,116)")
        self._norm = "instance" if norm == "instance" else "batch"
        stack = []
        for cin, cout, k in _backbone_plan(input_nc, ndf, patch_size):
            stack.append(nn.Conv2d(cin, cout, kernel_size=k, stride=k, padding=0, bias=False))
            if self._norm == "instance":
                stack.append(nn.InstanceNorm2d(cout, affine=True))
            else:
                stack.append(nn.BatchNorm2d(cout, affine=True))
            stack.append(nn.LeakyReLU(0.2, inplace=True))
        self.feature_extractor = nn.Sequential(*stack)
        self.C = stack[-3].out_channels

        self.se = SEBlock(self.C) if use_channel_attention else nn.Identity()
        self.chan_proj = nn.Conv2d(self.C, gram_channels, kernel_size=1, bias=False)
        self.Cr = gram_channels
        self.token_proj = nn.Linear(self.Cr * self.Cr, d_model)
        n_tf = transformer_layers if (use_token_attention and transformer_layers > 0) else 0
        self.transformer = (TransformerBlock(d_model=d_model, nhead=transformer_heads,
                                             dim_feedforward=2 * d_model, dropout=0.1, num_layers=n_tf)
                            if n_tf > 0 else nn.Identity())
        self.task_pool = TaskAttentionPooling(d_model=d_model, tasks=self.num_classes_per_task)
        self.classifiers = nn.ModuleDict({t: nn.Linear(d_model, n) for t, n in self.num_classes_per_task.items()})

        # ---- kernel-side state: never part of state_dict ----
        self._cfg = _lib.make_config(input_nc, ndf, int(patch_size), patch_div, gram_channels, d_model, n_tf,
                                     transformer_heads, list(self.num_classes_per_task.values()), self._norm,
                                     use_channel_attention)
        if _lib.lib.pmg_param_count(ctypes.byref(self._cfg)) < 0:
            raise ValueError("configuration rejected by the PMG CUDA library (pmg_param_count < 0)")
        self._n_logits = _lib.lib.pmg_logits_per_image(ctypes.byref(self._cfg))
        self._packed: Optional[torch.Tensor] = None
        self._packed_key: Optional[Tuple] = None
        self._ws: Optional[torch.Tensor] = None
        self._lock = threading.Lock()
        self._weights_epoch = 0         # bumped by load_state_dict / .to(): captured CUDA graphs (runner.GraphedPMG) re-capture
        self.register_load_state_dict_post_hook(_epoch_hook)
        self.force_generic = False      # tests: route through the shape-generic kernels
        self.fused_front = True         # fast path: TMA-fed fused conv0+conv1 kernel (fp32 input: + hi/lo plane split); False = im2col + GEMMs
        self.fused_encoder = True       # fast path: transformer + pooling + heads in one kernel; False = per-op kernels
        self.fused_mid = True           # fast path: conv2 + chan_proj in one kernel; False = two GEMMs
        self.fused_midgram = True       # fast path: conv2 + chan_proj + patches + Gram in one kernel from 96 images up; "always" = at every
                                        # batch size; False = mid kernel + Gram kernel

    # kernel-side state is rebuilt on demand; keep it out of pickles / deepcopies
    def __getstate__(self):
        st = dict(self.__dict__)
        for k in ("_packed", "_packed_key", "_ws", "_lock", "_staged", "_abi_cache"):
            st.pop(k, None)
        c = st.pop("_cfg")
        st["_cfg_tuple"] = (c.input_nc, c.ndf, c.patch_size, c.patch_div, c.gram_channels, c.d_model,
                            c.transformer_layers, c.transformer_heads, [c.ncls[i] for i in range(c.n_tasks)],
                            c.norm, c.use_channel_attention)
        return st

    def __setstate__(self, st):
        st = dict(st)
        cfg = st.pop("_cfg_tuple")
        super().__setstate__(st)
        self._cfg = _lib.make_config(*cfg)
        self._packed = self._packed_key = self._ws = None
        self._lock = threading.Lock()

    def _bump_epoch(self):
        self._weights_epoch = getattr(self, "_weights_epoch", 0) + 1

    def _apply(self, fn, *args, **kwargs):
        self._bump_epoch()              # .to() / .cuda() / .float(): parameter storage may move
        return super()._apply(fn, *args, **kwargs)

    # ------------------------------------------------------------------ weight packing
    def _abi_params(self):
        """Parameter tensors in the order include/pmg_b200.h documents for pmg_pack_weights."""
        out = []
        mods = list(self.feature_extractor)
        for i in range(0, len(mods), 3):
            conv, bn = mods[i], mods[i + 1]
            out += [conv.weight, bn.weight, bn.bias]
            if self._norm == "batch":
                out += [bn.running_mean, bn.running_var]
        if self.use_channel_attention:
            out += [self.se.fc[0].weight, self.se.fc[2].weight]
        out += [self.chan_proj.weight, self.token_proj.weight, self.token_proj.bias]
        if not isinstance(self.transformer, nn.Identity):
            for layer in self.transformer.encoder.layers:
                out += [layer.self_attn.in_proj_weight, layer.self_attn.in_proj_bias,
                        layer.self_attn.out_proj.weight, layer.self_attn.out_proj.bias,
                        layer.linear1.weight, layer.linear1.bias, layer.linear2.weight, layer.linear2.bias,
                        layer.norm1.weight, layer.norm1.bias, layer.norm2.weight, layer.norm2.bias]
        out += [self.task_pool.query[t] for t in self.num_classes_per_task]
        for t in self.num_classes_per_task:
            out += [self.classifiers[t].weight, self.classifiers[t].bias]
        return out

    def invalidate_packed(self) -> None:
        """Call after REPLACING a Parameter object of the module (``m.token_proj.weight = nn.Parameter(…)``).  In-place updates,
        ``load_state_dict`` and ``.to()`` are noticed without it (tensor version counters / storage pointers / the epoch)."""
        self._bump_epoch()

    def _ensure_packed(self, device: torch.device, stream_ptr: int) -> torch.Tensor:
        # the walk over the module tree is half of this function's host time (the other half: 78 version / pointer reads): the
        # Parameter objects themselves only change identity on explicit replacement, so the list is kept per weights epoch
        c = getattr(self, "_abi_cache", None)
        if c is None or c[0] != self._weights_epoch:
            c = self._abi_cache = (self._weights_epoch, self._abi_params())
        params = c[1]
        key = (device, tuple((p.data_ptr(), p._version) for p in params))
        if self._packed is not None and key == self._packed_key:
            return self._packed
        staged = []
        for p in params:
            if p.device != device:
                raise RuntimeError(f"model parameters live on {p.device} but the input is on {device}; "
                                   "call model.to(device) first")
            t = p.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            staged.append(t)
        n = len(staged)
        ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in staged])
        nbytes = _lib.lib.pmg_packed_bytes(ctypes.byref(self._cfg))
        if nbytes == 0:
            raise _lib.PmgError("pmg_packed_bytes: unsupported configuration")
        packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _lib.check(_lib.lib.pmg_pack_weights(ctypes.byref(self._cfg), ptrs, n, packed.data_ptr(), nbytes, stream_ptr),
                   "pmg_pack_weights")
        self._packed, self._packed_key = packed, key
        self._staged = staged   # keep alive until the pack kernels have run
        return packed

    # ------------------------------------------------------------------ forward
    def _run(self, x: torch.Tensor, want_gram: bool = False, taps: bool = False, out: Optional[torch.Tensor] = None,
             ws: Optional[torch.Tensor] = None, also=None):
        """``ws``: caller-owned workspace (uint8, >= pmg_workspace_bytes) used instead of the module's own — a captured
        CUDA graph must keep replaying into memory nobody reallocates (runner.GraphedPMG)."""
        if self.training:
            raise RuntimeError("PMG B200 kernels implement the eval forward only (BatchNorm running stats, "
                               "dropout inert): call model.eval()")
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise ValueError("expected a (B, C, H, W) tensor")
        if not x.is_cuda:
            raise RuntimeError("PMG B200 forward needs a CUDA tensor: there is no CPU fallback")
        if torch.is_grad_enabled() and x.requires_grad:
            raise RuntimeError("the PMG B200 forward is inference-only: no autograd graph is recorded "
                               "(input requires grad). Use torch.no_grad().")
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
        if B == 0:                             # empty batch (e.g. an empty last DataLoader batch): empty outputs, no launch
            np_ = self.patch_div * self.patch_div
            logits = out if out is not None else torch.empty((0, self._n_logits), dtype=torch.float32, device=x.device)
            gram = torch.empty((0, np_, self.Cr, self.Cr), dtype=torch.float32, device=x.device) if want_gram else None
            return logits, gram, {}
        x = x.contiguous()
        if x.data_ptr() % 16:                  # a view at an odd storage offset: the kernels use 16-byte loads / TMA
            x = x.clone(memory_format=torch.contiguous_format)
        cfgp = ctypes.byref(self._cfg)
        with self._lock, torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            packed = self._ensure_packed(x.device, stream)
            need = _lib.lib.pmg_workspace_bytes(cfgp, B, H, W)
            if need == 0:
                raise ValueError(f"input {tuple(x.shape)} is too small for this backbone / patch split")
            if ws is not None:
                if ws.device != x.device or ws.dtype != torch.uint8 or ws.numel() < need or ws.data_ptr() % 16:
                    raise ValueError(f"ws must be a 16-byte aligned uint8 tensor of >= {need} bytes on {x.device}")
            else:
                if self._ws is None or self._ws.device != x.device or self._ws.numel() < need:
                    self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
                ws = self._ws
            if out is not None:
                if (out.shape != (B, self._n_logits) or out.dtype != torch.float32 or out.device != x.device
                        or not out.is_contiguous()):
                    raise ValueError(f"out must be a contiguous ({B}, {self._n_logits}) float32 tensor on {x.device}")
                logits = out
            else:
                logits = torch.empty((B, self._n_logits), dtype=torch.float32, device=x.device)
            gram = None
            extra = {}
            tp = None
            np_ = self.patch_div * self.patch_div
            if want_gram:
                gram = torch.empty((B, np_, self.Cr, self.Cr), dtype=torch.float32, device=x.device)
            if taps:
                hc, wc = H, W
                for m in self.feature_extractor:
                    if isinstance(m, nn.Conv2d) and m.kernel_size[0] == 4:
                        hc, wc = hc // 4, wc // 4
                d = self.token_proj.out_features
                extra = {"chan": torch.empty((B, hc, wc, self.Cr), dtype=torch.float32, device=x.device),
                         "tokens_in": torch.empty((B, np_, d), dtype=torch.float32, device=x.device),
                         "tokens_out": torch.empty((B, np_, d), dtype=torch.float32, device=x.device)}
                tp = _lib.PmgTaps(extra["chan"].data_ptr(), extra["tokens_in"].data_ptr(), extra["tokens_out"].data_ptr())
            flags = self._flags()
            if also is not None and (hasattr(also, "ptrs") or len(also)):
                # further destinations of the logits (peer-mapped buffers of other GPUs): stored by the token-stage kernel itself
                if want_gram or taps:
                    raise ValueError("`also` cannot be combined with gram / taps outputs")
                if hasattr(also, "ptrs"):                  # runner.PeerSlots: validated once, pointer array ready
                    if also.shape != (B, self._n_logits):
                        raise ValueError(f"`also` slots are {also.shape}, the logits ({B}, {self._n_logits})")
                    ptrs, n_also = also.ptrs, also.n
                else:
                    for t in also:
                        if t.shape != (B, self._n_logits) or t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
                            raise ValueError(f"every `also` tensor must be a contiguous CUDA ({B}, {self._n_logits}) float32 tensor")
                    ptrs, n_also = (ctypes.c_void_p * len(also))(*[t.data_ptr() for t in also]), len(also)
                rc = _lib.lib.pmg_forward_bcast(cfgp, packed.data_ptr(), x.data_ptr(), dt, B, H, W, logits.data_ptr(), ptrs, n_also,
                                                ws.data_ptr(), ws.numel(), flags, stream)
                _lib.check(rc, "pmg_forward_bcast")
                return logits, gram, extra
            rc = _lib.lib.pmg_forward(cfgp, packed.data_ptr(), x.data_ptr(), dt, B, H, W, logits.data_ptr(),
                                      gram.data_ptr() if gram is not None else None,
                                      ctypes.byref(tp) if tp is not None else None,
                                      ws.data_ptr(), ws.numel(), flags, stream)
            _lib.check(rc, "pmg_forward")
        return logits, gram, extra

    def _split(self, logits: torch.Tensor) -> Dict[str, torch.Tensor]:
        out, o = {}, 0
        for t, n in self.num_classes_per_task.items():
            out[t] = logits[:, o:o + n]
            o += n
        return out

    def forward(self, x):
        logits, _, _ = self._run(x)
        return self._split(logits)

    def forward_packed(self, x, out: Optional[torch.Tensor] = None, ws: Optional[torch.Tensor] = None, also=None) -> torch.Tensor:
        """(B, Σncls) fp32 — the dict of ``forward`` as one tensor, columns in Tasks.json order.  ``out``: write the
        logits into this (B, Σncls) fp32 tensor (e.g. a slot of a gather bucket, runner.BucketedGather); ``ws``: a
        caller-owned workspace of ``workspace_bytes(...)`` bytes (graph capture); ``also``: up to 15 more (B, Σncls) fp32
        tensors — typically peer-mapped slots on the other GPUs (runner.PushBucketedGather) — that receive the same logits,
        written by the token-stage kernel's own epilogue (``pmg_forward_bcast``)."""
        return self._run(x, out=out, ws=ws, also=also)[0]

    def workspace_bytes(self, B: int, H: int = 224, W: int = 224) -> int:
        return int(_lib.lib.pmg_workspace_bytes(ctypes.byref(self._cfg), B, H, W))

    def forward_with_gram(self, x):
        """(dict, G) with G = per-patch Gram (B, patch_div², Cr, Cr) of reference line :188."""
        logits, gram, _ = self._run(x, want_gram=True)
        return self._split(logits), gram

    @torch.no_grad()
    def style_patch_grams(self, x: torch.Tensor) -> torch.Tensor:
        """Forward-only Gram export of the patch style transfer (utils/PMG/PMG_style_transfer.py:11-48,
        ``extract_patch_grams_for_image``) for a whole batch: ``feature_extractor(x)`` cropped to ``(Hf // ph)·ph x (Wf // pw)·pw``
        (``ph = Hf // patch_div``, no padding), cut — exactly as the reference's permute + reshape does (:37) — into runs of
        ``ph·pw`` consecutive row-major pixels, ``G = F·Fᵀ / (ph·pw)`` per run.  Returns (B, n_patches, C, C) fp32 —
        ``list(out[0])`` is what the reference's function returns for image 0."""
        if self.training:
            raise RuntimeError("PMG B200 kernels implement the eval forward only: call model.eval()")
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise ValueError("expected a (B, C, H, W) tensor")
        if not x.is_cuda:
            raise RuntimeError("PMG B200 forward needs a CUDA tensor: there is no CPU fallback")
        if x.dtype in (torch.float16, torch.float64):
            x = x.float()
        dt = {torch.bfloat16: _lib.PMG_BF16, torch.float32: _lib.PMG_F32}.get(x.dtype)
        if dt is None:
            raise TypeError(f"input dtype {x.dtype} not supported (a floating-point tensor is expected)")
        B, Cin, H, W = x.shape
        if Cin != self._cfg.input_nc:
            raise ValueError(f"expected {self._cfg.input_nc} input channels, got {Cin}")
        cfgp = ctypes.byref(self._cfg)
        n, c = ctypes.c_int(), ctypes.c_int()
        if _lib.lib.pmg_style_gram_shape(cfgp, H, W, ctypes.byref(n), ctypes.byref(c)) != 0:
            raise ValueError(f"input {tuple(x.shape)} is too small for this backbone / patch split")
        out = torch.empty((B, n.value, c.value, c.value), dtype=torch.float32, device=x.device)
        if B == 0:
            return out
        x = x.contiguous()
        if x.data_ptr() % 16:
            x = x.clone(memory_format=torch.contiguous_format)
        with self._lock, torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            packed = self._ensure_packed(x.device, stream)
            need = _lib.lib.pmg_workspace_bytes(cfgp, B, H, W)
            if self._ws is None or self._ws.device != x.device or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            _lib.check(_lib.lib.pmg_style_grams(cfgp, packed.data_ptr(), x.data_ptr(), dt, B, H, W, out.data_ptr(),
                                                self._ws.data_ptr(), self._ws.numel(), stream), "pmg_style_grams")
        return out

    def forward_debug(self, x):
        """Test hook: (packed logits, gram, {chan, tokens_in, tokens_out})."""
        return self._run(x, want_gram=True, taps=True)

    def _flags(self) -> int:
        return ((_lib.PMG_FLAG_FORCE_GENERIC if self.force_generic else 0)
                | (0 if self.fused_front else _lib.PMG_FLAG_NO_FUSED_FRONT)
                | (0 if getattr(self, "fused_encoder", True) else _lib.PMG_FLAG_NO_FUSED_ENCODER)
                | (0 if getattr(self, "fused_mid", True) else _lib.PMG_FLAG_NO_FUSED_MID)
                | (0 if getattr(self, "fused_midgram", True) else _lib.PMG_FLAG_NO_FUSED_MIDGRAM)
                | (_lib.PMG_FLAG_ALWAYS_FUSED_MIDGRAM if getattr(self, "fused_midgram", True) == "always" else 0))

    def launches_per_forward(self, B: int, H: int = 224, W: int = 224, dtype: torch.dtype = torch.bfloat16) -> int:
        flags = self._flags() | (0 if dtype == torch.bfloat16 else _lib.PMG_FLAG_COUNT_F32_INPUT)
        return _lib.lib.pmg_forward_launch_count(ctypes.byref(self._cfg), B, H, W, flags)
