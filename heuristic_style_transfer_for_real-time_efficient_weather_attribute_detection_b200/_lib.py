"""ctypes binding of include/pmg_b200.h.  The library is REQUIRED: there is no CPU or eager fallback —
if it is missing, importing this module raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpmg_b200.so")

PMG_MAX_TASKS = 32
PMG_F32, PMG_BF16 = 0, 1
PMG_FLAG_FORCE_GENERIC, PMG_FLAG_FORCE_FAST, PMG_FLAG_NO_FUSED_FRONT, PMG_FLAG_NO_FUSED_ENCODER, PMG_FLAG_NO_FUSED_MID = 1, 2, 4, 8, 16
PMG_FLAG_NO_FUSED_MIDGRAM, PMG_FLAG_ALWAYS_FUSED_MIDGRAM, PMG_FLAG_COUNT_F32_INPUT = 32, 64, 128
PM_FLAG_NO_IGEMM = 1
PM_FLAG_IGEMM_N128 = 2
PM_FLAG_NO_FUSED_STATS = 4
PM_FLAG_NO_FUSED_CONV0 = 8


class PmgConfig(C.Structure):
    _fields_ = [
        ("input_nc", C.c_int32), ("ndf", C.c_int32), ("patch_size", C.c_int32), ("patch_div", C.c_int32),
        ("gram_channels", C.c_int32), ("d_model", C.c_int32), ("transformer_layers", C.c_int32),
        ("transformer_heads", C.c_int32), ("n_tasks", C.c_int32), ("ncls", C.c_int32 * PMG_MAX_TASKS),
        ("norm", C.c_int32), ("use_channel_attention", C.c_int32),
    ]


class PmgTaps(C.Structure):
    _fields_ = [("chan", C.c_void_p), ("tokens_in", C.c_void_p), ("tokens_out", C.c_void_p)]


class PmConfig(C.Structure):
    _fields_ = [("input_nc", C.c_int32), ("ndf", C.c_int32), ("patch_size", C.c_float), ("n_tasks", C.c_int32),
                ("ncls", C.c_int32 * PMG_MAX_TASKS), ("attn_tau", C.c_float), ("ablate_attention", C.c_int32)]


class PmgStageTime(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ms", C.c_float)]


class PmgError(RuntimeError):
    pass


EXPORTS = [
    "pmg_abi_version", "pmg_status_string", "pmg_last_cuda_error", "pmg_param_count", "pmg_packed_bytes",
    "pmg_pack_weights", "pmg_workspace_bytes", "pmg_logits_per_image", "pmg_forward_launch_count",
    "pmg_forward", "pmg_forward_bcast", "pmg_forward_profile", "pmg_gram_batched", "pmg_tc_gemm", "pmg_debug_set_trace",
    "pmg_preprocess_u8", "pmg_postprocess", "pmg_style_gram_shape", "pmg_style_grams", "pmg_resize_scratch_bytes", "pmg_resize_crop_u8", "pmg_debug_tmem_a", "pmg_tc_gemm2", "pmg_debug_load_probe", "pm_debug_set_chunking",
    "pm_param_count", "pm_packed_bytes", "pm_pack_weights", "pm_workspace_bytes", "pm_feature_shape", "pm_forward",
]


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The PMG B200 path has no CPU / eager fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.pmg_abi_version.restype = C.c_int
    lib.pmg_status_string.restype = C.c_char_p
    lib.pmg_status_string.argtypes = [C.c_int]
    lib.pmg_last_cuda_error.restype = C.c_char_p
    cfgp = C.POINTER(PmgConfig)
    lib.pmg_param_count.argtypes = [cfgp]
    lib.pmg_param_count.restype = C.c_int
    lib.pmg_packed_bytes.argtypes = [cfgp]
    lib.pmg_packed_bytes.restype = C.c_size_t
    lib.pmg_pack_weights.argtypes = [cfgp, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.pmg_pack_weights.restype = C.c_int
    lib.pmg_workspace_bytes.argtypes = [cfgp, C.c_int, C.c_int, C.c_int]
    lib.pmg_workspace_bytes.restype = C.c_size_t
    lib.pmg_logits_per_image.argtypes = [cfgp]
    lib.pmg_logits_per_image.restype = C.c_int
    lib.pmg_forward_launch_count.argtypes = [cfgp, C.c_int, C.c_int, C.c_int, C.c_uint]
    lib.pmg_forward_launch_count.restype = C.c_int
    lib.pmg_forward.argtypes = [cfgp, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                C.c_void_p, C.POINTER(PmgTaps), C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p]
    lib.pmg_forward.restype = C.c_int
    lib.pmg_forward_bcast.argtypes = [cfgp, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p]
    lib.pmg_forward_bcast.restype = C.c_int
    lib.pmg_forward_profile.argtypes = [cfgp, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.POINTER(PmgStageTime),
                                        C.c_int, C.POINTER(C.c_int)]
    lib.pmg_forward_profile.restype = C.c_int
    lib.pmg_gram_batched.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    lib.pmg_gram_batched.restype = C.c_int
    lib.pmg_style_gram_shape.argtypes = [cfgp, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.pmg_style_gram_shape.restype = C.c_int
    lib.pmg_style_grams.argtypes = [cfgp, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_size_t, C.c_void_p]
    lib.pmg_style_grams.restype = C.c_int
    lib.pmg_tc_gemm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_void_p, C.c_void_p]
    lib.pmg_tc_gemm.restype = C.c_int
    lib.pmg_preprocess_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                      C.c_int, C.c_void_p, C.c_void_p]
    lib.pmg_preprocess_u8.restype = C.c_int
    lib.pmg_resize_scratch_bytes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    lib.pmg_resize_scratch_bytes.restype = C.c_size_t
    lib.pmg_resize_crop_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                       C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.pmg_resize_crop_u8.restype = C.c_int
    lib.pmg_postprocess.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_float, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
    lib.pmg_postprocess.restype = C.c_int
    lib.pmg_debug_tmem_a.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pmg_debug_tmem_a.restype = C.c_int
    lib.pmg_debug_load_probe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p]
    lib.pmg_debug_load_probe.restype = C.c_int
    lib.pmg_tc_gemm2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.pmg_tc_gemm2.restype = C.c_int
    lib.pm_debug_set_chunking.argtypes = [C.c_int, C.c_int]
    lib.pm_debug_set_chunking.restype = None
    lib.pmg_debug_set_trace.argtypes = [C.c_void_p]
    lib.pmg_debug_set_trace.restype = None
    pcfg = C.POINTER(PmConfig)
    lib.pm_param_count.argtypes = [pcfg]
    lib.pm_param_count.restype = C.c_int
    lib.pm_packed_bytes.argtypes = [pcfg]
    lib.pm_packed_bytes.restype = C.c_size_t
    lib.pm_pack_weights.argtypes = [pcfg, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.pm_pack_weights.restype = C.c_int
    lib.pm_workspace_bytes.argtypes = [pcfg, C.c_int, C.c_int, C.c_int]
    lib.pm_workspace_bytes.restype = C.c_size_t
    lib.pm_feature_shape.argtypes = [pcfg, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.pm_feature_shape.restype = C.c_int
    lib.pm_forward.argtypes = [pcfg, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p]
    lib.pm_forward.restype = C.c_int
    if lib.pmg_abi_version() != 2:
        raise ImportError("libpmg_b200.so ABI version mismatch — rebuild")
    return lib


lib = _load()


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib.pmg_status_string(status).decode()
        cuda = lib.pmg_last_cuda_error().decode()
        raise PmgError(f"{what}: {msg}" + (f" [{cuda}]" if status == 5 and cuda else ""))


def make_pm_config(input_nc, ndf, patch_size, ncls_list, attn_tau, ablate) -> PmConfig:
    if len(ncls_list) > PMG_MAX_TASKS:
        raise PmgError(f"at most {PMG_MAX_TASKS} tasks are supported")
    cfg = PmConfig()
    cfg.input_nc, cfg.ndf, cfg.patch_size = input_nc, ndf, float(patch_size)
    cfg.n_tasks = len(ncls_list)
    for i, n in enumerate(ncls_list):
        cfg.ncls[i] = int(n)
    cfg.attn_tau, cfg.ablate_attention = float(attn_tau), int(bool(ablate))
    return cfg


PMG_NORM = {"batch": 0, "instance": 1}


def make_config(input_nc, ndf, patch_size, patch_div, gram_channels, d_model, transformer_layers,
                transformer_heads, ncls_list, norm="batch", use_channel_attention=False) -> PmgConfig:
    if len(ncls_list) > PMG_MAX_TASKS:
        raise PmgError(f"at most {PMG_MAX_TASKS} tasks are supported")
    cfg = PmgConfig()
    cfg.input_nc, cfg.ndf, cfg.patch_size, cfg.patch_div = input_nc, ndf, patch_size, patch_div
    cfg.gram_channels, cfg.d_model = gram_channels, d_model
    cfg.transformer_layers, cfg.transformer_heads = transformer_layers, transformer_heads
    cfg.n_tasks = len(ncls_list)
    for i, n in enumerate(ncls_list):
        cfg.ncls[i] = int(n)
    cfg.norm = PMG_NORM[norm] if isinstance(norm, str) else int(norm)
    cfg.use_channel_attention = int(bool(use_channel_attention))
    return cfg
