"""The C-ABI library loads and exports every symbol include/pmg_b200.h declares; argument validation
that needs no GPU."""
import ctypes
import os
import re

import pytest

from helpers import ROOT, load_contract


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "pmg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pmg?_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    from pmg_b200 import _lib
    names = _declared_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(_lib.lib, n), n
    assert sorted(_lib.EXPORTS) == names
    assert _lib.lib.pmg_abi_version() == 2


def test_library_is_in_tree_shared_object():
    from pmg_b200 import _lib
    assert os.path.dirname(_lib.LIB_PATH).startswith(ROOT)
    assert os.path.getsize(_lib.LIB_PATH) > 50_000


def _cfg():
    from pmg_b200 import _lib
    _, hp, tasks = load_contract()
    return _lib, _lib.make_config(3, hp["ndf"], hp["patch_size"], hp["patch_div"], hp["gram_channels"], hp["d_model"],
                                  hp["transformer_layers"], hp["transformer_heads"], list(tasks.values()))


def test_host_side_queries_committed_config():
    _lib, cfg = _cfg()
    p = ctypes.byref(cfg)
    # 81 state_dict keys − 3 num_batches_tracked = 78 float tensors
    assert _lib.lib.pmg_param_count(p) == 78
    assert _lib.lib.pmg_logits_per_image(p) == 52
    assert _lib.lib.pmg_packed_bytes(p) > 2_431_284 * 4
    assert _lib.lib.pmg_workspace_bytes(p, 8, 224, 224) > 0
    assert _lib.lib.pmg_workspace_bytes(p, 8, 8, 8) == 0          # too small for two k4s4 layers
    assert _lib.lib.pmg_workspace_bytes(p, 0, 224, 224) == 0
    assert _lib.lib.pmg_forward_launch_count(p, 8, 224, 224, _lib.PMG_FLAG_FORCE_GENERIC) == 3 + 3 + 14 + 1
    # fast path: front | conv2+chan_proj+Gram (one kernel from 96 images up, mid + Gram kernels below) | token_proj | token stage
    lc = lambda B, flags=0: _lib.lib.pmg_forward_launch_count(p, B, 224, 224, flags)
    assert lc(1024) == 4 and lc(96) == 4 and lc(95) == 5 and lc(1) == 5
    assert lc(1, _lib.PMG_FLAG_ALWAYS_FUSED_MIDGRAM) == 4
    assert lc(1024, _lib.PMG_FLAG_NO_FUSED_MIDGRAM) == 5
    assert lc(1024, _lib.PMG_FLAG_NO_FUSED_MID) == 6
    assert lc(1024, _lib.PMG_FLAG_NO_FUSED_FRONT) == 6
    assert lc(1024, _lib.PMG_FLAG_COUNT_F32_INPUT) == 5          # fp32 input: the hi/lo plane split in front of the fused kernel
    assert lc(1024, _lib.PMG_FLAG_COUNT_F32_INPUT | _lib.PMG_FLAG_NO_FUSED_FRONT) == 6


def test_bad_arguments_are_status_codes_not_crashes():
    _lib, cfg = _cfg()
    p = ctypes.byref(cfg)
    assert _lib.lib.pmg_forward(p, None, None, 0, 1, 224, 224, None, None, None, None, 0, 0, None) == 1
    assert _lib.lib.pmg_pack_weights(p, None, 78, None, 0, None) == 1
    assert _lib.lib.pmg_gram_batched(None, 0, 1, 1, 1, 1.0, None, None) == 1
    # pointer alignment is validated before anything is launched (16 B for the image / workspace / packed weights)
    need = _lib.lib.pmg_packed_bytes(p)
    assert _lib.lib.pmg_forward(p, 4096, 4096 + 2, _lib.PMG_BF16, 1, 224, 224, 8192, None, None, 16384, 1 << 30, 0, None) == 1
    assert need > 0
    bad = _lib.make_config(3, 32, 64, 3, 96, 190, 2, 4, [2])       # d_model not divisible by heads
    assert _lib.lib.pmg_param_count(ctypes.byref(bad)) == -1
    assert _lib.lib.pmg_status_string(4) == b"device is not sm_100 (B200)"
    with pytest.raises(_lib.PmgError):
        _lib.check(2, "x")
