#!/usr/bin/env python
"""How fast can the SMs pull the image batch into shared memory?  (pmg_debug_load_probe, no compute.)
    python tools/load_probe.py        -> one JSON line per (mode, stages)"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pmg_b200 import _lib

B = 1024
x = torch.randn(B, 3, 224, 224, device="cuda").to(torch.bfloat16)
names = {0: "4-D tensor box 224x8x4 (448 B rows, 16 rows apart)", 1: "contiguous 14 KB cp.async.bulk", 2: "2-D tensor box 224x32 (contiguous rows)",
         3: "4-D tensor box, no evict-first hint"}
for mode in (0, 3, 2, 1):
    for stages in (3, 6, 12):
        ms = ctypes.c_float()
        _lib.check(_lib.lib.pmg_debug_load_probe(x.data_ptr(), B, mode, stages, ctypes.byref(ms), torch.cuda.current_stream().cuda_stream), "probe")
        gb = x.numel() * 2 / ms.value / 1e6
        print(json.dumps({"mode": names[mode], "stages": stages, "ms": round(ms.value, 4), "GB/s": round(gb, 1),
                          "B/clk/SM @1.965GHz": round(gb * 1e9 / 148 / 1.965e9, 1)}), flush=True)
