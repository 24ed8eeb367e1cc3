import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from pmg_b200.gram import gram_batched
gen = torch.Generator(device="cuda").manual_seed(0)
for (B, C, HW) in [(3, 128, 16), (5, 64, 16), (3, 80, 200), (2, 96, 3136), (2, 256, 1000), (2, 320, 64), (1, 512, 3136), (150, 128, 64), (2, 512, 16384), (2, 128, 128), (2,128,64), (1,128,72)]:
    F = torch.randn(B, C, HW, device="cuda", generator=gen).to(torch.bfloat16)
    ref = torch.matmul(F.double(), F.double().mT) / HW
    got = gram_batched(F).double()
    err = ((got - ref).norm() / ref.norm()).item()
    # per-tile error map
    bad = ((got - ref).abs() > 1e-3 * ref.abs().max()).nonzero()
    print((B, C, HW), f"rel {err:.2e}", "bad", bad.shape[0], bad[:3].tolist() if bad.shape[0] else "")
