"""Decode the clock64 role trace of the PM layer-0 kernel (CTA 0, its 10th tile).  Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pmg_b200
from pmg_b200 import _lib
from oracle import synth
m = pmg_b200.MultiTaskPatchGAN(tasks_dict=synth.PMG_TASKS, input_nc=3, ndf=64, norm="instance", patch_size=73, device=torch.device("cuda"), attn_use_se=False)
m.cuda().eval()
x = torch.randn(128, 3, 224, 224, device="cuda").to(torch.bfloat16)
with torch.no_grad():
    for _ in range(2):
        m(x)
    buf = torch.zeros(512, dtype=torch.int64, device="cuda")
    _lib.lib.pmg_debug_set_trace(buf.data_ptr())
    m(x)
    torch.cuda.synchronize()
    _lib.lib.pmg_debug_set_trace(None)
t = buf.cpu().numpy()
names = {0: "ld  iteration start", 1: "ld  rows landed (wait_group)", 2: "ld  bar.sync", 3: "ld  next rows issued", 4: "ld  a_empty ok",
         5: "ld  tile built", 6: "ld  a_full arrive", 8: "mma loop top", 9: "mma a_full ok", 10: "mma d_empty ok", 11: "mma issued + commits",
         12: "epi loop top", 13: "epi d_full ok", 14: "epi tile done"}
for ps in (1, 2):
    base = 400 + 20 * (ps - 1)
    ev = sorted((int(t[base + k]), n) for k, n in names.items() if t[base + k] > 0)
    if not ev:
        continue
    t0 = ev[0][0]
    print(f"---- pass {ps}")
    prev = 0
    for c, n in ev:
        print(f"{c - t0:8d} cyc (+{c - t0 - prev:6d})  {n}")
        prev = c - t0
