"""Decode the clock64 role trace of the fused conv2 + chan_proj + Gram kernel (tc_midgram.cuh): bands 6..8
of CTA 0 (global bands 6·grid, 7·grid, 8·grid: bands are dealt round-robin).  Run on the GPU box:  python tools/mg_trace.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pmg_b200
from pmg_b200 import _lib
from oracle import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
hp, tasks = synth.PMG_HPARAMS, synth.PMG_TASKS
m = pmg_b200.MultiTaskPatchGANGramModelNonOverlapV2(input_nc=3, ndf=hp["ndf"], norm="batch", patch_size=hp["patch_size"],
        patch_div=hp["patch_div"], num_classes_per_task=tasks, gram_channels=hp["gram_channels"], d_model=hp["d_model"],
        transformer_layers=hp["transformer_layers"], transformer_heads=hp["transformer_heads"])
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.pmg_state_dict(0).items()})
m.cuda().eval()
x = torch.randn(B, 3, 224, 224, device="cuda").to(torch.bfloat16)
with torch.no_grad():
    for _ in range(3):
        m.forward_packed(x)
    buf = torch.zeros(512, dtype=torch.int64, device="cuda")
    _lib.lib.pmg_debug_set_trace(buf.data_ptr())
    m.forward_packed(x)
    torch.cuda.synchronize()
    _lib.lib.pmg_debug_set_trace(None)
t = buf.cpu().numpy()[384:504]
names = {0: "tma slot free -> issue", 1: "mma1 conv2 waits ok", 2: "mma1 conv2 issued", 3: "mma1 chan waits ok", 4: "mma1 chan issued",
         5: "EA d2_full seen", 6: "EA chunk0 math done", 7: "EA a3_free ok", 8: "EA a3_ready arrive", 9: "EB d3_full seen"}
for j in range(3):
    names.update({10 + 3 * j: f"EB p{j} ld done", 11 + 3 * j: f"EB p{j} p_empty ok", 12 + 3 * j: f"EB p{j} p_full arrive",
                  19 + 2 * j: f"mma2 p{j} waits ok", 20 + 2 * j: f"mma2 p{j} issued",
                  25 + 3 * j: f"EC p{j} g_full seen", 26 + 3 * j: f"EC p{j} ld done", 27 + 3 * j: f"EC p{j} stores issued"})
ev = sorted((int(v), f"b{6 + i // 40} " + names.get(i % 40, str(i % 40))) for i, v in enumerate(t) if v > 0)
if not ev:
    print("no trace (B too small for 3 images on CTA 0?)")
else:
    t0 = ev[0][0]
    prev = 0
    for c, n in ev:
        print(f"{c - t0:8d} cyc  (+{c - t0 - prev:6d})  {n}")
        prev = c - t0
