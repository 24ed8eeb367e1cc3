"""Decode the clock64 trace of the fused token-stage kernel (CTA 0).  Run on the GPU box:
    python tools/enc_trace.py [B]"""
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
t = buf.cpu().numpy()
print(f"MMA thread: {int(t[250])} cycles in total waiting for weight blocks (w_full)")
t[250] = 0
t0 = min(v for v in t[:256] if v > 0)
names = {300: "  aux L1h2 scores done", 301: "  aux L1h2 softmax done", 302: "  aux L1h2 PV done", 303: "  aux L1h2 O stored", 0: "mma start", 100: "cmp pdl_wait done", 101: "cmp x0 loaded", 190: "cmp sc_full seen", 191: "cmp logits done"}
for l in range(2):
    o = 40 * l
    names.update({1 + o: f"L{l} mma a_ready(LN1)", 6 + o: f"L{l} mma o_ready", 7 + o: f"L{l} mma outproj issued", 8 + o: f"L{l} mma a_ready(LN2)",
                  20 + o: f"L{l} mma ffn done issued", 102 + o: f"L{l} cmp LN1 arrived", 107 + o: f"L{l} cmp x_full(attn)", 108 + o: f"L{l} cmp LN2 arrived",
                  115 + o: f"L{l} cmp x_full(ffn)"})
    for h in range(4):
        names[2 + o + h] = f"L{l} mma head{h} issued"
        names[103 + o + h] = f"L{l} cmp head{h} drained"
        names[200 + o + 2 * h] = f"L{l} aux head{h} qkv_full"
        names[201 + o + 2 * h] = f"L{l} aux head{h} done"
    for c in range(6):
        names[9 + o + c] = f"L{l} mma F1({c}) issued"
        names[15 + o + c] = f"L{l} mma F2({c}) issued"
        names[109 + o + c] = f"L{l} cmp H({c}) ready"
ev = sorted((int(v - t0), names.get(i, str(i))) for i, v in enumerate(t[:256].tolist() + t[300:304].tolist() + [0]*0) if v > 0) if False else sorted((int(v - t0), names.get(i, str(i))) for i, v in enumerate(t) if v > 0 and (i < 256 or 300 <= i < 304))
prev = 0
for c, n in ev:
    print(f"{c:8d} cyc  (+{c - prev:6d})  {n}")
    prev = c

# ---- front kernel trace (slots 256..): 3rd tile of CTA 0
fn = {36: "mma conv1 last + d1_full commit", 60: "epi d1_full seen", 61: "epi act1 stored"}
for sy in range(4):
    for c in range(3):
        fn[sy * 3 + c] = f"tma sy{sy} c{c} slot free -> issue"
    fn[20 + 4 * sy] = f"mma sy{sy} d0_empty ok"
    fn[21 + 4 * sy] = f"mma sy{sy} a0_full(c2) ok"
    fn[22 + 4 * sy] = f"mma sy{sy} conv0 issued"
    fn[23 + 4 * sy] = f"mma sy{sy} conv1(prev) issued"
    fn[40 + 4 * sy] = f"epi sy{sy} d0_full seen"
    fn[41 + 4 * sy] = f"epi sy{sy} math done"
    fn[42 + 4 * sy] = f"epi sy{sy} a1_free ok"
    fn[43 + 4 * sy] = f"epi sy{sy} a1_ready arrive"
ft = t[256:384]
if (ft > 0).any():
    f0 = min(v for v in ft if v > 0)
    ev = sorted((int(v - f0), fn.get(i, str(i))) for i, v in enumerate(ft) if v > 0)
    prev = 0
    print("---- front kernel, tile #2 of CTA 0")
    for c, n in ev:
        print(f"{c:8d} cyc  (+{c - prev:6d})  {n}")
        prev = c
