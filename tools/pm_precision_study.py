"""PM precision study (CPU, reference module): what happens to logits / attention when the inter-layer activations
(after InstanceNorm + LeakyReLU) are rounded to ONE bf16 instead of the hi/lo pair the CUDA path carries."""
import sys, numpy as np, torch
sys.path.insert(0, "/root/reference"); sys.path.insert(0, "/root/repo")
from Models.models_PM import MultiTaskPatchGAN
from oracle import synth
torch.manual_seed(0)
tasks = synth.PMG_TASKS
def build():
    m = MultiTaskPatchGAN(tasks_dict=tasks, input_nc=3, ndf=64, norm="instance", patch_size=73, device="cpu", attn_tau=0.7,
                          attn_use_se=False, attn_softmax_spatial=True)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.pm_state_dict(0).items()})
    return m.eval()
x = torch.from_numpy(synth.images(3, 8))
ref = build().double()
with torch.no_grad():
    want = ref(x.double())
for mode in ("all", "not_last", "input_only"):
    m = build()
    acts = [mod for mod in m.trunk.modules() if isinstance(mod, torch.nn.LeakyReLU)]
    hooks = []
    sel = acts if mode == "all" else acts[:-1] if mode == "not_last" else []
    for a in sel:
        hooks.append(a.register_forward_hook(lambda mod, i, o: o.to(torch.bfloat16).float()))
    with torch.no_grad():
        got = m(x)
    worst = 0.0; flips = 0
    for t in tasks:
        a, b = got[t].double(), want[t]
        worst = max(worst, float((a - b).norm() / b.norm()))
        flips += int((a.argmax(1) != b.argmax(1)).sum())
    allg = torch.cat([got[t].double() for t in tasks], 1); allw = torch.cat([want[t] for t in tasks], 1)
    print(mode, "n_rounded", len(sel), "logits rel", float((allg - allw).norm() / allw.norm()), "worst task", worst, "argmax flips", flips)
