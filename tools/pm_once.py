import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pmg_b200
from oracle import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = pmg_b200.MultiTaskPatchGAN(tasks_dict=synth.PMG_TASKS, input_nc=3, ndf=64, norm="instance", patch_size=73, device=torch.device("cuda"), attn_use_se=False)
m.cuda().eval()
x = torch.randn(B, 3, 224, 224, device="cuda").to(torch.bfloat16)
with torch.no_grad():
    for _ in range(2):
        m(x)
torch.cuda.synchronize()
print("ok")
