import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pmg_b200
from oracle import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
hp, tasks = synth.PMG_HPARAMS, synth.PMG_TASKS
m = pmg_b200.MultiTaskPatchGANGramModelNonOverlapV2(input_nc=3, ndf=hp["ndf"], norm="batch", patch_size=hp["patch_size"],
        patch_div=hp["patch_div"], num_classes_per_task=tasks, gram_channels=hp["gram_channels"], d_model=hp["d_model"],
        transformer_layers=hp["transformer_layers"], transformer_heads=hp["transformer_heads"])
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.pmg_state_dict(0).items()})
m.cuda().eval()
x = torch.randn(B, 3, 224, 224, device="cuda").to(torch.bfloat16)
with torch.no_grad():
    y = m.forward_packed(x)
torch.cuda.synchronize()
print("ok", y.abs().mean().item())
