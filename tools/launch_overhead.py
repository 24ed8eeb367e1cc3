"""Host-side cost of one forward (enqueue only) vs its device time, eager launches vs CUDA-graph replay, B = 1024.
    python tools/launch_overhead.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pmg_b200
from oracle import synth

hp, tasks = synth.PMG_HPARAMS, synth.PMG_TASKS
m = pmg_b200.MultiTaskPatchGANGramModelNonOverlapV2(input_nc=3, ndf=hp["ndf"], norm="batch", patch_size=hp["patch_size"],
        patch_div=hp["patch_div"], num_classes_per_task=tasks, gram_channels=hp["gram_channels"], d_model=hp["d_model"],
        transformer_layers=hp["transformer_layers"], transformer_heads=hp["transformer_heads"])
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.pmg_state_dict(0).items()})
m.cuda().eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.randn(B, 3, 224, 224, device="cuda").to(torch.bfloat16)
out = torch.empty((B, 52), device="cuda")
N = 200
with torch.no_grad():
    for _ in range(10):
        m.forward_packed(x, out=out)
    torch.cuda.synchronize()
    # enqueue cost: the GPU is kept busy by a long sleep kernel so the host never blocks on a full queue... simply time the loop
    t0 = time.perf_counter()
    for _ in range(N):
        m.forward_packed(x, out=out)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"eager: host enqueue {1e6 * (t1 - t0) / N:.1f} us/forward, wall incl. drain {1e6 * (t2 - t0) / N:.1f} us/forward")
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        m.forward_packed(x, out=out)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        m.forward_packed(x, out=out)
    for _ in range(10):
        g.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N):
        g.replay()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"graph: host enqueue {1e6 * (t1 - t0) / N:.1f} us/forward, wall incl. drain {1e6 * (t2 - t0) / N:.1f} us/forward")
