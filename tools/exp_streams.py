#!/usr/bin/env python
"""Experiments (one JSON line each), B200:
  streams : consecutive 1024-image steps issued round-robin on S CUDA streams, each with its own workspace and logits
            buffer (forward_packed(x, out=, ws=)); does the tail of one step's persistent kernels fill with the next step's?
  fp32    : device-resident fp32 NCHW input at B = 1024 vs bf16"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench_extra import pmg_model


def timed(fn, iters=40, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def streams():
    m = pmg_model()
    B = 1024
    xs = [torch.randn(B, 3, 224, 224, device="cuda").to(torch.bfloat16) for _ in range(2)]
    with torch.no_grad():
        m.forward_packed(xs[0])
    for S in (1, 2, 3):
        sts = [torch.cuda.Stream() for _ in range(S)]
        wss = [torch.empty(m.workspace_bytes(B), dtype=torch.uint8, device="cuda") for _ in range(S)]
        outs = [torch.empty((B, 52), device="cuda") for _ in range(S)]
        i = [0]

        def step():
            k = i[0] % S
            i[0] += 1
            with torch.cuda.stream(sts[k]), torch.no_grad():
                m.forward_packed(xs[i[0] % 2], out=outs[k], ws=wss[k])

        def run():
            cur = torch.cuda.current_stream()
            for s in sts:
                s.wait_stream(cur)
            for _ in range(S * 4):
                step()
            for s in sts:
                cur.wait_stream(s)
        ms = timed(run, iters=10) / (S * 4)
        print(json.dumps({"exp": "streams", "streams": S, "ms_per_step": round(ms, 4), "images_per_s": round(B / ms * 1e3, 1)}), flush=True)


def fp32():
    m = pmg_model()
    B = 1024
    xb = [torch.randn(B, 3, 224, 224, device="cuda").to(torch.bfloat16) for _ in range(2)]
    xf = [x.float() for x in xb]
    i = [0]
    for name, xs in (("bf16", xb), ("fp32", xf), ("fp32 im2col path (fused_front=False)", xf)):
        m.fused_front = "im2col" not in name
        def step():
            i[0] += 1
            with torch.no_grad():
                m.forward_packed(xs[i[0] % 2])
        ms = timed(step)
        print(json.dumps({"exp": "input dtype", "dtype": name, "launches": m.launches_per_forward(B, dtype=xs[0].dtype), "ms_per_step": round(ms, 4),
                          "images_per_s": round(B / ms * 1e3, 1)}), flush=True)


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    {"streams": streams, "fp32": fp32}[sys.argv[1]]()
