#!/usr/bin/env python
"""Measurements of the other BASELINE.json configs (parity cases in tests/, numbers here), one JSON line each:

    python tools/bench_extra.py sweep     # configs[2]: PMG forward, batch sweep 64..1024, 1 B200
    python tools/bench_extra.py pm        # configs[3]: PM forward, batch 1024 (chunks of 256), 1 B200
    python tools/bench_extra.py gram      # configs[1]: Gram microbench, batch 256, C x HW grid
    python tools/bench_extra.py latency   # B=1 latency, eager launches vs CUDA graph
All timed with CUDA events after warm-up, inputs resident in HBM; L2 note per line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pmg_b200
from oracle import synth   # seeded synthetic checkpoint (weights only)

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def ev_time(fn, iters, warm=3):
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


def pmg_model():
    hp, tasks = synth.PMG_HPARAMS, synth.PMG_TASKS
    m = pmg_b200.MultiTaskPatchGANGramModelNonOverlapV2(input_nc=3, ndf=hp["ndf"], norm="batch", patch_size=hp["patch_size"],
            patch_div=hp["patch_div"], num_classes_per_task=tasks, gram_channels=hp["gram_channels"], d_model=hp["d_model"],
            transformer_layers=hp["transformer_layers"], transformer_heads=hp["transformer_heads"])
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.pmg_state_dict(0).items()})
    return m.cuda().eval()


def emit(rows, r, quiet):
    rows.append(r)
    if not quiet:
        print(json.dumps(r), flush=True)


def sweep(quiet=False, iters=40):
    m = pmg_model()
    rows = []
    for B in (64, 128, 256, 512, 1024):
        # rotate over enough distinct batches that the inputs of consecutive iterations do not sit in the 126 MB L2
        nb = max(2, int(np.ceil(260e6 / (B * 3 * 224 * 224 * 2))))
        xs = [torch.randn(B, 3, 224, 224, device="cuda").to(torch.bfloat16) for _ in range(nb)]
        i = [0]
        def step():
            with torch.no_grad():
                m.forward_packed(xs[i[0] % nb]); i[0] += 1
        ms = ev_time(step, iters)
        v = B / ms * 1e3
        emit(rows, {"config": "PMG full forward bf16 batch sweep (BASELINE configs[2])", "batch": B, "ms": round(ms, 4),
                    "images_per_s": round(v, 1), "tensor_frac_of_measured": round(v * 77348864 / 1e12 / PEAK["bf16_tflops"], 4),
                    "hbm_frac_of_measured": round(v * 301264 / 1e9 / PEAK["hbm_gbs"], 4), "l2": f"{nb} rotating input batches"}, quiet)
        del xs
    return rows


def pm(quiet=False, sweep=False):
    """PM forward, B = 1024 in ONE call (pm_forward pipelines 128-image passes on two streams internally).
    sweep=True also times other (images per pass, passes in flight) settings through the tuning hook."""
    from pmg_b200 import _lib
    tasks = synth.PMG_TASKS
    rows = []
    m = pmg_b200.MultiTaskPatchGAN(tasks_dict=tasks, input_nc=3, ndf=64, norm="instance", patch_size=73, device=torch.device("cuda"), attn_use_se=False)
    m.cuda().eval()
    B = 1024
    x = torch.randn(B, 3, 224, 224, device="cuda").to(torch.bfloat16)
    def step():
        with torch.no_grad():
            m(x)
    configs = [(128, 2)] + ([(256, 1), (256, 2), (64, 2), (128, 1)] if sweep else [])
    for chunk, lanes in configs:
        _lib.lib.pm_debug_set_chunking(chunk, lanes)
        ms = ev_time(step, 3, warm=1)
        v = B / ms * 1e3
        emit(rows, {"config": f"PM forward batch 1024 (BASELINE configs[3]; one call, {chunk}-image passes, {lanes} in flight, random-init weights)",
                    "ms": round(ms, 3), "images_per_s": round(v, 1), "flops_per_image": 5.40e9,
                    "tensor_frac_of_measured": round(v * 5.40e9 / 1e12 / PEAK["bf16_tflops"], 4)}, quiet)
    _lib.lib.pm_debug_set_chunking(128, 2)
    if sweep:        # what each round-2 change is worth, one at a time switched off
        for attr, off, what in (("fused_conv0", False, "layer 0 as im2col + GEMM + InstanceNorm kernels"),
                                ("fused_stats", False, "InstanceNorm statistics by a separate pass"),
                                ("igemm_n128", True, "128-wide N tiles in every implicit GEMM")):
            setattr(m, attr, off)
            ms = ev_time(step, 3, warm=1)
            setattr(m, attr, not off)
            emit(rows, {"config": "PM forward batch 1024, default chunking, WITHOUT: " + what, "ms": round(ms, 3), "images_per_s": round(B / ms * 1e3, 1)}, quiet)
    return rows


def gram(quiet=False, Cs=(64, 96, 128, 256, 512), HWs=(16, 25, 196, 3136, 16384)):
    from pmg_b200.gram import gram_batched
    rows = []
    for C in Cs:
        for HW in HWs:
            F = torch.randn(256, C, HW, device="cuda").to(torch.bfloat16)
            # small shapes finish in microseconds: time 10 launches replayed from a CUDA graph so that the Python / ctypes call
            # overhead (~20 us) is not what gets measured
            gram_batched(F)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(10):
                    out = gram_batched(F)
            ms = ev_time(g.replay, 5, warm=2) / 10
            del g, out
            fl, by = 2.0 * C * C * HW * 256, (2.0 * C * HW + 4.0 * C * C) * 256
            bound_ms = max(fl / (PEAK["bf16_tflops"] * 1e12), by / (PEAK["hbm_gbs"] * 1e9)) * 1e3
            emit(rows, {"config": "Gram microbench batch 256 (BASELINE configs[1])", "C": C, "HW": HW, "ms": round(ms, 4),
                        "tflops": round(fl / ms / 1e9, 2), "gbs": round(by / ms / 1e6, 1),
                        "bound": "tensor" if fl / (PEAK["bf16_tflops"] * 1e12) > by / (PEAK["hbm_gbs"] * 1e9) else "hbm",
                        "frac_of_roofline": round(bound_ms / ms, 4)}, quiet)
            del F
    return rows


def latency(quiet=False, batches=(1, 8)):
    from pmg_b200.runner import GraphedPMG
    m = pmg_model()
    rows = []
    for B in batches:
        x = torch.randn(B, 3, 224, 224, device="cuda")     # fp32, as the reference's transform delivers it
        def eager():
            with torch.no_grad():
                m(x)
        g = GraphedPMG(m)
        gh = GraphedPMG(m, host_outputs=True)
        def wall(fn, n=300):
            for _ in range(20):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                fn(); torch.cuda.synchronize()
            return (time.perf_counter() - t0) / n * 1e6
        def ref_style(fn):                                   # the reference's read-back: one .cpu() per task (benchmark_utils.py:73)
            out = fn(x)
            return [torch.softmax(v[0], dim=0).cpu() for v in out.values()]
        r = {"config": "B=%d latency, fp32 input (SURVEY 8f N1)" % B,
             "eager_us": round(wall(eager), 1), "graph_us": round(wall(lambda: g(x)), 1),
             "eager_plus_12_cpu_reads_us": round(wall(lambda: ref_style(m), 100), 1),
             "graph_host_outputs_plus_12_reads_us": round(wall(lambda: ref_style(gh), 100), 1),
             "device_ms_eager": round(ev_time(eager, 200), 4), "device_ms_graph": round(ev_time(lambda: g(x), 200), 4)}
        emit(rows, r, quiet)
    return rows


if __name__ == "__main__":
    {"sweep": sweep, "pm": pm, "pm_sweep": lambda: pm(sweep=True), "gram": gram, "latency": latency}[sys.argv[1]]()
