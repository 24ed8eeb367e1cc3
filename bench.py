#!/usr/bin/env python
"""bench.py — PMG images/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
    python bench.py --impl reference [--gpus N] --steps K --warmup W  # CPU baseline arm (host cores)

For N>1 launch with torchrun (one rank per GPU); rank 0 prints the line.  Workload (config.workload):
BASELINE.json configs[2] at its largest single-GPU batch — PMG eval forward, committed hparams,
1024 images / GPU, 3×224×224 bf16, synthetic images + seeded synthetic checkpoint.  Weak scaling: the
per-GPU batch is fixed, the only collective is an all-gather of the packed logits (one per --gather-steps steps).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_IMAGE = 77_348_864
CONV_GRAM_FLOPS = 34_654_208            # conv stack + chan_proj + Gram
BYTES_PER_IMAGE = 3 * 224 * 224 * 2 + 52 * 4   # bf16 image in + fp32 logits out
# per image, committed config: (algorithmic FLOPs, algorithmic HBM bytes) of each stage as it is split into kernels
# (SURVEY.md §8a for the FLOPs; bytes = the stage's compulsory reads + writes)
STAGES = {
    "front_conv0_conv1": (9_633_792 + 12_845_056, 3 * 224 * 224 * 2 + 2 * 196 * 64 * 2),   # bf16 image in, act1 hi/lo out
    "im2col0": (0, 3 * 224 * 224 * 2 + 3136 * 64 * 2),
    "conv0": (9_633_792, 3136 * 64 * 2 + 2 * 196 * 512 * 2),
    "conv1": (12_845_056, 2 * 196 * 512 * 2 + 2 * 196 * 64 * 2),
    "conv2_chan_proj": (3_211_264 + 4_816_896, 2 * 196 * 64 * 2 + 196 * 96 * 4),
    "conv2": (3_211_264, 2 * 196 * 64 * 2 + 2 * 196 * 128 * 2),
    "chan_proj": (4_816_896, 2 * 196 * 128 * 2 + 196 * 96 * 4),
    # conv2 + chan_proj + Gram in one kernel: act1 hi/lo in, 6 upper 32x32 blocks of G hi/lo out (chan never leaves the SM)
    "conv2_chan_gram": (3_211_264 + 4_816_896 + 4_147_200, 2 * 196 * 64 * 2 + 2 * 9 * 6144 * 2),
    "gram": (4_147_200, 196 * 96 * 4 + 2 * 9 * 6144 * 2),            # chan in, 6 upper 32x32 blocks of G hi/lo out
    "token_proj": (31_850_496, 2 * 9 * 6144 * 2 + 4 * 9 * 192 * 4),  # G' hi/lo in, split-K partials out (two launches)
    "tok_reduce_ln": (0, 3 * 9 * 192 * 4 + 2 * 9 * 192 * 2),
    "encoder_layer": (5_370_624, 9 * 192 * 4 * 2),
    "pool_heads": (102_912, 9 * 192 * 4 + 52 * 4),
    "encoder_heads": (10_741_248 + 102_912, 2 * 9 * 192 * 4 + 52 * 4),
}


def bind_to_gpu_numa(index):
    """Multi-GPU runs: pin this rank's CPU threads to the cores NVML reports as local to its GPU, so the pinned host
    buffers of the end-to-end leg are first-touched on the NUMA node behind that GPU's PCIe root (8 ranks streaming from
    one node's DRAM halve each other's H2D rate).  Best effort: silently skipped when NVML / affinity are unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json, burst)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [s.strip() for s in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port(batch, seed=0):
    import torch
    from oracle import synth
    from oracle.pmg_torch_port import PMGTorchPort
    torch.set_num_threads(os.cpu_count() or 1)
    port = PMGTorchPort(synth.pmg_state_dict(0), synth.PMG_TASKS, 3, 4)
    x = torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(seed))
    return port, x


def time_cpu_baseline(budget_s=12.0, batch=64):
    """Reference-equivalent CPU forward (oracle/pmg_torch_port.py) on the host cores: bounded sample."""
    port, x = cpu_port(batch)
    for _ in range(2):
        port(x)
    n, t0 = 0, time.perf_counter()
    while True:
        port(x); n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= 5000:
            break
    return {"value": round(n * batch / el, 1), "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{n} forwards of batch {batch} (3x224x224 fp32) in {el:.1f}s, torch CPU ATen ops = the "
                      f"reference's op sequence (oracle/pmg_torch_port.py), {os.cpu_count()} threads"}


def extra_block(model, x):
    """N = 1 only: the other BASELINE.json configs and the library comparator, measured in the same run (each a few
    hundred ms of GPU time): configs[1] Gram microbench rows, configs[2] batch sweep, configs[3] PM forward B = 1024,
    the B = 1 latency path, device-resident fp32 input, and the reference's op sequence in eager PyTorch ON
    THE GPU (cuDNN / cuBLAS, default flags) — the library baseline BASELINE.md section 4 item 6 names."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_extra as bx
    from oracle import synth
    from oracle.pmg_torch_port import PMGTorchPort
    out = {}
    keep = lambda rows, keys: [{k: r[k] for k in keys if k in r} for r in rows]
    try:
        out["batch_sweep_bf16"] = keep(bx.sweep(quiet=True, iters=20), ("batch", "ms", "images_per_s"))
        out["pm_forward_b1024"] = keep(bx.pm(quiet=True), ("ms", "images_per_s", "tensor_frac_of_measured"))[0]
        out["gram_microbench_b256"] = keep(bx.gram(quiet=True, Cs=(96, 128, 512), HWs=(25, 196, 3136, 16384)),
                                           ("C", "HW", "ms", "tflops", "gbs", "bound", "frac_of_roofline"))
        out["latency"] = keep(bx.latency(quiet=True, batches=(1,)),
                              ("config", "eager_us", "graph_us", "eager_plus_12_cpu_reads_us", "graph_host_outputs_plus_12_reads_us"))[0]
        xf = x.float()
        ms = bx.ev_time(lambda: model.forward_packed(xf), 10)
        out["fp32_device_input_b1024"] = {"ms": round(ms, 4), "images_per_s": round(x.shape[0] / ms * 1e3, 1),
                                          "launches": model.launches_per_forward(x.shape[0], dtype=torch.float32)}
        port = PMGTorchPort(synth.pmg_state_dict(0), synth.PMG_TASKS, 3, 4)
        port.sd = {k: v.cuda() for k, v in port.sd.items()}
        port.encoder = port.encoder.cuda()
        ms = bx.ev_time(lambda: port(xf), 5, warm=2)
        out["gpu_eager_pytorch_baseline"] = {
            "ms": round(ms, 3), "images_per_s": round(x.shape[0] / ms * 1e3, 1), "dtype": "fp32 (default flags)",
            "what": "the reference's ATen op sequence (oracle/pmg_torch_port.py) in eager PyTorch on this B200, B = 1024, "
                    "cuDNN/cuBLAS kernels — a library comparator, not the reference arm"}
        del xf
    except Exception as e:      # the headline line must still print
        out["error"] = f"{type(e).__name__}: {e}"
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    batch = 256
    port, x = cpu_port(batch)
    for _ in range(max(1, args.warmup)):
        port(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        port(x)
    el = time.perf_counter() - t0
    v = args.steps * batch / el
    line = {
        "impl": "reference", "metric": "pmg_images_per_sec", "value": round(v, 1), "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * el / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_per_step": f"{batch} images (bounded CPU sample of the 1024/GPU batch)"},
        "cpu_baseline": {"value": round(v, 1), "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{args.steps} steps x {batch} images, torch CPU, {os.cpu_count()} threads"},
        "e2e": {"value": round(v, 1), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


WORKLOAD = ("PMG eval forward, committed hparams (ndf32 patch_div3 gram96 d192 L2 H4, 12 tasks/52 classes), "
            "3x224x224 bf16, 1024 images per GPU, synthetic images + seeded synthetic checkpoint")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="images per GPU per step")
    ap.add_argument("--e2e-chunk", type=int, default=512, help="largest H2D piece of HostPipeline (the last one is split 256, 128, ... 32)")
    ap.add_argument("--gather-steps", type=int, default=4, help="N>1: steps whose logits share one all_gather")
    ap.add_argument("--nccl-gather", action="store_true", help="N>1: use the NCCL all-gather instead of peer memory")
    ap.add_argument("--exchange", default="pull", choices=["push", "pull"],
                    help="N>1, peer memory: copy-engine pulls (pull, default: measured faster) or logits stored to every rank by the "
                         "token-stage kernel itself (push, pmg_forward_bcast)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--force-generic", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-memory end-to-end leg (ncu runs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the `extra` block (other BASELINE configs, N=1 only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    import pmg_b200
    from pmg_b200 import _lib
    from pmg_b200.runner import BucketedGather, HostPipeline, make_bucketed_gather
    from oracle import synth   # seeded synthetic checkpoint generator (test infra; weights only, not compute)

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path); use --impl reference for the CPU arm"
    W = max(3, args.warmup)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hp, tasks = synth.PMG_HPARAMS, synth.PMG_TASKS
    model = pmg_b200.MultiTaskPatchGANGramModelNonOverlapV2(
        input_nc=3, ndf=hp["ndf"], norm="batch", patch_size=hp["patch_size"], patch_div=hp["patch_div"],
        num_classes_per_task=tasks, gram_channels=hp["gram_channels"], d_model=hp["d_model"],
        transformer_layers=hp["transformer_layers"], transformer_heads=hp["transformer_heads"])
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.pmg_state_dict(0).items()})
    model.to(dev).eval()
    model.force_generic = args.force_generic
    B = args.batch
    gen = torch.Generator(device=dev).manual_seed(rank)
    x = torch.randn(B, 3, 224, 224, device=dev, generator=gen).to(torch.bfloat16)
    GB = B * world

    # the logits of --gather-steps consecutive steps go out in ONE all_gather (runner.BucketedGather): each forward writes
    # straight into its slot of the bucket, the gather of a full bucket runs under the forwards that fill the next one, and
    # drain() gathers what is left and waits for everything, inside the timed region
    bucket, exchange = (make_bucketed_gather(B, 52, args.gather_steps, dev, prefer_peer=not args.nccl_gather, kind=args.exchange)
                        if world > 1 else (None, "none"))

    lg_single = torch.empty((B, 52), dtype=torch.float32, device=dev)

    def step():
        with torch.no_grad():
            if bucket is None:
                return model.forward_packed(x, out=lg_single)
            lg = model.forward_packed(x, out=bucket.slot(), also=bucket.peers())
            bucket.commit()
        return lg

    def drain():
        if bucket is not None:
            bucket.flush()

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_enqueue_ms = []

    def timed(step_fn, drain_fn):
        """W warm-up + args.steps timed steps of step_fn on this rank; returns this rank's device ms (events on the
        launching stream, barrier + synchronize on both sides)."""
        for _ in range(W):
            step_fn()
        drain_fn()
        fence()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        h0 = time.perf_counter()
        for _ in range(args.steps):
            step_fn()
        host_enqueue_ms.append((time.perf_counter() - h0) * 1e3 / args.steps)     # host time to ENQUEUE a step (not in the JSON value)
        drain_fn()
        b.record()
        fence()
        return a.elapsed_time(b)

    def over_ranks(ms_local):
        """(max, [per-rank ms]) — the job's time is the slowest rank's."""
        if world == 1:
            return ms_local, [ms_local]
        t = torch.zeros(world, device=dev)
        dist.all_gather_into_tensor(t, torch.tensor([ms_local], device=dev))
        per = [float(v) for v in t.tolist()]
        return max(per), per

    def step_nocoll():
        with torch.no_grad():
            model.forward_packed(x, out=lg_single)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # N > 1: the control leg (same forwards, no exchange) runs FIRST.  On a multi-GPU box whatever is measured in the first
    # second of a process comes out a few per cent slow (tools/exp_exchange.py --quick: control 0.294 -> 0.283 -> 0.279 ms over
    # three rounds); the diagnostic leg takes that, the recorded leg is measured after it.  (Order only: every leg is still W
    # warm-up + exactly K timed steps between barriers.)
    if world > 1:
        ms_nc, rank_ms_nc = over_ranks(timed(step_nocoll, lambda: None))
    ms, rank_ms = over_ranks(timed(step, drain))
    value = GB * args.steps / (ms / 1e3)

    # ---- N > 1: (a) control leg without the collective (same forwards, logits stay on their rank) so that the cost of the
    # exchange and the rank-to-rank spread are separately visible; (b) correctness of the exchange on the real GPUs: one
    # gathered bucket must hold, at every rank's slot, exactly the logits that rank computed.
    multi = None
    if world > 1:
        chk, _ = make_bucketed_gather(B, 52, args.gather_steps, dev, keep=True, prefer_peer=not args.nccl_gather, kind=args.exchange)
        with torch.no_grad():
            for _ in range(3 * args.gather_steps + 1):      # several buckets and a partly filled last one: every index is reused
                model.forward_packed(x, out=chk.slot(), also=chk.peers())
                chk.commit()
            chk.flush()
            mine = model.forward_packed(x)
        got = chk.pop()
        sums = torch.zeros(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(sums, mine.double().sum().reshape(1))
        ok = len(got) == 3 * args.gather_steps + 1
        for t in got:                                   # t: (world, B, 52) — slot r = rank r's logits of that micro-batch
            ok = ok and bool(torch.equal(t[rank], mine))
            ok = ok and bool(torch.equal(t.double().sum(dim=(1, 2)), sums))
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        assert int(flag.item()) == 1, "gathered logits differ from the per-rank logits"
        # what the end-of-window drain of the LAST bucket costs (nothing is left to overlap it with): events around flush()
        dr = []
        for _ in range(5):
            with torch.no_grad():
                for sidx in range(args.gather_steps):
                    model.forward_packed(x, out=bucket.slot(), also=bucket.peers())
                    if sidx < args.gather_steps - 1:
                        bucket.commit()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record()
            bucket.commit()
            bucket.flush()
            d1.record()
            fence()
            dr.append(d0.elapsed_time(d1))
        drain_ms = over_ranks(statistics.median(dr))[0]
        stat = lambda v: {"min": round(min(v) / args.steps, 4), "median": round(statistics.median(v) / args.steps, 4),
                          "max": round(max(v) / args.steps, 4)}
        multi = {"rank_ms_per_step": stat(rank_ms),
                 "host_enqueue_ms_per_step": {"exchange_leg": round(host_enqueue_ms[1], 4), "no_collective_leg": round(host_enqueue_ms[0], 4),
                                              "what": "rank 0's host time to enqueue one step; the device time per step is above"},
                 "no_collective": {"value": round(GB * args.steps / (ms_nc / 1e3), 1), "rank_ms_per_step": stat(rank_ms_nc)},
                 "last_bucket_drain_ms": round(drain_ms, 4),
                 "gather_check": "every rank: gathered (world,B,52) bucket == own logits at its slot, peers' slots match "
                                 "their fp64 checksums"}

    # ---- end to end through the public API: pinned host batch → H2D → (GPU ToTensor+Normalize) → forward → D2H logits,
    # every step.  Three host input formats: decoded uint8 HWC images (headline: 150 KB/image over PCIe), the
    # normalised bf16 batch, and the fp32 batch the reference's own DataLoader delivers.
    e2e_ms, h2d_ms, streamed_ms = {}, {}, None
    if not args.no_e2e:
        gcpu = torch.Generator().manual_seed(rank)
        hosts = {
            "uint8_nhwc": torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, generator=gcpu).pin_memory(),
            "bf16_nchw": x.cpu().pin_memory(),
            "fp32_nchw": x.float().cpu().pin_memory(),
        }
        pipe = HostPipeline(model, chunk=args.e2e_chunk, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def leg(fn):
            for _ in range(3):
                fn()
            fence()
            e0.record()
            for _ in range(args.steps):
                fn()
            e1.record()
            fence()
            return over_ranks(e0.elapsed_time(e1))[0]

        for name, x_host in hosts.items():
            # logits land in pinned host memory per rank: the D2H read is the result delivery
            e2e_ms[name] = leg(lambda: pipe.infer(x_host))
            if name == "uint8_nhwc":
                # streaming form of the same API: step i+1 is submitted before step i's result is read, so the copy
                # engine never waits for the forward of a step's last chunk.  Every step still copies its inputs from
                # pinned host memory and its logits are read on the host, all inside the timed region.
                def streamed_leg():
                    for _ in range(3):
                        pipe.infer(x_host)
                    fence()
                    e0.record()
                    prev = pipe.submit(x_host)
                    for _ in range(args.steps - 1):
                        nxt = pipe.submit(x_host)
                        pipe.collect(prev)
                        prev = nxt
                    pipe.collect(prev)
                    e1.record()
                    fence()
                    return over_ranks(e0.elapsed_time(e1))[0]
                streamed_ms = streamed_leg()
            # the wire itself: a bare pinned -> device copy of the same batch on every rank at once, no compute — the
            # ceiling any end-to-end number from this host format can reach on this box
            dbuf = torch.empty(x_host.shape, dtype=x_host.dtype, device=dev)
            h2d_ms[name] = leg(lambda: dbuf.copy_(x_host, non_blocking=True))
            del dbuf
        del hosts
    clocks = sampler.stop() if sampler else None
    E2E_BYTES = {"uint8_nhwc": 3 * 224 * 224, "bf16_nchw": 3 * 224 * 224 * 2, "fp32_nchw": 3 * 224 * 224 * 4}

    def e2e_entry(name):
        ms_leg = e2e_ms[name]
        wire = GB * args.steps / (h2d_ms[name] / 1e3)
        return {"value": round(GB * args.steps / (ms_leg / 1e3), 1), "unit": "images/s", "h2d_bytes_per_step": GB * E2E_BYTES[name],
                "d2h_bytes_per_step": GB * 52 * 4, "ms_per_step": round(ms_leg / args.steps, 4), "host_input": name,
                "bytes_are": "whole job (all ranks) per step",
                "h2d_ceiling": {"images_per_s": round(wire, 1), "gb_per_s": round(wire * E2E_BYTES[name] / 1e9, 2),
                                "what": "bare pinned->device copy of the same batch, all ranks at once, timed in this run"},
                "frac_of_h2d_ceiling": round(h2d_ms[name] / ms_leg, 4),
                "reference_contract_format": name == "fp32_nchw"}

    def e2e_block():
        """Headline = the streaming form of the public API (HostPipeline.submit / collect, two batches outstanding): every step
        copies ITS inputs from pinned host memory and ITS logits are read on the host inside the timed region; only the wait
        for step i's last small forward no longer idles the copy engine.  `blocking` = one HostPipeline.infer() per step
        (host sync after every step), the shape of the reference's own loop."""
        blk = e2e_entry("uint8_nhwc")
        out = dict(blk)
        if streamed_ms is not None:
            out.update({"value": round(GB * args.steps / (streamed_ms / 1e3), 1), "ms_per_step": round(streamed_ms / args.steps, 4),
                        "frac_of_h2d_ceiling": round(h2d_ms["uint8_nhwc"] / streamed_ms, 4)})
        out.update(
            api="pmg_b200.runner.HostPipeline.submit(pinned host uint8 HWC images) / collect(ticket) -> GPU ToTensor+Normalize "
                "(prepost.preprocess_u8) -> forward -> pinned host logits; two batches outstanding",
            blocking={"value": blk["value"], "unit": "images/s", "ms_per_step": blk["ms_per_step"],
                      "frac_of_h2d_ceiling": blk["frac_of_h2d_ceiling"],
                      "api": "HostPipeline.infer(x): same pipeline, host blocks on every step's logits before the next step's "
                             "first copy is issued"},
            chunk=args.e2e_chunk,
            format_note="headline leg = decoded uint8 HWC images (what a DataLoader holds before ToTensor/Normalize, "
                        "utils/datasets_utils.py:212-214, done on the GPU here); the reference module's own forward(x) "
                        "contract is the normalised fp32 NCHW batch = other_host_inputs.fp32_nchw (blocking infer())",
            host_numa_binding=(f"rank threads bound to the {numa_cpus} cores NVML reports local to the GPU" if numa_cpus else "none"),
            other_host_inputs={k: e2e_entry(k) for k in ("bf16_nchw", "fp32_nchw")})
        return out

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-stage device times (events between kernels, on the launching stream) → roofline
    import ctypes
    pk = peaks()
    cfgp = ctypes.byref(model._cfg)
    stages = (_lib.PmgStageTime * 64)()
    n_st = ctypes.c_int(0)
    acc = {}
    reps = 5
    lg = torch.empty((B, 52), dtype=torch.float32, device=dev)
    ws_prof = torch.empty(model.workspace_bytes(B), dtype=torch.uint8, device=dev)
    with torch.no_grad():
        model.forward_packed(x, out=lg, ws=ws_prof)      # makes sure the packed weights exist
    flags = _lib.PMG_FLAG_FORCE_GENERIC if args.force_generic else 0
    for _ in range(reps):
        rc = _lib.lib.pmg_forward_profile(cfgp, model._packed.data_ptr(), x.data_ptr(), _lib.PMG_BF16, B, 224, 224,
                                          lg.data_ptr(), ws_prof.data_ptr(), ws_prof.numel(), flags,
                                          torch.cuda.current_stream().cuda_stream, stages, 64, ctypes.byref(n_st))
        _lib.check(rc, "pmg_forward_profile")
        for i in range(n_st.value):
            nm = stages[i].name.decode()
            acc.setdefault(nm, []).append(stages[i].ms)
    # a stage name can repeat inside one forward (encoder_layer ×2): average per launch
    per_launch = {k: sum(v) / len(v) for k, v in acc.items()}
    per_forward = {k: sum(v) / reps for k, v in acc.items()}
    dom = max(per_forward, key=per_forward.get)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(dom)
    dom_flops, dom_bytes = STAGES.get(dom, (0, 0))
    ridge = pk["bf16_tflops"] * 1e12 / (pk["hbm_gbs"] * 1e9)
    sec = per_launch[dom] / 1e3
    if dom_bytes and dom_flops / dom_bytes < ridge:       # arithmetic intensity below the ridge: HBM-bound kernel
        ach = dom_bytes * B / sec / 1e9
        roofline = {"kernel": dom, "bound": "hbm", "achieved": round(ach, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": round(ach / pk["hbm_gbs"], 4)}
    else:
        ach = dom_flops * B / sec / 1e12
        roofline = {"kernel": dom, "bound": "tensor", "achieved": round(ach, 2), "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": round(ach / pk["bf16_tflops"], 4)}
    roofline.update({
        "traffic": traffic, "peak_source": pk["source"],
        "algorithmic_per_image": {"flops": dom_flops, "bytes": dom_bytes},
        "ms_per_launch": round(per_launch[dom], 4), "share_of_forward": round(per_forward[dom] / sum(per_forward.values()), 3),
        "stage_ms": {k: round(v, 4) for k, v in per_forward.items()},
        "stage_frac_of_own_roofline": {
            k: round(max(STAGES[k][0] * B / (per_launch[k] / 1e3) / 1e12 / pk["bf16_tflops"],
                         STAGES[k][1] * B / (per_launch[k] / 1e3) / 1e9 / pk["hbm_gbs"]), 3)
            for k in per_launch if k in STAGES},
        "forward": {
            "tensor_frac": round(value / world * FLOPS_PER_IMAGE / 1e12 / pk["bf16_tflops"], 4),
            "conv_gram_tensor_frac": round(value / world * CONV_GRAM_FLOPS / 1e12 / pk["bf16_tflops"], 4),
            "hbm_frac": round(value / world * BYTES_PER_IMAGE / 1e9 / pk["hbm_gbs"], 4),
            "flops_per_image": FLOPS_PER_IMAGE, "bytes_per_image": BYTES_PER_IMAGE,
        },
    })
    launches = model.launches_per_forward(B) * args.steps
    line = {
        "metric": "pmg_images_per_sec", "value": round(value, 1), "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": GB, "per_gpu_batch": B, "parallelism": f"dp{world}",
                   "l2_policy": f"input {B * 3 * 224 * 224 * 2 / 1e6:.0f} MB per GPU per step > 126 MB L2",
                   "path": "generic-fp32" if args.force_generic else "default",
                   "collective": (f"{exchange}: ({args.gather_steps},B_local,52) fp32 logits of every rank on every rank, once per "
                                  f"{args.gather_steps} steps, async under the next forwards; all waited for inside the timed region")
                                 if world > 1 else "none"},
        "clocks": clocks,
        "e2e": None if not e2e_ms else e2e_block(),
        "gpu_launches": launches,
        "roofline": roofline,
    }
    if multi:
        line["multi_gpu"] = multi
    if world == 1 and not args.no_extra:
        line["extra"] = extra_block(model, x)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = time_cpu_baseline()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
