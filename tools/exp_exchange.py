#!/usr/bin/env python
"""Where does the per-bucket cost of the multi-GPU logit exchange come from?  torchrun, 2+ GPUs.  Variants of the timed
loop (20 steps of 1024 images, buckets of 4): control (no exchange) | event record per bucket | peer gather without its
barriers | without its copies | full peer gather | NCCL gather; plus 40/80 steps to separate the end-of-run drain."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch, torch.distributed as dist
from bench_extra import pmg_model
from pmg_b200.runner import PeerBucketedGather, BucketedGather

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
m = pmg_model()
B = 1024
x = torch.randn(B, 3, 224, 224, device=dev).to(torch.bfloat16)


def run(name, steps, make):
    step, drain = make()
    for _ in range(8):
        step()
    drain(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    drain()
    b.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"variant": name, "steps": steps, "ms_per_step": round(float(t) / steps, 4)}), flush=True)


def control():
    out = torch.empty((B, 52), device=dev)
    def step():
        with torch.no_grad():
            m.forward_packed(x, out=out)
    return step, lambda: None


def events_only():
    out = torch.empty((B, 52), device=dev)
    side = torch.cuda.Stream()
    st = {"n": 0, "ev": None}
    def step():
        with torch.no_grad():
            m.forward_packed(x, out=out)
        st["n"] += 1
        if st["n"] % 4 == 0:
            ev = torch.cuda.Event(); ev.record()
            side.wait_event(ev)
    return step, lambda: None


class NoBarrier(PeerBucketedGather):
    def _launch(self):
        hdl = self._hdl
        class H:
            def barrier(self, channel=0): pass
        self._hdl = H()
        try:
            super()._launch()
        finally:
            self._hdl = hdl


class NoCopies(PeerBucketedGather):
    def _launch(self):
        w = self.world
        self.world = 0
        try:
            super()._launch()
        finally:
            self.world = w


def symm_write_only():
    """logits written into the symmetric (peer-mapped) send area, nothing else: is it the memory the encoder stores to?"""
    g = PeerBucketedGather(B, 52, 4, dev)
    st = {"n": 0}
    def step():
        with torch.no_grad():
            m.forward_packed(x, out=g.send[0][st["n"] % 4])
        st["n"] += 1
    return step, lambda: None


def bucketed(cls):
    def make():
        g = cls(B, 52, 4, dev)
        def step():
            with torch.no_grad():
                m.forward_packed(x, out=g.slot())
            g.commit()
        return step, g.flush
    return make


legs = [("control", 20, control), ("event record + side-stream wait per bucket", 20, events_only),
        ("symmetric-memory logits, no exchange", 20, symm_write_only),
        ("peer gather, no barriers", 20, bucketed(NoBarrier)), ("peer gather, no copies", 20, bucketed(NoCopies)),
        ("peer gather", 20, bucketed(PeerBucketedGather)), ("peer gather", 80, bucketed(PeerBucketedGather)),
        ("nccl all_gather", 20, bucketed(BucketedGather)), ("nccl all_gather", 80, bucketed(BucketedGather)),
        ("control", 80, control)]
if "--quick" in sys.argv:          # the 20-step legs only, three rounds (leg-to-leg noise on a multi-GPU box is +-3 %)
    legs = [l for l in legs if l[1] == 20 and "nccl" not in l[0]] * 3
for name, steps, mk in legs:
    run(name, steps, mk)
dist.destroy_process_group()
