"""Host logic of the batch-sharded runner on CPU: world_size-2 gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_bounds_cover_batch():
    from pmg_b200.runner import shard_bounds
    for n in (1, 7, 8, 8192, 8193):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pmg_b200.runner import gather_logits, shard_bounds
    full = torch.arange(n * 52, dtype=torch.float32).reshape(n, 52)
    lo, hi = shard_bounds(n, world, rank)
    got = gather_logits(full[lo:hi].clone(), n)
    q.put((rank, bool(torch.equal(got, full))))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 9])
def test_gather_logits_world2_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _worker_async(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pmg_b200.runner import gather_logits_async, shard_bounds
    full = torch.arange(n * 52, dtype=torch.float32).reshape(n, 52)
    lo, hi = shard_bounds(n, world, rank)
    h1 = gather_logits_async(full[lo:hi].clone(), n)
    h2 = gather_logits_async(full[lo:hi].clone() * 2, n)      # two gathers in flight
    ok = bool(torch.equal(h1.result(), full)) and bool(torch.equal(h2.result(), full * 2))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gather_logits_async_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_async, args=(r, 2, port, 9, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _worker_bucket(rank, world, port, n_steps, steps_per_gather, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pmg_b200.runner import BucketedGather
    b, L = 5, 52
    bg = BucketedGather(b, L, steps_per_gather, torch.device("cpu"), keep=True)
    got = []
    for s in range(n_steps):
        # stand-in for model.forward_packed(x, out=bg.slot()): micro-batch s of rank r holds 1000 s + 100 r + row
        bg.slot().copy_((1000 * s + 100 * rank + torch.arange(b, dtype=torch.float32))[:, None].expand(b, L))
        bg.commit()
        got.extend(t.reshape(-1, L).clone() for t in bg.pop())     # views live two buckets: clone as a consumer would
    bg.flush()
    got.extend(t.reshape(-1, L).clone() for t in bg.pop())
    ok = len(got) == n_steps
    for s, t in enumerate(got):
        want = torch.cat([1000 * s + 100 * r + torch.arange(b, dtype=torch.float32) for r in range(world)])[:, None].expand(world * b, L)
        ok = ok and bool(torch.equal(t, want))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_steps,steps_per_gather", [(8, 4), (7, 3), (2, 4), (9, 1)])
def test_bucketed_gather_world2_gloo(n_steps, steps_per_gather):
    """Several micro-batches per collective, two buckets in flight, partial last bucket (runner.BucketedGather)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_bucket, args=(r, 2, port, n_steps, steps_per_gather, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
