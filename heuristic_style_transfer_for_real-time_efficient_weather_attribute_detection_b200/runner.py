"""Batch-sharded data-parallel PMG inference: one process per GPU, a contiguous slice of the batch per
rank, and a single collective — an all-gather of the packed (B_local, Σncls) fp32 logits — at the end
(SURVEY.md §8e).  The reference has no multi-GPU path; images are independent (eval BatchNorm uses
running stats, models_PMG.py:176-199 has no cross-sample op), so there is no other exchange step.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n items for `rank`; the first n % world ranks get one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class PendingGather:
    """Handle of an in-flight logits all-gather: ``result()`` waits for the collective and returns the
    (global_batch, L) tensor.  Lets the next micro-batch's forward run under the gather."""

    def __init__(self, work, out, sizes, mx):
        self.work, self.out, self.sizes, self.mx = work, out, sizes, mx

    def result(self) -> torch.Tensor:
        if self.work is not None:
            self.work.wait()
            self.work = None
        if all(hi - lo == self.mx for lo, hi in self.sizes):
            return self.out
        return torch.cat([self.out[r * self.mx: r * self.mx + (hi - lo)] for r, (lo, hi) in enumerate(self.sizes)], dim=0)


def gather_logits_async(local: torch.Tensor, global_batch: int, group=None) -> PendingGather:
    """Same single collective as gather_logits, issued with async_op=True."""
    world = dist.get_world_size(group)
    L = local.shape[1]
    sizes = [shard_bounds(global_batch, world, r) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    send = local
    if local.shape[0] != mx:
        send = local.new_zeros((mx, L))
        send[: local.shape[0]] = local
    out = local.new_empty((world * mx, L))
    work = dist.all_gather_into_tensor(out, send.contiguous(), group=group, async_op=True)
    return PendingGather(work, out, sizes, mx)


def gather_logits(local: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
    """All-gather ragged (B_local, L) shards into (global_batch, L) on every rank with ONE collective:
    shards are padded to the largest shard so a single all_gather_into_tensor moves them."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    L = local.shape[1]
    sizes = [shard_bounds(global_batch, world, r) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    send = local
    if local.shape[0] != mx:
        send = local.new_zeros((mx, L))
        send[: local.shape[0]] = local
    out = local.new_empty((world * mx, L))
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    if all(hi - lo == mx for lo, hi in sizes):
        return out
    return torch.cat([out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


class BucketedGather:
    """Gathers the logits of several consecutive micro-batches with ONE collective.

    A (B_local, L) logits shard is 213 KB at 1024 images: over NVLink 5 the all-gather is pure launch latency, and a
    collective kernel that runs beside the next forward also takes SMs away from the persistent 148-CTA kernels.  So the
    forwards of ``steps`` micro-batches write their logits straight into consecutive slots of one bucket
    (``slot()`` → ``forward_packed(x, out=slot)`` → ``commit()``); when the bucket is full ``commit()`` starts a single
    asynchronous all_gather_into_tensor of the whole bucket and a second bucket takes the next micro-batches meanwhile.
    ``flush()`` gathers a partly filled bucket and waits for everything in flight.  Every rank must use the same
    ``b_local`` (pad ragged shards; ``gather_logits`` handles the ragged one-shot case).

    With ``keep=True`` the gathered logits are handed out by ``pop()``: a list of (world, B_local, L) tensors in
    micro-batch order — ``t.reshape(-1, L)`` is the (global_batch, L) tensor ``gather_logits`` returns.  They are views of
    the receive buffer of their bucket, valid until that bucket is filled again (two buckets later)."""

    def __init__(self, b_local: int, n_logits: int, steps: int, device, group=None, keep: bool = False):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.steps, self.b, self.L, self.keep = int(steps), int(b_local), int(n_logits), keep
        self.send = [torch.empty((self.steps, self.b, self.L), dtype=torch.float32, device=device) for _ in range(2)]
        self.recv = [torch.empty((self.world, self.steps, self.b, self.L), dtype=torch.float32, device=device) for _ in range(2)]
        self.work = [None, None]
        self.filled = [0, 0]
        self.cur, self.n = 0, 0
        self._order = []          # launched buckets whose gather has not been waited for, oldest first
        self._done = []

    def slot(self) -> torch.Tensor:
        """The (B_local, L) view the next forward writes its logits into."""
        if self.n == 0:
            self._retire(self.cur)          # the previous gather of this bucket must have read the send buffer
        return self.send[self.cur][self.n]

    def peers(self):
        """Further destinations the forward itself should store the slot's logits to (``forward_packed(x, out=slot(),
        also=peers())``; call after ``slot()``).  None here: the exchange is a separate step of this class."""
        return None

    def commit(self) -> None:
        """The forward that fills the current slot has been enqueued on the current stream."""
        self.n += 1
        if self.n == self.steps:
            self._launch()

    def _launch(self) -> None:
        i = self.cur
        self.filled[i] = self.n
        if self.world > 1:
            self.work[i] = dist.all_gather_into_tensor(self.recv[i].view(self.world * self.steps, self.b, self.L), self.send[i],
                                                        group=self.group, async_op=True)
        else:
            self.recv[i][0].copy_(self.send[i], non_blocking=True)
        self._order.append(i)
        self.cur ^= 1
        self.n = 0

    def _retire(self, i: int) -> None:
        while i in self._order:
            j = self._order.pop(0)
            if self.work[j] is not None:
                self.work[j].wait()
                self.work[j] = None
            if self.keep:
                self._done.extend(self.recv[j][:, s] for s in range(self.filled[j]))

    def flush(self) -> None:
        """Gather what is in the current bucket and wait for every gather in flight."""
        if self.n > 0:
            self._launch()
        while self._order:
            self._retire(self._order[0])

    def pop(self):
        out, self._done = self._done, []
        return out


class PeerBucketedGather(BucketedGather):
    """BucketedGather whose exchange uses NO SMs: the bucket's send area lives in symmetric (peer-mapped) device memory
    (``torch.distributed._symmetric_memory``) and every rank PULLS its peers' buckets over NVLink with copy-engine
    ``cudaMemcpyAsync`` transfers on up to four side streams, fenced by two one-CTA signal-pad barriers (bucket complete on
    every rank / every rank has pulled; only the first is on the critical path of the results).  An NCCL all-gather kernel needs several CTAs of 512+ threads; beside the forward's persistent
    148-CTA kernels (one CTA per SM, registers and shared memory nearly exhausted) it can only be placed as forward CTAs
    retire and then delays the next kernel's wave — measured 0.014 ms per 1024-image step on 2 B200 (4.6 %).  The copy engines
    and a 32-thread barrier kernel do not compete for those resources.  Same interface as BucketedGather; CUDA + NVLink
    peer access only (construction raises otherwise — callers fall back to BucketedGather)."""

    def __init__(self, b_local: int, n_logits: int, steps: int, device, group=None, keep: bool = False):
        import torch.distributed._symmetric_memory as symm
        if not dist.is_initialized() or dist.get_world_size(group) < 2:
            raise RuntimeError("PeerBucketedGather needs an initialised process group with at least 2 ranks")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.steps, self.b, self.L, self.keep = int(steps), int(b_local), int(n_logits), keep
        shape = (2, self.steps, self.b, self.L)
        self._sym = symm.empty(shape, dtype=torch.float32, device=device)
        self._hdl = symm.rendezvous(self._sym, self.group)
        self._peer = [self._hdl.get_buffer(p, shape, torch.float32) for p in range(self.world)]
        self.send = [self._sym[0], self._sym[1]]
        self.recv = [torch.empty((self.world, self.steps, self.b, self.L), dtype=torch.float32, device=device) for _ in range(2)]
        self._free = [None, None]
        self.xs = torch.cuda.Stream(device)
        self.cs = [torch.cuda.Stream(device) for _ in range(min(4, self.world - 1))]    # peer pulls run side by side
        self.work = [None, None]          # (results ready, send area free) events of bucket i
        self.filled = [0, 0]
        self.cur, self.n = 0, 0
        self._order, self._done = [], []

    def _launch(self) -> None:
        i = self.cur
        self.filled[i] = self.n
        ready = torch.cuda.Event()
        ready.record()                                     # the forwards that fill this bucket are enqueued
        with torch.cuda.stream(self.xs):
            self.xs.wait_event(ready)
            self.recv[i][self.rank].copy_(self.send[i], non_blocking=True)     # own slot: local copy, needs no peer
            self._hdl.barrier(channel=i)                   # bucket i is complete on every rank
            go = torch.cuda.Event()
            go.record()
        pulled = []
        for d in range(1, self.world):                     # staggered sources: no two ranks pull from the same peer at once
            p = (self.rank + d) % self.world
            st = self.cs[(d - 1) % len(self.cs)]
            with torch.cuda.stream(st):
                st.wait_event(go)
                self.recv[i][p].copy_(self._peer[p][i], non_blocking=True)
                if d > self.world - 1 - len(self.cs):      # last copy on this stream
                    e = torch.cuda.Event()
                    e.record()
                    pulled.append(e)
        with torch.cuda.stream(self.xs):
            for e in pulled:
                self.xs.wait_event(e)
            got = torch.cuda.Event()
            got.record()                                   # every peer's bucket has landed here: the results are readable
            self._hdl.barrier(channel=2 + i)               # every rank has pulled: send area i may be overwritten
            free = torch.cuda.Event()
            free.record()
        self.work[i] = (got, free)
        self._order.append(i)
        self.cur ^= 1
        self.n = 0

    def slot(self) -> torch.Tensor:
        if self.n == 0:
            self._retire(self.cur)
            if self._free[self.cur] is not None:           # peers may still be reading this send area (two buckets ago)
                if not self._free[self.cur].query():
                    torch.cuda.current_stream().wait_event(self._free[self.cur])
                self._free[self.cur] = None
        return self.send[self.cur][self.n]

    def _retire(self, i: int) -> None:
        while i in self._order:
            j = self._order.pop(0)
            if self.work[j] is not None:
                got, free = self.work[j]
                if not got.query():                        # a completed event needs no stream op (keeps the forwards' PDL chain intact)
                    torch.cuda.current_stream().wait_event(got)
                self._free[j] = free
                self.work[j] = None
            if self.keep:
                self._done.extend(self.recv[j][:, s] for s in range(self.filled[j]))

class PeerSlots:
    """Pre-validated extra destinations for ``forward_packed(..., also=…)``: a ctypes array of device pointers to contiguous
    (B, Σncls) fp32 buffers (``tensors`` keeps them alive)."""
    __slots__ = ("ptrs", "n", "shape", "tensors")

    def __init__(self, ptrs, n, shape, tensors):
        self.ptrs, self.n, self.shape, self.tensors = ptrs, n, shape, tensors


class PushBucketedGather(BucketedGather):
    """MEASURED SLOWER than PeerBucketedGather's copy-engine pull (2 B200: 0.2899 vs 0.2815 ms per 1024-image step, 8 B200: 0.3029
    vs 0.3000) and therefore not the default: a kernel is not complete before its remote writes have been flushed, so the NVLink
    round trip of the epilogue's stores sits between every forward and the next one, where the pull keeps all of the exchange
    off the compute stream.  Kept as the reference point for that comparison (``make_bucketed_gather(kind="push")``).

    The exchange fused into the forward: there is no collective, no copy and no extra kernel on the data path.  The
    receive buckets live in symmetric (peer-mapped) memory and the token-stage kernel of every forward stores its logits
    into its slot on EVERY rank from its own epilogue (``forward_packed(x, out=slot(), also=peers())`` →
    ``pmg_forward_bcast``: 208 B per image and peer as posted NVLink writes).  What is left per bucket is synchronisation,
    two one-CTA signal-pad barriers on side streams: "every rank's forwards of this bucket are complete" (the results are
    readable) and, one bucket later, "every rank has taken its results out" (the index may be written again — remote ranks
    write into this rank's memory, so reuse needs everybody's consent).  Three bucket indices: the second barrier of an
    index has a whole bucket of time before anybody waits for it.  ``pop()`` hands out copies (a remote rank may be writing
    the next use of a buffer while this rank still holds its results).  Same interface as BucketedGather; CUDA + NVLink peer
    access only (construction raises otherwise)."""

    NBUF = 3

    def __init__(self, b_local: int, n_logits: int, steps: int, device, group=None, keep: bool = False):
        import torch.distributed._symmetric_memory as symm
        if not dist.is_initialized() or dist.get_world_size(group) < 2:
            raise RuntimeError("PushBucketedGather needs an initialised process group with at least 2 ranks")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world - 1 > 15:
            raise RuntimeError("pmg_forward_bcast takes at most 15 peer destinations")
        self.steps, self.b, self.L, self.keep = int(steps), int(b_local), int(n_logits), keep
        shape = (self.NBUF, self.world, self.steps, self.b, self.L)
        self._sym = symm.empty(shape, dtype=torch.float32, device=device)
        self._hdl = symm.rendezvous(self._sym, self.group)
        self._peer = [self._hdl.get_buffer(p, shape, torch.float32) for p in range(self.world)]
        self.recv = [self._sym[i] for i in range(self.NBUF)]        # (world, steps, b, L): slot [r] is written by rank r
        self.xs, self.fs = torch.cuda.Stream(device), torch.cuda.Stream(device)
        self.work = [None] * self.NBUF        # event: bucket in index i is complete on every rank
        self.free = [None] * self.NBUF        # event: every rank has taken the results of index i out
        self.filled = [0] * self.NBUF
        self.cur, self.n = 0, 0
        self._order, self._done = [], []
        self._cache = {}

    def _views(self):
        """(local slot, peers' slots as a ready-made pointer array) of the current (index, step): built once — a forward is
        enqueued in ~80 us of host time, seven tensor views and their checks per step would add half of that again."""
        key = (self.cur, self.n)
        v = self._cache.get(key)
        if v is None:
            import ctypes
            i, n, r = self.cur, self.n, self.rank
            local = self.recv[i][r][n]
            remote = [self._peer[p][i][r][n] for p in range(self.world) if p != r]
            ptrs = (ctypes.c_void_p * len(remote))(*[t.data_ptr() for t in remote])
            v = self._cache[key] = (local, PeerSlots(ptrs, len(remote), tuple(local.shape), remote))
        return v

    def slot(self) -> torch.Tensor:
        if self.n == 0:
            i = self.cur
            self._retire(i)                                   # (only on the first laps / after a flush: normally retired a bucket ago)
            self._retire((i + 1) % self.NBUF)                 # the bucket two launches back: hand its results out NOW, so that
            #                                                   everybody's consent for that index is in long before its next use
            if self.free[i] is not None:
                if not self.free[i].query():
                    torch.cuda.current_stream().wait_event(self.free[i])
                self.free[i] = None
        return self._views()[0]

    def peers(self):
        return self._views()[1]

    def _launch(self) -> None:
        i = self.cur
        self.filled[i] = self.n
        done = torch.cuda.Event()
        done.record()                                          # this rank's forwards of the bucket (and their remote stores) are enqueued
        with torch.cuda.stream(self.xs):
            self.xs.wait_event(done)
            self._hdl.barrier(channel=i)                       # … and complete on every rank: the bucket is readable everywhere
            got = torch.cuda.Event()
            got.record()
        self.work[i] = got
        self._order.append(i)
        self.cur = (i + 1) % self.NBUF
        self.n = 0

    def _retire(self, i: int) -> None:
        while i in self._order:
            j = self._order.pop(0)
            got = self.work[j]
            self.work[j] = None
            if not got.query():
                torch.cuda.current_stream().wait_event(got)
            if self.keep:
                self._done.extend(self.recv[j][:, s].clone() for s in range(self.filled[j]))
            taken = torch.cuda.Event()
            taken.record()                                     # this rank no longer needs index j …
            with torch.cuda.stream(self.fs):
                self.fs.wait_event(taken)
                self._hdl.barrier(channel=self.NBUF + j)       # … and neither does anybody else: it may be written again
                free = torch.cuda.Event()
                free.record()
            self.free[j] = free


def make_bucketed_gather(b_local: int, n_logits: int, steps: int, device, group=None, keep: bool = False, prefer_peer: bool = True,
                         kind: str = "pull"):
    """(gather, description): the copy-engine pull (``kind="pull"``: PeerBucketedGather, the default — measured faster) or the
    exchange fused into the forward (``kind="push"``: PushBucketedGather) when the ranks can map each other's memory, else —
    or with ``prefer_peer=False`` — the NCCL / gloo all-gather.  Every rank must take the same branch: the outcome of the attempt is agreed on with one
    all-reduce."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = torch.device(device)
    if prefer_peer and world > 1 and dev.type == "cuda":
        cls = PushBucketedGather if kind == "push" else PeerBucketedGather
        g, ok = None, 1
        try:
            g = cls(b_local, n_logits, steps, dev, group, keep)
        except Exception:                                  # no peer access / symmetric memory unavailable on this box
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 1:
            return g, ("logits stored to every rank's symmetric-memory bucket by the token-stage kernel itself (pmg_forward_bcast), "
                       "2 one-CTA barriers per bucket on side streams" if kind == "push" else
                       "peer-memory pull (symmetric memory + copy engines, 2 one-CTA barriers per bucket)")
        del g
    return BucketedGather(b_local, n_logits, steps, device, group, keep), "all_gather_into_tensor"


class ShardedPMG:
    """Wraps a (replicated) PMG module; ``infer(x_local)`` runs the local shard and gathers the logits."""

    def __init__(self, model, group=None):
        self.model = model
        self.group = group

    @torch.no_grad()
    def infer_local(self, x_local: torch.Tensor) -> torch.Tensor:
        return self.model.forward_packed(x_local)

    @torch.no_grad()
    def infer(self, x_local: torch.Tensor, global_batch: int) -> torch.Tensor:
        return gather_logits(self.infer_local(x_local), global_batch, self.group)

    @torch.no_grad()
    def infer_async(self, x_local: torch.Tensor, global_batch: int) -> PendingGather:
        """Forward of the local shard + non-blocking gather; call ``.result()`` when the logits are needed."""
        return gather_logits_async(self.infer_local(x_local), global_batch, self.group)

    @torch.no_grad()
    def infer_stream(self, batches, steps_per_gather: int = 4):
        """Micro-batches of equal local size → list of (global_batch, L) logits, ``steps_per_gather`` micro-batches per
        collective (BucketedGather)."""
        bg, out = None, []
        for xb in batches:
            if bg is None:
                bg, _ = make_bucketed_gather(xb.shape[0], self.model._n_logits, steps_per_gather, xb.device, self.group, keep=True)
            self.model.forward_packed(xb, out=bg.slot(), also=bg.peers())
            bg.commit()
            out.extend(t.reshape(-1, t.shape[-1]).clone() for t in bg.pop())     # the views die when the bucket is reused
        if bg is not None:
            bg.flush()
            out.extend(t.reshape(-1, t.shape[-1]).clone() for t in bg.pop())
        return out

    def split(self, logits: torch.Tensor) -> Dict[str, torch.Tensor]:
        return self.model._split(logits)


class HostPipeline:
    """End-to-end inference from HOST memory: ``infer(x_host)`` streams a pinned (B,3,H,W) host batch
    through the GPU in chunks — H2D of chunk i+1 on a copy stream overlaps the forward of chunk i on the
    compute stream — and returns the (B, Σncls) logits in pinned host memory (one D2H per chunk).
    This is the call the end-to-end bench times; the reference's equivalent is the
    ``inputs.to(device)`` → ``model(inputs)`` → ``.cpu()`` loop of Functions/PMG_evaluation.py:124-150."""

    def __init__(self, model, chunk: int = 512, device: Optional[torch.device] = None, u8_layout: str = "NHWC",
                 tail_chunk: int = 32):
        """``x_host`` may be the normalised float batch (B,3,H,W) fp32 / bf16 — what the reference's loaders deliver — or
        the decoded uint8 images, (B,H,W,3) for ``u8_layout="NHWC"`` / (B,3,H,W) for "NCHW": then ToTensor + Normalize
        (utils/datasets_utils.py:212-214) run on the GPU (prepost.preprocess_u8) and only 150 KB/image cross PCIe.

        The wire is the bottleneck, so what a call adds to the bare copy time is the forward of the LAST chunk (nothing is
        left to overlap it with): the last ``chunk`` images go out in halving pieces down to ``tail_chunk``."""
        self.model = model
        self.u8_layout = u8_layout
        self.chunk = int(chunk)
        self.tail_chunk = max(1, min(int(tail_chunk), self.chunk))
        self.device = device or next(model.parameters()).device
        self.copy_stream = torch.cuda.Stream(self.device)
        self.compute_stream = torch.cuda.Stream(self.device)
        self._dev = [None, None]
        self._pre = [None, None]
        self._consumed = [None, None]          # compute has finished reading device buffer i (persists across calls)
        self._nchunk = 0                       # chunks enqueued so far (buffer parity runs across calls)
        self._outs = [None, None]              # pinned result buffers, alternating per call
        self._done = [None, None]
        self._ncall = 0
        self._pending = []                     # tickets submitted and not yet collected
        self._pending_B = {}

    def _schedule(self, B: int):
        """[(lo, hi)] — uniform chunks, the last one split into halving pieces (…, 64, 32, 32)."""
        edges, lo = [], 0
        while B - lo > self.chunk:
            edges.append((lo, lo + self.chunk))
            lo += self.chunk
        rest = B - lo
        while rest > 2 * self.tail_chunk:
            step = rest // 2
            edges.append((lo, lo + step))
            lo += step
            rest -= step
        if rest > self.tail_chunk:
            edges.append((lo, lo + rest - self.tail_chunk))
            lo += rest - self.tail_chunk
            rest = self.tail_chunk
        if rest:
            edges.append((lo, lo + rest))
        return edges

    def _enqueue(self, x_host: torch.Tensor) -> int:
        assert x_host.device.type == "cpu" and x_host.is_pinned(), "pass a pinned host tensor"
        B = x_host.shape[0]
        L = self.model._n_logits
        slot = self._ncall & 1
        assert slot not in [t & 1 for t in self._pending], "collect() the result submitted two calls ago first"
        if self._outs[slot] is None or self._outs[slot].shape[0] < B:
            self._outs[slot] = torch.empty((B, L), dtype=torch.float32).pin_memory()
        out = self._outs[slot]
        shape = (self.chunk,) + tuple(x_host.shape[1:])
        for i in range(2):
            if self._dev[i] is None or self._dev[i].shape != shape or self._dev[i].dtype != x_host.dtype:
                if self._consumed[i] is not None:
                    self._consumed[i].synchronize()
                self._dev[i] = torch.empty(shape, dtype=x_host.dtype, device=self.device)
        cur = torch.cuda.current_stream(self.device)
        self.copy_stream.wait_stream(cur)
        self.compute_stream.wait_stream(cur)
        for lo, hi in self._schedule(B):
            k = self._nchunk & 1
            self._nchunk += 1
            buf = self._dev[k]
            copied = torch.cuda.Event()
            with torch.cuda.stream(self.copy_stream):
                if self._consumed[k] is not None:
                    self.copy_stream.wait_event(self._consumed[k])
                buf[: hi - lo].copy_(x_host[lo:hi], non_blocking=True)
                copied.record(self.copy_stream)
            with torch.cuda.stream(self.compute_stream):
                self.compute_stream.wait_event(copied)
                xin = buf[: hi - lo]
                if xin.dtype == torch.uint8:
                    from . import prepost
                    n, hh, ww = (xin.shape[0], xin.shape[1], xin.shape[2]) if self.u8_layout == "NHWC" else (xin.shape[0], xin.shape[2], xin.shape[3])
                    if self._pre[k] is None or self._pre[k].shape[1:] != (3, hh, ww) or self._pre[k].shape[0] < n:
                        self._pre[k] = torch.empty((self.chunk, 3, hh, ww), dtype=torch.bfloat16, device=self.device)
                    xin = prepost.preprocess_u8(xin, self.u8_layout, torch.bfloat16, out=self._pre[k][:n])
                logits = self.model.forward_packed(xin)
                self._consumed[k] = torch.cuda.Event()
                self._consumed[k].record(self.compute_stream)
                out[lo:hi].copy_(logits, non_blocking=True)
        self._done[slot] = torch.cuda.Event()
        self._done[slot].record(self.compute_stream)
        ticket = self._ncall
        self._ncall += 1
        self._pending.append(ticket)
        self._pending_B[ticket] = B
        return ticket

    @torch.no_grad()
    def infer(self, x_host: torch.Tensor, sync: bool = True) -> torch.Tensor:
        """Returns the logits in pinned host memory.  ``sync=True`` (default) blocks until the last device→host copy has
        landed, i.e. the returned tensor is readable; ``sync=False`` only orders the current stream after the work."""
        t = self._enqueue(x_host)
        self._pending.remove(t)
        B = self._pending_B.pop(t)
        torch.cuda.current_stream(self.device).wait_stream(self.compute_stream)
        if sync:
            self._done[t & 1].synchronize()
        return self._outs[t & 1][:B]

    @torch.no_grad()
    def submit(self, x_host: torch.Tensor) -> int:
        """Streaming form: enqueue a batch and return a ticket without waiting.  At most two batches may be outstanding
        (two pinned result buffers); the copy engine starts on batch i+1 while the forward of batch i's last chunk and its
        device→host read are still running, so a loop ``t = submit(next); collect(prev)`` keeps the wire busy all the
        time.  ``x_host`` must stay untouched until its ticket is collected."""
        return self._enqueue(x_host)

    def collect(self, ticket: Optional[int] = None) -> torch.Tensor:
        """Blocks until the batch of ``ticket`` (default: the oldest outstanding one) is in pinned host memory; the
        returned view is valid until two more batches have been submitted."""
        if ticket is None:
            ticket = self._pending[0]
        self._pending.remove(ticket)
        B = self._pending_B.pop(ticket)
        self._done[ticket & 1].synchronize()
        return self._outs[ticket & 1][:B]


class GraphedPMG:
    """Latency path for the B = 1 / small ragged batches of the reference's own callers:
    utils/benchmark_utils.py:67-75 and utils/camera_utils.py:119-123 call ``model(x)`` one image at a time and then
    read every task's result with its own device→host sync.  Here the whole forward (6 kernel launches, plus the
    per-task softmax / arg-max kernel when ``postprocess=True``) is captured once per input shape into a CUDA graph;
    a call is one device-to-device copy into the graph's static input + one ``cudaGraphLaunch``.

    ``__call__(x)`` keeps the module contract (``Dict[task, Tensor(B, ncls)]`` on the device, views of one packed
    tensor).  With ``host_outputs=True`` the packed logits are copied to pinned host memory inside the graph and the
    call synchronises once, so the caller's 12 ``.cpu()`` reads cost nothing.  Outputs are views of static buffers:
    consume them before the next call."""

    def __init__(self, model, postprocess: bool = False, prob_threshold: float = 0.0, host_outputs: bool = False):
        self.model = model
        self.postprocess = postprocess
        self.prob_threshold = float(prob_threshold)
        self.host_outputs = host_outputs
        self._graphs = {}
        self._epoch = None

    def _capture(self, x: torch.Tensor):
        """Every captured graph owns what its kernels point at: the static input, its OWN workspace (the module's shared
        ``_ws`` is reallocated whenever a later eager call or a larger capture needs more bytes — a graph replaying into
        the freed buffer would corrupt memory), the packed-weights tensor of the capture epoch, and the outputs."""
        from . import prepost
        dev = x.device
        B, _, H, W = x.shape
        static_in = torch.empty_like(x)
        static_in.copy_(x)
        ws = torch.empty(self.model.workspace_bytes(B, H, W), dtype=torch.uint8, device=dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                       # packs weights, sets kernel attributes, warms the allocator
                lg = self.model.forward_packed(static_in, ws=ws)
                if self.postprocess:
                    prepost.postprocess(lg, self.model.num_classes_per_task, self.prob_threshold)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        host = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g), torch.no_grad():
            logits = self.model.forward_packed(static_in, ws=ws)
            post = prepost.postprocess(logits, self.model.num_classes_per_task, self.prob_threshold) if self.postprocess else None
            if self.host_outputs:
                host = torch.empty(logits.shape, dtype=logits.dtype, pin_memory=True)
                host.copy_(logits, non_blocking=True)
        return {"graph": g, "in": static_in, "logits": logits, "post": post, "host": host, "ws": ws,
                "packed": self.model._packed}

    def invalidate(self) -> None:
        """Drop every captured graph (they re-capture on the next call).  Called automatically after
        ``load_state_dict`` / ``.to()`` on the module; call it by hand after modifying a parameter in place."""
        self._graphs.clear()

    @torch.no_grad()
    def forward_packed(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("GraphedPMG needs a CUDA tensor: there is no CPU fallback")
        if self.model.training:
            raise RuntimeError("PMG B200 kernels implement the eval forward only: call model.eval()")
        epoch = getattr(self.model, "_weights_epoch", 0)
        if epoch != self._epoch:                     # weights were reloaded / moved: the graphs point at stale packed weights
            self._graphs.clear()
            self._epoch = epoch
        if x.shape[0] == 0:
            return self.model.forward_packed(x)
        key = (tuple(x.shape), x.dtype, x.device)
        st = self._graphs.get(key)
        if st is None:
            st = self._graphs[key] = self._capture(x.contiguous())
        else:
            st["in"].copy_(x, non_blocking=True)
        st["graph"].replay()
        if self.host_outputs:
            torch.cuda.current_stream(x.device).synchronize()
            return st["host"]
        return st["logits"]

    def __call__(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        return self.model._split(self.forward_packed(x))

    def predictions(self, x: torch.Tensor):
        """(probs, preds, max_probs) of prepost.postprocess for this batch (requires ``postprocess=True``)."""
        if not self.postprocess:
            raise RuntimeError("construct GraphedPMG(..., postprocess=True)")
        self.forward_packed(x)
        return self._graphs[(tuple(x.shape), x.dtype, x.device)]["post"]
