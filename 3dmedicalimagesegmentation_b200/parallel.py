"""Multi-GPU plumbing: one process per GPU, torch.distributed over NCCL/NVLink (gloo on CPU for the tests).

The reference is single-GPU (no distributed code).  The path shards in exactly two places:
  * training: samples are independent (InstanceNorm is per sample) -> data parallel, ONE gradient all-reduce of the
    92.45 M fp32 gradients per step, issued as a few large flat buckets;
  * sliding-window inference: windows are independent -> contiguous chunks of the window list per rank
    (inferers.sliding_window_inference(rank=, world_size=)).
"""
from __future__ import annotations

import os
from typing import List

import torch
import torch.distributed as dist

__all__ = ["init_from_env", "barrier", "max_over_ranks", "shutdown", "GradientAllReduce"]


def init_from_env(backend: str = None):
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def barrier(world: int):
    if world > 1:
        dist.barrier()


def max_over_ranks(value: float, world: int, device) -> float:
    if world == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def shutdown(world: int):
    if world > 1 and dist.is_initialized():
        dist.destroy_process_group()


class GradientAllReduce:
    """Averages gradients across ranks after `backward()`.

    The UNETR autograd node hands back all parameter gradients as views of ONE flat fp32 buffer, so when that
    buffer is found the whole reduction is a single all-reduce (369.8 MB at 96^3); otherwise gradients are packed
    into `bucket_mb` buckets.  Parameters whose grad is None on every rank are skipped (ranking stages)."""

    def __init__(self, module: torch.nn.Module, world_size: int, bucket_mb: int = 512, group=None, overlap: bool = True, compress: str = None):
        self.params: List[torch.nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        self.world, self.group = world_size, group
        self.bucket_elems = bucket_mb * 1024 * 1024 // 4
        self.module = module
        self.side = None
        # compress="bf16" (overlapped path only): every gradient group crosses NVLink as bf16 -- cast on a side stream behind the group's
        # event, all-reduce (AVG) of half the bytes, cast back into the fp32 gradient views on a second side stream.  The averaged
        # gradients the optimizer sees are bf16-rounded (relative 2^-9 per element, the precision the bf16 engine computed them from);
        # fp32 mode and the correctness checks keep fp32 on the wire.  At 8 GPUs the 340 MB of ViT gradients become final in the last
        # millisecond of the backward: fp32 needs ~410 GB/s of bus bandwidth to hide behind it, bf16 half of that.
        if compress not in (None, "bf16"):
            raise ValueError("compress must be None or 'bf16'")
        self.compress = compress
        self._stage16 = None
        self.post = None
        # overlap: the UNETR backward records an event when each of its gradient groups (UNETR.grad_groups: 7 by default) is final; their all-reduces run on a side
        # stream behind those events while the rest of the backward is still executing (the host enqueues far ahead of the GPU)
        if overlap and world_size > 1 and hasattr(module, "overlap_grad_reduce") and torch.cuda.is_available():
            module.overlap_grad_reduce = True
            self.side = torch.cuda.Stream()
            self.post = torch.cuda.Stream()

    def _avg(self, t):
        if dist.get_backend(self.group) == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        w = dist.all_reduce(t, group=self.group, async_op=True)
        w.wait()
        t.div_(self.world)
        return None

    def _flat_view(self, grads):
        """All grads contiguous slices of one storage, in order -> return the covering flat tensor."""
        if not grads:
            return None
        base = grads[0].untyped_storage().data_ptr()
        start = grads[0].data_ptr()
        off = start
        for g in grads:
            if g.untyped_storage().data_ptr() != base or g.data_ptr() != off or not g.is_contiguous():
                return None
            off += g.numel() * g.element_size()
        n = (off - start) // 4
        return torch.as_strided(grads[0], (n,), (1,))

    class _After:
        """completion of one compressed group: the current stream waits for the event recorded after the cast back"""

        def __init__(self, ev):
            self.ev = ev

        def wait(self):
            torch.cuda.current_stream().wait_event(self.ev)

    def _issue(self, flat, ev, lo, hi):
        """on the side stream, behind `ev`: all-reduce flat[lo:hi]; returns an object whose .wait() orders the current stream after it"""
        self.side.wait_event(ev)
        if self.compress is None:
            return self._avg(flat[lo:hi])
        if self._stage16 is None or self._stage16.numel() < flat.numel() or self._stage16.device != flat.device:
            self._stage16 = torch.empty(flat.numel(), dtype=torch.bfloat16, device=flat.device)
        c = self._stage16[lo:hi]
        c.copy_(flat[lo:hi])                                   # side stream: fp32 -> bf16
        work = self._avg(c)
        done = torch.cuda.Event()
        with torch.cuda.stream(self.post):
            if work is not None:
                work.wait()                                    # post stream waits for the collective
            flat[lo:hi].copy_(c)                               # bf16 -> fp32 into the gradient views
            done.record(self.post)
        return self._After(done)

    def reduce_and_step(self, optimizer):
        """`reduce(); optimizer.step()` with the optimizer update of every gradient group launched behind that group's own all-reduce
        (optim.FusedAdamW.step(grad_ranges=)): the update of the early groups overlaps the all-reduce of the last one, which has no
        backward left to hide behind.  Any other optimizer, or a backward that did not record group events: plain reduce + step."""
        ready = getattr(self.module, "_grad_ready", None)
        ok = self.world > 1 and ready is not None and self.side is not None and hasattr(optimizer, "advance_host_steps")
        if ok:
            flat, groups, views = ready
            ok = not any(p.grad is None or p.grad.data_ptr() != ptr for p, ptr in views)
        if not ok:
            self.reduce()
            return optimizer.step()
        self.module._grad_ready = None
        ranges = []
        base = flat.data_ptr()
        with torch.cuda.stream(self.side):
            for ev, lo, hi in groups:
                if hi > lo:
                    ranges.append((self._issue(flat, ev, lo, hi), base + 4 * lo, base + 4 * hi))
        return optimizer.step(grad_ranges=ranges)

    def reduce(self):
        if self.world == 1:
            return
        ready = getattr(self.module, "_grad_ready", None)
        if ready is not None and self.side is not None:
            flat, groups, views = ready
            self.module._grad_ready = None
            # the events cover ranges of the autograd node's flat buffer: the averages only reach the parameters if every p.grad
            # IS the view of it that the backward handed out (true when AccumulateGrad stole the views; false under gradient
            # accumulation, zero_grad(set_to_none=False) or a cloned gradient) -- otherwise take the generic path below
            if any(p.grad is None or p.grad.data_ptr() != ptr for p, ptr in views):
                ready = None
        if ready is not None and self.side is not None:
            works = []
            with torch.cuda.stream(self.side):
                for ev, lo, hi in groups:
                    if hi > lo:
                        works.append(self._issue(flat, ev, lo, hi))
            for w in works:
                if w is not None:
                    w.wait()              # the current stream waits for the reduction
            return
        grads = [p.grad for p in self.params if p.grad is not None]
        # one covering view if the gradients are consecutive slices of one buffer in ADDRESS order (the UNETR node lays them out in
        # its own parameter-table order, which is the same on every rank); the bucket path below must keep the rank-independent
        # parameter order
        flat = self._flat_view(sorted(grads, key=lambda g: g.data_ptr()))
        if flat is not None:
            w = self._avg(flat)
            if w is not None:
                w.wait()
            return
        bucket, size = [], 0
        for g in grads + [None]:
            if g is None or size + g.numel() > self.bucket_elems:
                if bucket:
                    buf = torch.cat([b.reshape(-1) for b in bucket])
                    dist.all_reduce(buf, group=self.group)
                    buf.div_(self.world)
                    o = 0
                    for b in bucket:
                        b.copy_(buf[o:o + b.numel()].view_as(b))
                        o += b.numel()
                bucket, size = [], 0
            if g is not None:
                bucket.append(g)
                size += g.numel()
