"""A whole training step -- `loss = loss_fn(model(x), y); loss.backward(); [all-reduce]; optimizer.step(); optimizer.zero_grad()`
(unetr_segmentation_3d.py:218-226) -- captured once as a CUDA graph and replayed.

The step is ~400 kernel launches issued from C++ and Python (3.3 ms of host work at configs[1] against ~6 ms of GPU time): replaying
it costs the host one `cudaGraphLaunch`, the device sees back-to-back kernel nodes without launch gaps, and a blocking `.item()` on
the loss no longer exposes the enqueue time of the next step.

What makes the step capturable (all of it lives in this package, nothing here is a different arithmetic path):
  * the UNETR autograd node keeps ONE flat gradient buffer across steps (stable gradient addresses, unetr.py) and allocates its
    workspace from the graph's private pool;
  * `FusedAdamW(capturable=True)` reads the update count from device memory and, with `mirror=model`, keeps the packed bf16 weights
    current, so the forward never needs a host-side decision about re-packing;
  * no host read-back happens inside the step (the loss stays on the device; read it after the replay).
"""
from __future__ import annotations

import torch

__all__ = ["GraphedTrainStep"]


class GraphedTrainStep:
    """`step = GraphedTrainStep(model, loss_fn, optimizer, x, y)`; then `loss = step(x, y)` runs one optimisation step on the batch
    and returns the (device) loss of that step.  `x` / `y` fix the shapes; their contents are only used for the warm-up steps, which
    are REAL optimisation steps (the optimizer state must exist before capture).  `reducer`: a parallel.GradientAllReduce to call
    between backward and the optimizer step (NCCL collectives are captured with the rest).  Drop every reference to losses / outputs of
    earlier eager steps before constructing it: a live autograd graph keeps AccumulateGrad nodes bound to the stream they were created
    on, and the capture (which runs on its own stream) then has to synchronise with that stream, which CUDA refuses during capture."""

    def __init__(self, model, loss_fn, optimizer, x: torch.Tensor, y: torch.Tensor, reducer=None, warmup: int = 3):
        if not x.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors (there is no CPU path)")
        if not getattr(optimizer, "capturable", True):
            raise RuntimeError("the optimizer must be capturable (FusedAdamW(..., capturable=True))")
        self.model, self.loss_fn, self.optimizer, self.reducer = model, loss_fn, optimizer, reducer
        self.x, self.y = x.clone(), y.clone()
        self.replays = 0
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        lib = None
        try:
            from . import _lib
            lib = _lib.load()
            l0 = lib.b200_launch_count()
        except Exception:
            l0 = 0
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()
        self.launches_per_step = int(lib.b200_launch_count() - l0) if lib is not None else 0
        # the capture itself executed nothing: the host-side step counts it advanced are taken back
        if hasattr(optimizer, "advance_host_steps"):
            optimizer.advance_host_steps(-1)

    def _eager(self):
        loss = self.loss_fn(self._logits(self.model(self.x)), self.y)
        loss.backward()
        if self.reducer is not None:
            self.reducer.reduce()
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
        return loss.detach()

    @staticmethod
    def _logits(out):
        return out[1] if isinstance(out, tuple) else out

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        if y.data_ptr() != self.y.data_ptr():
            self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        if hasattr(self.optimizer, "advance_host_steps"):
            self.optimizer.advance_host_steps(1)
        return self.loss
