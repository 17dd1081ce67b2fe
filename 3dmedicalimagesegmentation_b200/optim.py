"""Fused multi-tensor AdamW -- the optimizer step that follows the hot path in both training scripts
(`torch.optim.AdamW(model.parameters(), lr, weight_decay)` at unetr_segmentation_3d.py:522 / unetr_ranking_pretraining_3d.py:466,
`optimizer.step(); optimizer.zero_grad()` at seg:225-226 / rank:214-215).

Same arithmetic and the same skip rule as torch.optim.AdamW (parameters whose `.grad is None` keep their state: the ranking stages
rely on it), but ONE kernel launch over all parameters instead of torch's multi_tensor_apply passes."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

__all__ = ["FusedAdamW"]


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}          # group index -> (key, tensors_dev, chunks_dev, n_chunks)

    def _build(self, gi, plist):
        chunk = _lib.load().b200_adamw_chunk()
        rows, chunks = [], []
        for ti, p in enumerate(plist):
            st = self.state[p]
            rows.append((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()))
            for s in range(0, p.numel(), chunk):
                chunks.append((ti, s))
        tens = np.array(rows, dtype=np.int64)                                   # {p, g, m, v, n}: five 8-byte fields
        ch = np.zeros(len(chunks), dtype=[("t", "<i4"), ("pad", "<i4"), ("s", "<i8")])
        ch["t"] = [c[0] for c in chunks]; ch["s"] = [c[1] for c in chunks]
        dev = plist[0].device
        tens_d = torch.from_numpy(tens.view(np.uint8).reshape(-1)).to(dev)
        ch_d = torch.from_numpy(ch.view(np.uint8).reshape(-1)).to(dev)
        return tens_d, ch_d, len(chunks)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous() or p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdamW needs contiguous fp32 CUDA parameters and gradients")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            steps = {self.state[p]["step"] for p in plist}
            # one launch per distinct step count (all equal in ordinary training; ranking stages that wake parameters up later differ)
            for si, s0 in enumerate(sorted(steps)):
                sub = [p for p in plist if self.state[p]["step"] == s0]
                key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in sub)      # tables are rebuilt only when a pointer moves
                ent = self._tables.get((gi, si))
                if ent is None or ent[0] != key:
                    ent = (key,) + self._build(gi, sub)
                    self._tables[(gi, si)] = ent
                b1, b2 = group["betas"]
                _lib.check(lib.b200_adamw_step(_lib.ptr(ent[1]), _lib.ptr(ent[2]), ent[3], float(group["lr"]), float(b1), float(b2),
                                               float(group["eps"]), float(group["weight_decay"]), s0 + 1, _lib.stream_ptr()), "b200_adamw_step")
                for p in sub:
                    self.state[p]["step"] = s0 + 1
                # the kernel wrote through raw pointers: bump the autograd version counters so that everything keyed on them (the
                # UNETR inference cache of packed bf16 weights, saved-tensor checks) sees the in-place update
                torch.autograd.graph.increment_version(sub)
        return loss
