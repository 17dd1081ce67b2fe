"""Fused multi-tensor AdamW -- the optimizer step that follows the hot path in both training scripts
(`torch.optim.AdamW(model.parameters(), lr, weight_decay)` at unetr_segmentation_3d.py:522 / unetr_ranking_pretraining_3d.py:466,
`optimizer.step(); optimizer.zero_grad()` at seg:225-226 / rank:214-215).

Same arithmetic and the same skip rule as torch.optim.AdamW (parameters whose `.grad is None` keep their state: the ranking stages
rely on it), but ONE kernel launch over all parameters instead of torch's multi_tensor_apply passes."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

__all__ = ["FusedAdamW"]


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2, mirror=None,
                 capturable: bool = False):
        """`mirror`: a b200 UNETR module whose parameters this optimizer updates.  In bf16 mode the same launch then also writes the
        module's packed bf16 weight copies (the operands of the tcgen05 engine), so the next forward skips its cast / re-layout
        launches.  Without it nothing changes: the forward re-packs whenever a parameter's version counter moved."""
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}          # group index -> (key, tensors_dev, chunks_dev, n_chunks)
        self._mirror = mirror
        # capturable: the update count is read from device memory (advanced by a one-thread kernel before each update), so the launch
        # sequence of a step does not change from step to step and can be replayed from a CUDA graph (graph.GraphedTrainStep).  The
        # host-side `state[p]["step"]` is advanced by every Python-driven step; replays advance it through `advance_host_steps`.
        self.capturable = bool(capturable)
        self._dev_steps = {}       # (group, cohort) -> int32 device counter
        self._last_cohorts = []

    def _mirror_rows(self, device):
        """id(parameter) -> device address of its plain bf16 mirror (parameters that have one), or {}."""
        m = self._mirror
        if m is None or getattr(m, "compute_mode", None) != "bf16" or not getattr(self, "_mirror_was_ok", False):
            return {}
        rows = m.packed_mirrors(device)
        return {id(p): r for p, r in zip(m._ordered_params(), rows) if r}

    def _build(self, gi, plist):
        chunk = _lib.load().b200_adamw_chunk()
        rows, chunks, gaddr = [], [], []
        mirrors = self._mirror_rows(plist[0].device)
        # table in gradient-address order: a range of one flat gradient buffer is then a contiguous range of chunks (step(grad_ranges=))
        plist = sorted(plist, key=lambda q: q.grad.data_ptr())
        for ti, p in enumerate(plist):
            st = self.state[p]
            rows.append((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(), mirrors.get(id(p), 0)))
            for s in range(0, p.numel(), chunk):
                chunks.append((ti, s))
                gaddr.append(p.grad.data_ptr() + 4 * s)
        tens = np.array(rows, dtype=np.int64)                                   # {p, g, m, v, n, s0}: six 8-byte fields
        ch = np.zeros(len(chunks), dtype=[("t", "<i4"), ("pad", "<i4"), ("s", "<i8")])
        ch["t"] = [c[0] for c in chunks]; ch["s"] = [c[1] for c in chunks]
        dev = plist[0].device
        tens_d = torch.from_numpy(tens.view(np.uint8).reshape(-1)).to(dev)
        ch_d = torch.from_numpy(ch.view(np.uint8).reshape(-1)).to(dev)
        return tens_d, ch_d, len(chunks), np.array(gaddr, dtype=np.int64)

    @torch.no_grad()
    def step(self, closure=None, grad_ranges=None):
        """`grad_ranges` (parallel.GradientAllReduce.reduce_and_step): [(work, lo_addr, hi_addr)] -- gradient address ranges in the order
        their all-reduces were issued.  The update is then launched range by range, each behind its own reduction, so the update of
        the early groups runs while the last all-reduce is still on the wire."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        self._last_cohorts = []
        # the packed mirrors stay current through this step only if they were current before it (every parameter unchanged since
        # the forward that packed or verified them)
        m = self._mirror
        mirror_dev, mirror_ok = None, False
        if m is not None and getattr(m, "compute_mode", None) == "bf16":
            mp = m._ordered_params()
            mirror_dev = mp[0].device
            mirror_ok = m._packed_key.get(mirror_dev) == m._version_key(mp)
            if mirror_ok != getattr(self, "_mirror_was_ok", None):
                self._tables = {}          # tables carry (or omit) the mirror pointers
                self._mirror_was_ok = mirror_ok
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
                dev_step = None
                if self.capturable:
                    dev_step = self._dev_steps.get((gi, si))
                    if dev_step is None or dev_step[1] != key:
                        dev_step = (torch.full((1,), s0, dtype=torch.int32, device=sub[0].device), key)
                        self._dev_steps[(gi, si)] = dev_step
                    self._last_cohorts.append(sub)
                hyper = (float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]))
                dptr = _lib.ptr(dev_step[0]) if dev_step else None
                if grad_ranges:
                    # chunks outside every range (none in practice) would be skipped: require full cover
                    done, first = 0, True
                    for work, lo, hi in grad_ranges:
                        c0, c1 = int(np.searchsorted(ent[4], lo, "left")), int(np.searchsorted(ent[4], hi, "left"))
                        if work is not None:
                            work.wait()
                        if c1 > c0:
                            # the device-side update count advances once per step: the first launch advances it (step > 0), the others read it
                            _lib.check(lib.b200_adamw_step_capturable(_lib.ptr(ent[1]), ctypes.c_void_p(ent[2].data_ptr() + 16 * c0), c1 - c0, *hyper,
                                                                      s0 + 1 if (first or not dev_step) else 0, dptr, _lib.stream_ptr()), "b200_adamw_step")
                            first = False
                            done += c1 - c0
                    if done != ent[3]:
                        raise RuntimeError("FusedAdamW.step(grad_ranges=): the ranges do not cover every gradient exactly once")
                else:
                    _lib.check(lib.b200_adamw_step_capturable(_lib.ptr(ent[1]), _lib.ptr(ent[2]), ent[3], *hyper, s0 + 1, dptr, _lib.stream_ptr()),
                               "b200_adamw_step")
                for p in sub:
                    self.state[p]["step"] = s0 + 1
                # the kernel wrote through raw pointers: bump the autograd version counters so that everything keyed on them (the
                # UNETR inference cache of packed bf16 weights, saved-tensor checks) sees the in-place update
                torch.autograd.graph.increment_version(sub)
        if mirror_ok:
            m.repack_convs(mirror_dev)       # the re-laid-out conv copies (one launch); the plain casts were written by the update itself
            m._packed_key[mirror_dev] = m._version_key(m._ordered_params())
        return loss

    def advance_host_steps(self, n: int = 1):
        """After `n` replays of a captured step: bring the host-side step counts of the parameters the captured step updated in line
        with the device counters (which the replays advanced)."""
        for sub in self._last_cohorts:
            for p in sub:
                self.state[p]["step"] += n
