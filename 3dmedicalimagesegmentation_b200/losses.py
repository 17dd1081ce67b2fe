"""Loss callables of the hot path, backed by csrc/loss.cuh.

* `DiceCELoss(to_onehot_y=True, softmax=True)`  -- monai.losses.DiceCELoss as configured at
  unetr_segmentation_3d.py:404 and called at :222; `DiceCELoss(to_onehot_y=False, sigmoid=True)` as at :480.
* `extract_triplets_more_partitions` / `BTLoss` -- unetr_ranking_pretraining_3d.py:59-133 and :202-217, same
  call signatures, including BTLoss's side effects (backward, optimizer.step, optimizer.zero_grad, returns float).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib

__all__ = ["DiceCELoss", "extract_triplets_more_partitions", "BTLoss", "ranking_loss", "configure_ranking"]


# ------------------------------------------------------------------------------------------------ DiceCE
class _DiceCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        lib = _lib.load()
        _lib.require_device(logits)
        if logits.dtype != torch.float32:
            logits = logits.float()
        logits = logits.contiguous()
        target = target.contiguous().float()
        b, c = logits.shape[:2]
        v = logits[0, 0].numel()
        scratch = torch.empty(lib.b200_dicece_scratch_bytes(b, c), dtype=torch.uint8, device=logits.device)
        out = torch.empty(3, dtype=torch.float32, device=logits.device)
        _lib.check(lib.b200_dicece_forward(_lib.ptr(logits), _lib.ptr(target), b, c, v, _lib.ptr(scratch), _lib.ptr(out),
                                           _lib.stream_ptr()), "b200_dicece_forward")
        ctx.save_for_backward(logits, target, scratch)
        ctx.terms = out
        return out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        logits, target, scratch = ctx.saved_tensors
        b, c = logits.shape[:2]
        v = logits[0, 0].numel()
        dlogits = torch.empty_like(logits)
        up = grad_out.contiguous().float().reshape(1)
        _lib.check(lib.b200_dicece_backward(_lib.ptr(logits), _lib.ptr(target), b, c, v, _lib.ptr(scratch), _lib.ptr(up),
                                            _lib.ptr(dlogits), _lib.stream_ptr()), "b200_dicece_backward")
        return dlogits, None


class _DiceCESigmoidFunction(torch.autograd.Function):
    """DiceCELoss(to_onehot_y=False, sigmoid=True) (seg:480): csrc/loss.cuh dicece_sig_* kernels."""

    @staticmethod
    def forward(ctx, logits, target):
        lib = _lib.load()
        _lib.require_device(logits)
        logits = logits.float().contiguous()
        target = target.float().contiguous()
        b, c = logits.shape[:2]
        v = logits[0, 0].numel()
        scratch = torch.empty(lib.b200_dicece_scratch_bytes(b, c), dtype=torch.uint8, device=logits.device)
        out = torch.empty(3, dtype=torch.float32, device=logits.device)
        _lib.check(lib.b200_dicece_sigmoid_forward(_lib.ptr(logits), _lib.ptr(target), b, c, v, _lib.ptr(scratch), _lib.ptr(out),
                                                   _lib.stream_ptr()), "b200_dicece_sigmoid_forward")
        ctx.save_for_backward(logits, target, scratch)
        ctx.terms = out
        return out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        logits, target, scratch = ctx.saved_tensors
        b, c = logits.shape[:2]
        v = logits[0, 0].numel()
        dlogits = torch.empty_like(logits)
        up = grad_out.contiguous().float().reshape(1)
        _lib.check(lib.b200_dicece_sigmoid_backward(_lib.ptr(logits), _lib.ptr(target), b, c, v, _lib.ptr(scratch), _lib.ptr(up),
                                                    _lib.ptr(dlogits), _lib.stream_ptr()), "b200_dicece_sigmoid_backward")
        return dlogits, None


class DiceCELoss(nn.Module):
    """monai.losses.DiceCELoss in the two configurations the reference uses:

    * `DiceCELoss(to_onehot_y=True, softmax=True)` (seg:404): softmax Dice (include_background, smooth 1e-5/1e-5, mean over
      (b,c)) + mean cross-entropy against the `[B,1,...]` label map.
    * `DiceCELoss(to_onehot_y=False, sigmoid=True)` (seg:480): sigmoid Dice against the `[B,C,...]` multi-hot
      target + cross-entropy against `argmax(target, 1)` (MONAI 0.6.0 `ce` rule for a C-channel target).

    Any other configuration raises NotImplementedError (no fallback)."""

    def __init__(self, include_background: bool = True, to_onehot_y: bool = False, sigmoid: bool = False,
                 softmax: bool = False, squared_pred: bool = False, jaccard: bool = False, reduction: str = "mean",
                 smooth_nr: float = 1e-5, smooth_dr: float = 1e-5, batch: bool = False, lambda_dice: float = 1.0,
                 lambda_ce: float = 1.0, **unused) -> None:
        super().__init__()
        common = (include_background and not squared_pred and not jaccard and reduction == "mean" and smooth_nr == 1e-5
                  and smooth_dr == 1e-5 and not batch and lambda_dice == 1.0 and lambda_ce == 1.0 and not unused)
        if common and to_onehot_y and softmax and not sigmoid:
            self.variant = "softmax"
        elif common and not to_onehot_y and sigmoid and not softmax:
            self.variant = "sigmoid"
        else:
            raise NotImplementedError(
                "b200 DiceCELoss implements DiceCELoss(to_onehot_y=True, softmax=True) (seg:404) and "
                "DiceCELoss(to_onehot_y=False, sigmoid=True) (seg:480) with MONAI defaults")

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if self.variant == "sigmoid":
            if target.shape != input.shape:
                raise AssertionError(f"ground truth has differing shape ({target.shape}) from input ({input.shape})")
            if input.shape[1] > 16:
                raise NotImplementedError("sigmoid DiceCELoss takes at most 16 channels")
            return _DiceCESigmoidFunction.apply(input, target)
        if target.shape[1] != 1 or target.shape[0] != input.shape[0] or target.shape[2:] != input.shape[2:]:
            raise AssertionError(f"ground truth has differing shape ({target.shape}) from input ({input.shape})")
        return _DiceCEFunction.apply(input, target)


# ------------------------------------------------------------------------------------------------ ranking
_RANK = {"num_partitions": 4, "temperature": 0.1, "verbose": False}


def configure_ranking(temperature: float = None, num_partitions: int = None, verbose: bool = None):
    """The reference reads module globals `temperature` (rank:327) and `num_partitions` (rank:330)."""
    if temperature is not None:
        _RANK["temperature"] = float(temperature)
    if num_partitions is not None:
        if num_partitions != 4:
            raise NotImplementedError("the fused ranking kernel enumerates the reference's 4 partitions x 4 samples")
        _RANK["num_partitions"] = 4
    if verbose is not None:
        _RANK["verbose"] = bool(verbose)


class TripletList(list):
    """What `extract_triplets_more_partitions` returns: 576 slice ids (0..15 = partition*4 + sample) plus the plan the
    fused kernel needs.  The reference materialises 3 x 576 tensor views here; the ids are the same information."""
    plan = None


class _RankFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, batch1, batch2, slice_dimension, idx, temperature):
        lib = _lib.load()
        _lib.require_device(batch1)
        # [C,X,Y,Z] samples in the order rank:80-84 uses; dense so values and gradients share one geometry
        vols = [v.float().contiguous() for v in (batch1[0], batch1[1], batch2[0], batch2[1])]
        ref = vols[0]
        if any(v.shape != ref.shape for v in vols):
            raise ValueError("the four samples must have one shape")
        axis = slice_dimension - 1
        free = [a for a in (1, 2, 3) if a != axis]
        g = _lib.RankGeom()
        for i, v in enumerate(vols):
            g.src[i] = v.data_ptr()
            g.grad[i] = None
        g.stride_c, g.stride_slice = ref.stride(0), ref.stride(axis)
        g.stride_f0, g.stride_f1 = ref.stride(free[0]), ref.stride(free[1])
        g.channels, g.f0, g.f1 = ref.shape[0], ref.shape[free[0]], ref.shape[free[1]]
        for i in range(4):
            if not 0 <= int(idx[i]) < ref.shape[axis]:
                raise IndexError("slice index out of range")
            g.idx[i] = int(idx[i])
        g.temperature = float(temperature)
        scratch = torch.empty(lib.b200_ranking_scratch_bytes(ref.shape[0]), dtype=torch.uint8, device=ref.device)
        out = torch.empty(1, dtype=torch.float32, device=ref.device)
        _lib.check(lib.b200_ranking_forward(ctypes.byref(g), _lib.ptr(scratch), _lib.ptr(out), _lib.stream_ptr()),
                   "b200_ranking_forward")
        ctx.geom, ctx.vols, ctx.scratch = g, vols, scratch
        ctx.shapes = (tuple(batch1.shape), tuple(batch2.shape))
        return out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        g, vols = ctx.geom, ctx.vols
        dev = vols[0].device
        g1 = torch.zeros(ctx.shapes[0], dtype=torch.float32, device=dev)
        g2 = torch.zeros(ctx.shapes[1], dtype=torch.float32, device=dev)
        for i, o in enumerate((g1[0], g1[1], g2[0], g2[1])):
            g.grad[i] = o.data_ptr()           # dense [C,X,Y,Z] views: same strides as the forward copies
        up = grad_out.contiguous().float().reshape(1)
        _lib.check(lib.b200_ranking_backward(ctypes.byref(g), _lib.ptr(ctx.scratch), _lib.ptr(up), _lib.stream_ptr()),
                   "b200_ranking_backward")
        return g1, g2, None, None, None


def ranking_loss(batch1, batch2, slice_dimension, slice_idx_list, temperature=None):
    """sum over the 576 triplets of mean_c log(1+exp(-(cos(r,s)-cos(r,d))/T))  as a differentiable 0-d tensor."""
    t = _RANK["temperature"] if temperature is None else temperature
    return _RankFunction.apply(batch1, batch2, int(slice_dimension), [int(i) for i in slice_idx_list], float(t))


def extract_triplets_more_partitions(batch1, batch2, slice_dimension=2):
    """Same draws from the global numpy RNG as rank:73-75; returns (reference, similar, dissimilar) id lists."""
    if slice_dimension not in (2, 3, 4):
        raise ValueError("slice_dimension must be 2, 3 or 4")
    if batch1.shape[0] < 2 or batch2.shape[0] < 2:
        raise IndexError("each half needs the two transforms of a volume (rank:80-83)")
    npart = _RANK["num_partitions"]
    dims = batch1.shape
    partition_size = int(dims[slice_dimension] / npart)
    init_idx = np.random.choice(np.arange(0, partition_size))
    slice_idx_list = [int(init_idx + p * partition_size) for p in range(npart)]
    if _RANK["verbose"]:
        print("Shapes of loss batch pair", batch1.shape, batch2.shape)
        print("Slicing dimension", slice_dimension)
        print("Slice indices:", slice_idx_list)
    ref, sim, dis = TripletList(), TripletList(), TripletList()
    for p in range(npart):
        mine = [p * 4 + j for j in range(4)]
        others = [q * 4 + j for q in range(npart) if q != p for j in range(4)]
        for r in mine:
            for s in mine:
                if s == r:
                    continue
                for d in others:
                    ref.append(r); sim.append(s); dis.append(d)
    plan = (batch1, batch2, slice_dimension, slice_idx_list)
    ref.plan = sim.plan = dis.plan = plan
    return ref, sim, dis


def BTLoss(reference, similar, dissimilar, optimizer):
    """Bradley-Terry ranking loss step (rank:202-217): loss -> backward -> optimizer.step -> zero_grad -> float."""
    plan = getattr(reference, "plan", None)
    if plan is None or getattr(similar, "plan", None) is not plan or getattr(dissimilar, "plan", None) is not plan:
        raise TypeError("BTLoss expects the three lists returned together by this package's "
                        "extract_triplets_more_partitions (there is no per-view fallback path)")
    batch1, batch2, sd, idx = plan
    loss = ranking_loss(batch1, batch2, sd, idx)
    loss.backward()
    optimizer.step()
    optimizer.zero_grad()
    return loss.item()
