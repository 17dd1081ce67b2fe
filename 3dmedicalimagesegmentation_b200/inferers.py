"""Sliding-window whole-volume inference on the GPU -- `monai.inferers.sliding_window_inference` as called at
unetr_segmentation_3d.py:109 (positional, overlap 0.25), :143 and :694-695 (keyword, overlap 0.8).

Window enumeration, padding, scan interval and accumulation order follow MONAI 0.6.0; the
gather / overlap-add / divide-by-count arithmetic runs in csrc/sliding.cuh.  `rank`/`world_size` shard the window
list across GPUs (windows are independent; the overlap-add is completed with one all-reduce).
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Callable, List, Sequence, Tuple

import torch

from . import _lib

__all__ = ["sliding_window_inference", "window_starts", "shard_windows"]


def _scan_interval(image_size, roi, overlap):
    return tuple(int(r) if r == s else max(int(r * (1 - overlap)), 1) for r, s in zip(roi, image_size))


def _axis_starts(size, roi, step) -> List[int]:
    num = int(math.ceil(float(size) / step))
    scan = next((d for d in range(num) if d * step + roi >= size), -1)
    n = scan + 1 if scan != -1 else 1
    return [i * step - max(i * step + roi - size, 0) for i in range(n)]


def window_starts(image_size, roi, overlap):
    """Per-axis window starts and the flat list in MONAI order (first spatial axis slowest)."""
    step = _scan_interval(image_size, roi, overlap)
    per_axis = [_axis_starts(s, r, st) for s, r, st in zip(image_size, roi, step)]
    flat = [(a, b, c) for a in per_axis[0] for b in per_axis[1] for c in per_axis[2]]
    return per_axis, flat


def shard_windows(n_items: int, rank: int, world_size: int) -> range:
    """Contiguous chunk of the ij-ordered window list owned by `rank` (x-slabs)."""
    per = (n_items + world_size - 1) // world_size
    return range(min(rank * per, n_items), min((rank + 1) * per, n_items))


def sliding_window_inference(inputs: torch.Tensor, roi_size, sw_batch_size: int, predictor: Callable,
                             overlap: float = 0.25, mode: str = "constant", sigma_scale=0.125,
                             padding_mode: str = "constant", cval: float = 0.0, sw_device=None, device=None,
                             *args, rank: int = 0, world_size: int = 1, process_group=None,
                             return_argmax: bool = False, labels: torch.Tensor = None, return_logits: bool = True, **kwargs):
    """Extras beyond MONAI's signature (all keyword-only, defaults reproduce MONAI): `rank`/`world_size`/`process_group`
    shard the windows; `return_argmax` adds the uint8 class mask; `labels` ([B,1,D,H,W] class ids) fuses the validation
    tail (seg:110-126) into the normalise pass and adds the [B,C,3] counts DiceMetric/ConfusionMatrixMetric consume
    (`metric.update_from_counts`); `return_logits=False` skips writing the 3.76 GB normalised logits when only the mask /
    counts are wanted.  Return: logits | (logits, mask) | (logits, mask, counts), `None` in place of skipped logits."""
    if str(mode).lower() not in ("constant", "blendmode.constant"):
        raise NotImplementedError("only constant blending (the mode both reference call sites use) is implemented")
    if str(padding_mode).lower() not in ("constant", "pytorchpadmode.constant"):
        raise NotImplementedError("only constant padding is implemented")
    if inputs.dim() != 5:
        raise ValueError("inputs must be [B,C,D,H,W]")
    if not 0 <= overlap < 1:
        raise AssertionError("overlap must be >= 0 and < 1.")
    lib = _lib.load()
    _lib.require_device(inputs)
    x = inputs.contiguous().float()
    batch, chan = x.shape[:2]
    orig = tuple(x.shape[2:])
    roi = tuple(int(r) for r in (roi_size if isinstance(roi_size, Sequence) else (roi_size,) * 3))
    roi = tuple(r if r > 0 else o for r, o in zip(roi, orig))          # MONAI fall_back_tuple
    size = tuple(max(o, r) for o, r in zip(orig, roi))
    pad = tuple((s - o) // 2 for s, o in zip(size, orig))
    per_axis, flat = window_starts(size, roi, overlap)
    if max(len(a) for a in per_axis) > 64:
        raise NotImplementedError("more than 64 windows along one axis")
    num_win = len(flat)
    items = [(b, *flat[w]) for b in range(batch) for w in range(num_win)]   # idx -> (idx // num_win, idx % num_win)
    mine = shard_windows(len(items), rank, world_size)
    st = _lib.stream_ptr()

    gin = _lib.SwGeom(chan, *orig, *pad, *size, *roi)
    sw_batch_size = max(1, min(int(sw_batch_size), 16))
    acc = None
    gout = None
    # our own UNETR replays its forward as one CUDA graph inside this loop (each prediction is accumulated at once, so the
    # graph's static output buffers may be overwritten by the next call); B200_NO_GRAPH=1 keeps the eager launches
    use_graph = hasattr(predictor, "_graph_forward") and not getattr(predictor, "tuple_output", True) and \
        not torch.is_grad_enabled() and not os.environ.get("B200_NO_GRAPH") and len(mine) >= 4 * sw_batch_size
    prev_graph = getattr(predictor, "inference_graph", False)
    if use_graph:
        predictor.inference_graph = True
    try:
        acc, gout = _sw_loop(lib, x, items, mine, sw_batch_size, chan, roi, gin, cval, st, predictor, args, kwargs, batch, size, orig, pad)
    finally:
        if use_graph:
            predictor.inference_graph = prev_graph
    return _sw_finish(lib, acc, gout, world_size, process_group, batch, orig, per_axis, return_argmax, x, st, labels, return_logits)


def _sw_loop(lib, x, items, mine, sw_batch_size, chan, roi, gin, cval, st, predictor, args, kwargs, batch, size, orig, pad):
    acc = None
    gout = None
    for g0 in range(mine.start, mine.stop, sw_batch_size):
        chunk = items[g0:min(g0 + sw_batch_size, mine.stop)]
        n = len(chunk)
        starts = (ctypes.c_int32 * (4 * n))(*[v for it in chunk for v in it])
        win = torch.empty((n, chan, *roi), dtype=torch.float32, device=x.device)
        _lib.check(lib.b200_sw_gather(_lib.ptr(x), _lib.ptr(win), ctypes.byref(gin), starts, n, float(cval), st), "b200_sw_gather")
        pred = predictor(win, *args, **kwargs)
        if isinstance(pred, (tuple, list)):
            raise TypeError("predictor must return a tensor (monai.networks.nets.UNETR flavour, seg:36)")
        pred = pred.contiguous().float()
        if acc is None:
            cout = pred.shape[1]
            acc = torch.zeros((batch, cout, *size), dtype=torch.float32, device=x.device)
            gout = _lib.SwGeom(cout, *orig, *pad, *size, *roi)
        # one launch per run of windows of the same batch item; every accumulator voxel adds its windows in window order, so the
        # sums are bit-identical to the reference's one-window-at-a-time loop
        k0 = 0
        while k0 < n:
            k1 = k0
            while k1 < n and chunk[k1][0] == chunk[k0][0]:
                k1 += 1
            sN = (ctypes.c_int32 * (4 * (k1 - k0)))(*[v for it in chunk[k0:k1] for v in it])
            _lib.check(lib.b200_sw_accumulate_n(_lib.ptr(acc), _lib.ptr(pred[k0]), ctypes.byref(gout), sN, k1 - k0, st),
                       "b200_sw_accumulate_n")
            k0 = k1
    return acc, gout


def _sw_finish(lib, acc, gout, world_size, process_group, batch, orig, per_axis, return_argmax, x, st, labels=None,
               return_logits=True):
    if acc is None:
        raise RuntimeError("this rank owns no windows; use fewer ranks than windows")
    if world_size > 1:
        import torch.distributed as dist
        dist.all_reduce(acc, group=process_group)
    cout = acc.shape[1]
    want_mask = return_argmax or labels is not None
    out = torch.empty((batch, cout, *orig), dtype=torch.float32, device=x.device) if return_logits else None
    mask = torch.empty((batch, 1, *orig), dtype=torch.uint8, device=x.device) if want_mask else None
    counts = None
    lab = None
    if labels is not None:
        if cout > 32:
            raise NotImplementedError("fused validation counts take at most 32 classes")
        lab = labels.to(x.device).float().contiguous()
        if lab.shape[0] != batch or tuple(lab.shape[-3:]) != tuple(orig) or lab.numel() != batch * orig[0] * orig[1] * orig[2]:
            raise ValueError(f"labels {tuple(labels.shape)} do not match the volume {(batch, 1, *orig)}")
        counts = torch.empty((batch, cout, 3), dtype=torch.float64, device=x.device)
    arrs = [(ctypes.c_int32 * len(a))(*a) for a in per_axis]
    _lib.check(lib.b200_sw_finalize_metric(_lib.ptr(acc), _lib.ptr(out), _lib.ptr(mask), ctypes.byref(gout), batch,
                                           arrs[0], len(arrs[0]), arrs[1], len(arrs[1]), arrs[2], len(arrs[2]),
                                           _lib.ptr(lab), _lib.ptr(counts), st), "b200_sw_finalize_metric")
    if counts is not None:
        counts.voxels = orig[0] * orig[1] * orig[2]
        return out, mask, counts
    return (out, mask) if return_argmax else out
