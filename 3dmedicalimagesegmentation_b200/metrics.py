"""Validation tail of the segmentation script on the GPU -- `monai.metrics.DiceMetric` and
`monai.metrics.ConfusionMatrixMetric` as constructed at unetr_segmentation_3d.py:485-494 and driven at :110-126 / :153-188:

    metric(y_pred=[one-hot tensors], y=[one-hot tensors])   # appends one row per sample to the metric's buffer
    metric.aggregate()                                      # NaN-aware "mean" (0-d) or "mean_batch" ([C]) over the buffer
    metric.reset()

The same three integer counts per (sample, class) -- |y & p|, |p|, |y| -- feed every metric, so besides the reference's
one-hot call form each metric also takes `update_from_counts(counts)` with the counts `sliding_window_inference(...,
labels=...)` produces inside its normalise pass, or `update_from_label_maps(mask, labels)` (uint8 argmax + label map).
All arithmetic is in csrc/metrics.cuh; nothing is computed by torch ops.
"""
from __future__ import annotations

from typing import List, Sequence, Union

import torch

from . import _lib

__all__ = ["DiceMetric", "ConfusionMatrixMetric", "segmentation_counts", "segmentation_counts_from_label_maps"]

_REDUCTIONS = {"mean": 0, "mean_batch": 1}
_CONFUSION = {"precision": 0, "positive predictive value": 0, "ppv": 0, "sensitivity": 1, "recall": 1, "hit_rate": 1,
              "true positive rate": 1, "tpr": 1}


def _stack(x: Union[torch.Tensor, Sequence[torch.Tensor]]) -> torch.Tensor:
    """MONAI accepts a list of channel-first tensors (decollated batch, seg:112-119) or one batch-first tensor."""
    if isinstance(x, (list, tuple)):
        x = torch.stack([t for t in x], dim=0)
    if x.dim() < 3:
        raise ValueError("y_pred should have at least three dimensions.")
    return x


def segmentation_counts(y_pred, y) -> torch.Tensor:
    """[B,C,3] float64 counts (|y & p|, |p|, |y|) from one-hot (binarised) `y_pred` / `y` of shape [B,C,spatial...]."""
    lib = _lib.load()
    y_pred, y = _stack(y_pred), _stack(y)
    if y_pred.shape != y.shape:
        raise ValueError("y_pred and y should have same shapes.")
    _lib.require_device(y_pred)
    p = y_pred.float().contiguous()
    t = y.to(p.device).float().contiguous()
    b, c = p.shape[:2]
    v = p[0, 0].numel()
    counts = torch.empty((b, c, 3), dtype=torch.float64, device=p.device)
    _lib.check(lib.b200_seg_counts_onehot(_lib.ptr(p), _lib.ptr(t), b, c, v, _lib.ptr(counts), _lib.stream_ptr()),
               "b200_seg_counts_onehot")
    counts.voxels = v
    return counts


def segmentation_counts_from_label_maps(mask: torch.Tensor, labels: torch.Tensor, n_classes: int) -> torch.Tensor:
    """Same counts from the uint8 argmax mask [B,1,spatial] (or [B,spatial]) and the label map holding class ids -- what
    `AsDiscrete(argmax=True, to_onehot=True)` / `AsDiscrete(to_onehot=True)` (seg:405-406) expand to one-hot tensors."""
    lib = _lib.load()
    _lib.require_device(mask)
    if mask.dtype != torch.uint8:
        raise TypeError("mask must be the uint8 argmax tensor")
    if not 1 <= n_classes <= 32:
        raise NotImplementedError("1..32 classes")
    b = mask.shape[0]
    m = mask.contiguous()
    t = labels.to(m.device).float().contiguous()
    v = m[0].numel()
    if t.shape[0] != b or t[0].numel() != v:
        raise ValueError("mask and labels should have same shapes.")
    counts = torch.empty((b, n_classes, 3), dtype=torch.float64, device=m.device)
    _lib.check(lib.b200_seg_counts_labels(_lib.ptr(m), _lib.ptr(t), b, n_classes, v, _lib.ptr(counts), _lib.stream_ptr()),
               "b200_seg_counts_labels")
    counts.voxels = v
    return counts


class _CumulativeSegMetric:
    """Buffer/aggregate/reset protocol of monai.metrics.CumulativeIterationMetric for metrics derived from the counts."""

    _width = 1            # trailing size of one buffer row entry

    def __init__(self, include_background: bool = True, reduction: str = "mean", get_not_nans: bool = False) -> None:
        if not include_background:
            raise NotImplementedError("include_background=False is not used by the reference (seg:485-494)")
        if str(reduction) not in _REDUCTIONS:
            raise NotImplementedError('reduction must be "mean" or "mean_batch" (the two the reference uses)')
        self.reduction = str(reduction)
        self.get_not_nans = get_not_nans
        self._rows: List[torch.Tensor] = []

    # -- per-iteration update in the three accepted forms
    def __call__(self, y_pred, y):
        return self.update_from_counts(segmentation_counts(y_pred, y))

    def update_from_label_maps(self, mask, labels, n_classes):
        return self.update_from_counts(segmentation_counts_from_label_maps(mask, labels, n_classes))

    def update_from_counts(self, counts: torch.Tensor, voxels: int = None):
        voxels = voxels if voxels is not None else getattr(counts, "voxels", None)
        if voxels is None:
            raise ValueError("voxels per sample is needed to form true negatives")
        rows = self._rows_from_counts(counts, int(voxels))
        self._rows.append(rows)
        return rows

    def reset(self) -> None:
        self._rows = []

    def get_buffer(self) -> torch.Tensor:
        if not self._rows:
            raise ValueError("the data to aggregate must be PyTorch Tensor.")     # MONAI's message on an empty buffer
        return torch.cat(self._rows, dim=0).contiguous()

    def _reduce(self, f: torch.Tensor):
        lib = _lib.load()
        n, c = f.shape[:2]
        k = f[0, 0].numel()
        # "mean" of per-class scalars has shape [1], as MONAI's `torch.where(..., t_zero)` with t_zero = zeros(1) yields;
        # the reference reads it with .item() (seg:122,154)
        shape = (k,) if self.reduction == "mean" else ((c, k) if k > 1 else (c,))
        out = torch.empty(shape, dtype=torch.float32, device=f.device)
        nn_ = torch.empty(shape, dtype=torch.float32, device=f.device)
        _lib.check(lib.b200_metric_reduce(_lib.ptr(f), n, c, k, _REDUCTIONS[self.reduction], _lib.ptr(out), _lib.ptr(nn_),
                                          _lib.stream_ptr()), "b200_metric_reduce")
        return out, nn_


class DiceMetric(_CumulativeSegMetric):
    """`DiceMetric(include_background=True, reduction="mean"|"mean_batch", get_not_nans=False)` (seg:485-486): per
    (sample, class) 2|y&p|/(|y|+|p|), NaN when the class is absent from the label."""

    def _rows_from_counts(self, counts, voxels):
        lib = _lib.load()
        n, c = counts.shape[:2]
        dice = torch.empty((n, c), dtype=torch.float32, device=counts.device)
        _lib.check(lib.b200_seg_metrics(_lib.ptr(counts.contiguous()), n, c, voxels, _lib.ptr(dice), None, _lib.stream_ptr()),
                   "b200_seg_metrics")
        return dice

    def aggregate(self):
        f, not_nans = self._reduce(self.get_buffer())
        return (f, not_nans) if self.get_not_nans else f


class ConfusionMatrixMetric(_CumulativeSegMetric):
    """`ConfusionMatrixMetric(include_background=True, metric_name="precision"|"sensitivity", reduction=..., get_not_nans=False)`
    (seg:487-494).  `aggregate()` returns a list with one tensor per metric name (the reference indexes `[0]`, seg:157,160).
    compute_sample=False (MONAI's default): the (tp, fp, tn, fn) rows are reduced first, the metric is formed from the
    reduced counts; compute_sample=True: per-sample metrics are formed first, then reduced NaN-aware."""

    def __init__(self, include_background: bool = True, metric_name="hit_rate", compute_sample: bool = False,
                 reduction: str = "mean", get_not_nans: bool = False) -> None:
        super().__init__(include_background, reduction, get_not_nans)
        names = [metric_name] if isinstance(metric_name, str) else list(metric_name)
        for nme in names:
            if nme.lower() not in _CONFUSION:
                raise NotImplementedError(f"confusion-matrix metric {nme!r}: precision and sensitivity (recall) are implemented")
        self.metric_name = [nme.lower() for nme in names]
        self.compute_sample = compute_sample

    def _rows_from_counts(self, counts, voxels):
        lib = _lib.load()
        n, c = counts.shape[:2]
        cm = torch.empty((n, c, 4), dtype=torch.float32, device=counts.device)
        _lib.check(lib.b200_seg_metrics(_lib.ptr(counts.contiguous()), n, c, voxels, None, _lib.ptr(cm), _lib.stream_ptr()),
                   "b200_seg_metrics")
        return cm

    def _metric(self, cm: torch.Tensor, which: int) -> torch.Tensor:
        lib = _lib.load()
        rows = cm.numel() // 4
        out = torch.empty(cm.shape[:-1], dtype=torch.float32, device=cm.device)
        _lib.check(lib.b200_confusion_metric(_lib.ptr(cm.contiguous()), rows, which, _lib.ptr(out), _lib.stream_ptr()),
                   "b200_confusion_metric")
        return out

    def aggregate(self):
        data = self.get_buffer()
        results = []
        for nme in self.metric_name:
            which = _CONFUSION[nme]
            if self.compute_sample:
                f, not_nans = self._reduce(self._metric(data, which).contiguous())
            else:
                red, not_nans = self._reduce(data)
                f = self._metric(red, which)
            results.append((f, not_nans) if self.get_not_nans else f)
        return results
