"""Sliding-window whole-volume inference on the GPU -- `monai.inferers.sliding_window_inference` as called at
unetr_segmentation_3d.py:109 (positional, overlap 0.25), :143 and :694-695 (keyword, overlap 0.8).

Window enumeration, padding, scan interval and accumulation order follow MONAI 0.6.0; the
gather / overlap-add / divide-by-count arithmetic runs in csrc/sliding.cuh.

Multi-GPU (`rank`/`world_size`, one process per GPU): the window list of each volume is cut into contiguous chunks, and
rank r OWNS an equal share of the padded rows (first spatial axis) -- an x-slab.  Each rank predicts its windows, sends the rows of
those predictions that fall into other ranks' slabs as row-clipped pieces (NCCL send/recv: only halo rows cross NVLink, about one window depth per boundary), adds every piece of its own
slab in GLOBAL window order (bit-identical to the single-GPU loop, which is MONAI's order), normalises its slab locally, and the
uint8 mask / validation counts are combined with one small all-reduce.  The logits stay sharded unless the caller asks for them.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Callable, List, Sequence, Tuple

import torch

from . import _lib

__all__ = ["sliding_window_inference", "window_starts", "shard_windows", "slab_plan"]


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
    """Contiguous chunk of the ij-ordered window list owned by `rank` (x-slabs), balanced to within one window: the
    first `n_items % world_size` ranks take one more.  Every rank gets at least one window when n_items >= world_size."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def slab_plan(flat: Sequence[Tuple[int, int, int]], roi0: int, padded_rows: int, world_size: int):
    """Ownership and halo traffic of the slab-owned sliding window for ONE volume (pure host logic; unit-tested on the CPU).

    `flat`: window starts in MONAI order (first axis slowest).  Returns (chunks, bounds, pieces):
      chunks[r]  = range of window indices rank r predicts (contiguous, balanced: `shard_windows`);
      bounds     = world_size + 1 padded-row boundaries: rank r owns rows [bounds[r], bounds[r+1]) -- EQUAL slabs, so that every
                   rank accumulates and normalises the same number of rows (cutting at the first window of each chunk instead
                   leaves 48..128-row slabs at 8 ranks on 512 rows, and the widest one sets the time);
      pieces[d]  = the contributions to rank d's slab in global window order: (window, source rank, x_lo, x_hi) with
                   [x_lo, x_hi) = that window's rows inside d's slab.  source == d: an own window, otherwise a halo piece.
    A window reaches at most roi0 rows, i.e. into the slabs next to its own chunk's; traffic flows both ways."""
    n = len(flat)
    chunks = [shard_windows(n, r, world_size) for r in range(world_size)]
    bounds = [(padded_rows * r) // world_size for r in range(world_size)] + [padded_rows]
    pieces = [[] for _ in range(world_size)]
    for src, ch in enumerate(chunks):
        for w in ch:
            xs = flat[w][0]
            for d in range(world_size):
                lo, hi = max(xs, bounds[d]), min(xs + roi0, bounds[d + 1])
                if hi > lo:
                    pieces[d].append((w, src, lo, hi))
    for d in range(world_size):
        pieces[d].sort(key=lambda p: p[0])
    return chunks, bounds, pieces


def sliding_window_inference(inputs: torch.Tensor, roi_size, sw_batch_size: int, predictor: Callable,
                             overlap: float = 0.25, mode: str = "constant", sigma_scale=0.125,
                             padding_mode: str = "constant", cval: float = 0.0, sw_device=None, device=None,
                             *args, rank: int = 0, world_size: int = 1, process_group=None,
                             return_argmax: bool = False, labels: torch.Tensor = None, return_logits: bool = True,
                             gather_logits: bool = True, **kwargs):
    """Extras beyond MONAI's signature (all keyword-only, defaults reproduce MONAI): `rank`/`world_size`/`process_group`
    shard the windows; `return_argmax` adds the uint8 class mask; `labels` ([B,1,D,H,W] class ids) fuses the validation
    tail (seg:110-126) into the normalise pass and adds the [B,C,3] counts DiceMetric/ConfusionMatrixMetric consume
    (`metric.update_from_counts`); `return_logits=False` skips writing the 3.76 GB normalised logits when only the mask /
    counts are wanted.  Return: logits | (logits, mask) | (logits, mask, counts), `None` in place of skipped logits.
    `world_size > 1`: slab-owned accumulation (module docstring); mask and counts are complete on every rank; the logits are
    complete on every rank with `gather_logits=True` (default, MONAI semantics: one broadcast per slab) and otherwise stay
    sharded: the returned tensor is this rank's slab `[1,C,rows,H,W]` with `.row_offset` (batch 1 only, else None)."""
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
    st = _lib.stream_ptr()
    if world_size > 1:
        return _sw_slabs(lib, x, per_axis, flat, roi, size, orig, pad, chan, batch, max(1, min(int(sw_batch_size), 16)), predictor, args, kwargs,
                         float(cval), rank, world_size, process_group, return_argmax, labels, return_logits, gather_logits)
    items = [(b, *flat[w]) for b in range(batch) for w in range(num_win)]   # idx -> (idx // num_win, idx % num_win)
    mine = range(len(items))

    gin = _lib.SwGeom(chan, *orig, *pad, *size, *roi)
    sw_batch_size = max(1, min(int(sw_batch_size), 16))
    acc = None
    gout = None
    # our own UNETR replays its forward as one CUDA graph inside this loop (each prediction is accumulated at once, so the
    # graph's static output buffers may be overwritten by the next call); B200_NO_GRAPH=1 keeps the eager launches
    with _graphed(predictor, len(mine) >= 4 * sw_batch_size):
        acc, gout = _sw_loop(lib, x, items, mine, sw_batch_size, chan, roi, gin, cval, st, predictor, args, kwargs, batch, size, orig, pad)
    return _sw_finish(lib, acc, gout, batch, orig, per_axis, return_argmax, x, st, labels, return_logits)


class _graphed:
    """our own UNETR replays its forward as one CUDA graph inside the window loop (see sliding_window_inference)"""

    def __init__(self, predictor, worth_it):
        self.p = predictor
        self.on = hasattr(predictor, "_graph_forward") and not getattr(predictor, "tuple_output", True) and \
            not torch.is_grad_enabled() and not os.environ.get("B200_NO_GRAPH") and worth_it

    def __enter__(self):
        if self.on:
            self.prev = getattr(self.p, "inference_graph", False)
            self.p.inference_graph = True

    def __exit__(self, *exc):
        if self.on:
            self.p.inference_graph = self.prev


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


def _sw_finish(lib, acc, gout, batch, orig, per_axis, return_argmax, x, st, labels=None, return_logits=True):
    cout = acc.shape[1]
    want_mask = return_argmax or labels is not None
    out = torch.empty((batch, cout, *orig), dtype=torch.float32, device=x.device) if return_logits else None
    mask = torch.empty((batch, 1, *orig), dtype=torch.uint8, device=x.device) if want_mask else None
    lab, counts = _label_args(labels, x, batch, cout, orig)
    arrs = [(ctypes.c_int32 * len(a))(*a) for a in per_axis]
    _lib.check(lib.b200_sw_finalize_metric(_lib.ptr(acc), _lib.ptr(out), _lib.ptr(mask), ctypes.byref(gout), batch,
                                           arrs[0], len(arrs[0]), arrs[1], len(arrs[1]), arrs[2], len(arrs[2]),
                                           _lib.ptr(lab), _lib.ptr(counts), st), "b200_sw_finalize_metric")
    if counts is not None:
        counts.voxels = orig[0] * orig[1] * orig[2]
        return out, mask, counts
    return (out, mask) if return_argmax else out


def _label_args(labels, x, batch, cout, orig):
    if labels is None:
        return None, None
    if cout > 32:
        raise NotImplementedError("fused validation counts take at most 32 classes")
    lab = labels.to(x.device).float().contiguous()
    if lab.shape[0] != batch or tuple(lab.shape[-3:]) != tuple(orig) or lab.numel() != batch * orig[0] * orig[1] * orig[2]:
        raise ValueError(f"labels {tuple(labels.shape)} do not match the volume {(batch, 1, *orig)}")
    return lab, torch.zeros((batch, cout, 3), dtype=torch.float64, device=x.device)


# ------------------------------------------------------------------------------------------------ slab-owned multi-GPU path
class _SlabItem:
    """One volume (batch item) on one rank: predict own windows, pack the halo pieces, accumulate the slab, normalise it."""

    def __init__(self, lib, x, item, flat, roi, size, orig, pad, chan, rank, world, cval):
        self.lib, self.x, self.item, self.flat, self.roi, self.size, self.orig, self.pad = lib, x, item, flat, roi, size, orig, pad
        self.chan, self.rank, self.world, self.cval = chan, rank, world, cval
        self.chunks, self.bounds, self.pieces = slab_plan(flat, roi[0], size[0], world)
        self.mine = self.chunks[rank]
        self.x0, self.x1 = self.bounds[rank], self.bounds[rank + 1]
        self.gin = _lib.SwGeom(chan, *orig, *pad, *size, *roi)
        self.preds = None
        self.gout = None
        self.cout = None

    # The halo exchange is cut into ROUNDS so that it overlaps the prediction loop: the pieces of the windows in the j-th part of a
    # rank's chunk travel as soon as that part is predicted (NCCL runs them on its own stream); only the last round is exposed.
    ROUNDS = 4

    def round_of(self, w, src):
        ch = self.chunks[src]
        return min(self.ROUNDS - 1, (w - ch.start) * self.ROUNDS // max(len(ch), 1))

    def round_end(self, j):
        """number of own windows that must be predicted before round j can be sent"""
        n = len(self.mine)
        return n if j >= self.ROUNDS - 1 else max(i for i in range(n + 1) if i == 0 or self.round_of(self.mine.start + i - 1, self.rank) <= j)

    def predict(self, predictor, sw_batch, args, kwargs, after_call=None):
        lib, st = self.lib, _lib.stream_ptr()
        for g0 in range(self.mine.start, self.mine.stop, sw_batch):
            ws = range(g0, min(g0 + sw_batch, self.mine.stop))
            n = len(ws)
            starts = (ctypes.c_int32 * (4 * n))(*[v for w in ws for v in (self.item, *self.flat[w])])
            win = torch.empty((n, self.chan, *self.roi), dtype=torch.float32, device=self.x.device)
            _lib.check(lib.b200_sw_gather(_lib.ptr(self.x), _lib.ptr(win), ctypes.byref(self.gin), starts, n, self.cval, st), "b200_sw_gather")
            pred = predictor(win, *args, **kwargs)
            if isinstance(pred, (tuple, list)):
                raise TypeError("predictor must return a tensor (monai.networks.nets.UNETR flavour, seg:36)")
            if self.preds is None:
                self.cout = pred.shape[1]
                self.preds = torch.empty((len(self.mine), self.cout, *self.roi), dtype=torch.float32, device=self.x.device)
                self.gout = _lib.SwGeom(self.cout, *self.orig, *self.pad, *self.size, *self.roi)
            # every prediction is kept until the halo pieces of the lower ranks have arrived: a voxel's additions must happen in
            # global window order, and those pieces come first
            self.preds[g0 - self.mine.start:g0 - self.mine.start + n].copy_(pred)
            if after_call is not None:
                after_call(self, g0 - self.mine.start + n)

    def piece_elems(self, lo, hi):
        return self.cout * (hi - lo) * self.roi[1] * self.roi[2]

    def pack_sends(self, rnd=None):
        """{dst: flat fp32 buffer} of the row-clipped pieces of own windows (of exchange round `rnd`; None = all) that land in dst's
        slab, in window order."""
        lib, st, out = self.lib, _lib.stream_ptr(), {}
        for d in range(self.world):
            if d == self.rank:
                continue
            mine = [p for p in self.pieces[d] if p[1] == self.rank and (rnd is None or self.round_of(p[0], self.rank) == rnd)]
            if not mine:
                continue
            buf = torch.empty(sum(self.piece_elems(lo, hi) for _, _, lo, hi in mine), dtype=torch.float32, device=self.x.device)
            off = 0
            for w, _, lo, hi in mine:
                _lib.check(lib.b200_sw_pack_rows(_lib.ptr(self.preds[w - self.mine.start]), ctypes.c_void_p(buf.data_ptr() + 4 * off),
                                                 ctypes.byref(self.gout), lo - self.flat[w][0], hi - lo, st), "b200_sw_pack_rows")
                off += self.piece_elems(lo, hi)
            out[d] = buf
        return out

    def recv_sizes(self, cout, rnd=None):
        """{src: element count} this rank receives (in exchange round `rnd`; None = all); cout is known to every rank: same predictor"""
        out = {}
        for w, src, lo, hi in self.pieces[self.rank]:
            if src != self.rank and (rnd is None or self.round_of(w, src) == rnd):
                out[src] = out.get(src, 0) + cout * (hi - lo) * self.roi[1] * self.roi[2]
        return out

    def accumulate(self, recv):
        """recv: {src: flat buffer} (one exchange) or {(src, round): flat buffer}.  Adds all pieces of the slab in global window order;
        returns the slab accumulator."""
        lib, st = self.lib, _lib.stream_ptr()
        nrows = self.x1 - self.x0
        acc = torch.zeros((self.cout, max(nrows, 1), self.size[1], self.size[2]), dtype=torch.float32, device=self.x.device)
        if nrows <= 0:
            return acc
        offs = {src: 0 for src in recv}
        todo = []
        for w, src, lo, hi in self.pieces[self.rank]:
            s0, s1, s2 = self.flat[w]
            if src == self.rank:
                ptr, nx, xbase = self.preds[w - self.mine.start].data_ptr(), self.roi[0], s0
            else:
                key = (src, self.round_of(w, src)) if (src, self.round_of(w, src)) in recv else src
                ptr, nx, xbase = recv[key].data_ptr() + 4 * offs[key], hi - lo, lo
                offs[key] += self.cout * (hi - lo) * self.roi[1] * self.roi[2]
            todo.append((ptr, (s0, s1, s2, lo, hi, nx, xbase)))
        for k0 in range(0, len(todo), 16):
            grp = todo[k0:k0 + 16]
            ptrs = (ctypes.c_void_p * len(grp))(*[g[0] for g in grp])
            desc = (ctypes.c_int32 * (7 * len(grp)))(*[v for g in grp for v in g[1]])
            _lib.check(lib.b200_sw_accumulate_slab(_lib.ptr(acc), ctypes.byref(self.gout), ptrs, desc, len(grp), self.x0, nrows, st),
                       "b200_sw_accumulate_slab")
        return acc

    def rows(self):
        """un-padded rows [d0, d1) of the volume this rank owns"""
        d0 = min(max(self.x0 - self.pad[0], 0), self.orig[0])
        d1 = min(max(self.x1 - self.pad[0], 0), self.orig[0])
        return d0, d1

    def finalize(self, acc, per_axis, out, out_d0, out_rows, mask, lab, counts):
        d0, d1 = self.rows()
        arrs = [(ctypes.c_int32 * len(a))(*a) for a in per_axis]
        _lib.check(self.lib.b200_sw_finalize_slab(_lib.ptr(acc), _lib.ptr(out), _lib.ptr(mask), ctypes.byref(self.gout), self.item,
                                                  arrs[0], len(arrs[0]), arrs[1], len(arrs[1]), arrs[2], len(arrs[2]),
                                                  _lib.ptr(lab), _lib.ptr(counts), d0, d1 - d0, self.x0, max(self.x1 - self.x0, 1),
                                                  out_d0, out_rows, _lib.stream_ptr()), "b200_sw_finalize_slab")


def _exchange_nccl(sends, recv_sizes, device, group, wait=True):
    """sends: {dst: buffer}; recv_sizes: {src: elements} -> {src: buffer}.  One batched NCCL send/recv round.  wait=False returns
    (recv, works): the caller waits on the works before it reads the buffers (and keeps `sends` alive until then)."""
    import torch.distributed as dist
    recv = {src: torch.empty(n, dtype=torch.float32, device=device) for src, n in recv_sizes.items()}
    ops = [dist.P2POp(dist.isend, buf, dst, group) for dst, buf in sorted(sends.items())] + \
          [dist.P2POp(dist.irecv, buf, src, group) for src, buf in sorted(recv.items())]
    works = dist.batch_isend_irecv(ops) if ops else []
    if not wait:
        return recv, works
    for w in works:
        w.wait()
    return recv


def _sw_slabs(lib, x, per_axis, flat, roi, size, orig, pad, chan, batch, sw_batch, predictor, args, kwargs, cval, rank, world, group,
              return_argmax, labels, return_logits, gather_logits, exchange=None):
    import torch.distributed as dist
    if len(flat) < world:
        raise RuntimeError("fewer windows than ranks; use fewer ranks")
    want_mask = return_argmax or labels is not None
    mask = torch.zeros((batch, 1, *orig), dtype=torch.uint8, device=x.device) if want_mask else None
    lab = counts = None
    full = None
    slab_out = None
    items = [_SlabItem(lib, x, b, flat, roi, size, orig, pad, chan, rank, world, cval) for b in range(batch)]
    timing = os.environ.get("B200_SW_TIMING")          # tuning aid: CUDA-event time of each phase, printed by every rank to stderr
    marks = []

    def mark(tag):
        if timing:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((tag, e))
    mark("start")
    # halo pieces leave in rounds while the prediction loop is still running (NCCL on its own stream); `exchange` (tests) = one blocking round
    inflight = {id(it): {"recv": {}, "works": [], "keep": [], "next": 0} for it in items}

    def after_call(it, done_local):
        if exchange is not None:
            return
        fl = inflight[id(it)]
        while fl["next"] < it.ROUNDS and done_local >= it.round_end(fl["next"]):
            j = fl["next"]
            sends = it.pack_sends(j)
            recv, works = _exchange_nccl(sends, it.recv_sizes(it.cout, j), x.device, group, wait=False)
            fl["recv"].update({(src, j): buf for src, buf in recv.items()})
            fl["works"] += works
            fl["keep"].append(sends)
            fl["next"] = j + 1
    with _graphed(predictor, len(items[0].mine) * batch >= 4 * sw_batch):
        for it in items:
            it.predict(predictor, sw_batch, args, kwargs, after_call)
    mark("predict+pack+send")
    cout = items[0].cout
    lab, counts = _label_args(labels, x, batch, cout, orig)
    for it in items:
        if exchange is not None:
            recv = exchange(it.pack_sends(), it.recv_sizes(cout))
        else:
            fl = inflight[id(it)]
            for wk in fl["works"]:
                wk.wait()
            recv = fl["recv"]
        mark("exchange tail")
        acc = it.accumulate(recv)
        mark("accumulate")
        d0, d1 = it.rows()
        out = None
        if return_logits and gather_logits:
            if full is None:
                full = torch.empty((batch, cout, *orig), dtype=torch.float32, device=x.device)
            out, out_d0, out_rows = full[it.item], 0, orig[0]
        elif return_logits and batch == 1:
            slab_out = torch.empty((1, cout, max(d1 - d0, 0), orig[1], orig[2]), dtype=torch.float32, device=x.device)
            slab_out.row_offset = d0
            out, out_d0, out_rows = slab_out, d0, max(d1 - d0, 1)
        else:
            out_d0, out_rows = 0, 1
        it.finalize(acc, per_axis, out, out_d0, out_rows, mask, lab, counts)
        it.preds = None
    # combine: every un-padded row has exactly one owner, the others hold zeros
    if mask is not None and dist.is_initialized():
        dist.all_reduce(mask, group=group)
    if counts is not None and dist.is_initialized():
        dist.all_reduce(counts, group=group)
    if full is not None and dist.is_initialized():
        works = []
        for it in items:
            for r in range(world):
                lo = min(max(it.bounds[r] - pad[0], 0), orig[0]); hi = min(max(it.bounds[r + 1] - pad[0], 0), orig[0])
                if hi > lo:      # rows of one owner are a strided view per channel: broadcast channel planes as contiguous chunks
                    for c in range(cout):
                        works.append(dist.broadcast(full[it.item, c, lo:hi], src=dist.get_global_rank(group, r) if group is not None else r,
                                                    group=group, async_op=True))
        for w in works:
            w.wait()
    mark("finalize+combine")
    if timing:
        import sys
        torch.cuda.synchronize()
        print(f"[sliding window rank {rank}] " + "  ".join(f"{b[0]} {a[1].elapsed_time(b[1]):.2f} ms" for a, b in zip(marks, marks[1:])) +
              f"  (windows {len(items[0].mine)}, slab rows {items[0].x1 - items[0].x0})", file=sys.stderr, flush=True)
    logits = full if full is not None else slab_out
    if counts is not None:
        counts.voxels = orig[0] * orig[1] * orig[2]
        return logits, mask, counts
    return (logits, mask) if return_argmax else logits
