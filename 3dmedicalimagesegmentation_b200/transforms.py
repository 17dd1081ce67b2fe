"""GPU-side crop sampling and augmentation -- the per-iteration tail of the reference's training transforms on
volumes that are resident in HBM, under MONAI's names and dictionary-transform call convention:

    RandCropByPosNegLabeld(keys, label_key, spatial_size, pos, neg, num_samples, image_key, image_threshold)   seg:341-350
    RandFlipd(keys, spatial_axis, prob) x3, RandRotate90d(keys, prob, max_k), RandShiftIntensityd(keys, offsets, prob)   seg:351-375
    RandSpatialCropSamplesd(keys, roi_size, random_size=False, num_samples)                                     rank:365-369
    ConvertToMultiChannelBasedOnBratsClassesd(keys)                                                             seg:65-93
    Compose([...])  -- maps transforms over the list a multi-sample crop returns, `set_random_state(seed)` as MONAI's

Random draws are made on the host with numpy `RandomState` streams in the order MONAI 0.6.0's `randomize()` methods make them (one
stream per transform, derived from the Compose seed), so a seeded pipeline is reproducible and comparable with a restatement of the
reference (oracle/transforms_oracle.py).  Everything that touches voxels runs in csrc/augment.cuh: the foreground / background
sets are per-block counts + a prefix scan (never index lists), and `Compose` FUSES a crop sampler followed by flips / rot90 /
intensity shift (and the BraTS conversion) into ONE gather launch per batch of crops: the composed signed axis permutation is
applied to the output coordinate, so each output voxel is read and written once.  Tensors are `[C, D, H, W]` fp32 on the GPU.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Sequence

import numpy as np
import torch

from . import _lib

__all__ = ["Compose", "RandCropByPosNegLabeld", "RandSpatialCropSamplesd", "RandFlipd", "RandRotate90d", "RandShiftIntensityd",
           "ConvertToMultiChannelBasedOnBratsClassesd", "AxisMap"]

MAX_SEED = np.iinfo(np.uint32).max + 1


class AxisMap:
    """A composition of axis flips and 90-degree rotations of a cubic crop as a signed axis permutation:
    crop coordinate b of the voxel that lands at output coordinate `o` is `o[src[b]]`, reversed when `rev[b]`."""

    def __init__(self):
        self.src, self.rev, self.shift = [0, 1, 2], [False, False, False], 0.0

    def then(self, t_src: Sequence[int], t_rev: Sequence[bool]) -> "AxisMap":
        """apply a further transform whose input coordinate c is `new[t_src[c]]` (reversed when t_rev[c])"""
        out = AxisMap()
        out.shift = self.shift
        out.src = [t_src[self.src[b]] for b in range(3)]
        out.rev = [self.rev[b] != t_rev[self.src[b]] for b in range(3)]
        return out

    def flip(self, axis: int) -> "AxisMap":
        rev = [False] * 3
        rev[axis] = True
        return self.then([0, 1, 2], rev)

    def rot90(self, k: int, axes=(0, 1)) -> "AxisMap":
        """numpy.rot90(m, k, axes): k=1: out[i,j] = m[j, N-1-i]; k=2: m[N-1-i, N-1-j]; k=3: m[N-1-j, i]"""
        a0, a1 = axes
        m = self
        for _ in range(k % 4):
            src, rev = [0, 1, 2], [False] * 3
            src[a0], src[a1] = a1, a0
            rev[a1] = True
            m = m.then(src, rev)
        return m

    def kernel_form(self):
        perm, flip = [0, 0, 0], [0, 0, 0]
        for b in range(3):
            perm[self.src[b]] = b
            flip[self.src[b]] = int(self.rev[b])
        return perm, flip

    def apply_numpy(self, crop: np.ndarray) -> np.ndarray:
        """the same gather with numpy (host logic check): crop [..., r, r, r]"""
        perm, flip = self.kernel_form()
        out = np.transpose(crop, list(range(crop.ndim - 3)) + [crop.ndim - 3 + p for p in perm])
        for a in range(3):
            if flip[a]:
                out = np.flip(out, axis=crop.ndim - 3 + a)
        return out + np.float32(self.shift) if self.shift else out


class _Randomizable:
    def __init__(self):
        self.R = np.random.RandomState()

    def set_random_state(self, seed=None, state=None):
        if seed is not None:
            self.R = np.random.RandomState(int(seed) % MAX_SEED)
        elif state is not None:
            self.R = state
        else:
            self.R = np.random.RandomState()
        return self


def _check(t: torch.Tensor, what: str):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dim() == 4):
        raise TypeError(f"{what} must be a [C, D, H, W] CUDA tensor (GPU-side transforms; there is no CPU path)")
    _lib.require_device(t)
    return t.contiguous().float()


def _gather(image, label, starts_dev, maps: List[AxisMap], roi, brats=False):
    """n augmented crops of (image, label) in one launch.  image / label may be None."""
    lib = _lib.load()
    ref = image if image is not None else label
    n, (d, h, w) = len(maps), ref.shape[1:]
    arr = (_lib.AugMap * n)()
    for i, m in enumerate(maps):
        perm, flip = m.kernel_form()
        arr[i].perm[:] = perm
        arr[i].flip[:] = flip
        arr[i].shift = float(np.float32(m.shift))
    ci = image.shape[0] if image is not None else 0
    cl = label.shape[0] if label is not None else 0
    out_i = torch.empty((n, ci, *roi), dtype=torch.float32, device=ref.device) if image is not None else None
    out_l = torch.empty((n, 4 if brats else cl, *roi), dtype=torch.float32, device=ref.device) if label is not None else None
    _lib.check(lib.b200_aug_crop(_lib.ptr(image), ci, _lib.ptr(label), cl, d, h, w, _lib.ptr(starts_dev), arr, n, *roi, int(brats),
                                 _lib.ptr(out_i), _lib.ptr(out_l), _lib.stream_ptr()), "b200_aug_crop")
    return out_i, out_l


class ConvertToMultiChannelBasedOnBratsClassesd:
    """seg:65-93: label map {0,1,2,3} -> float32 [4, ...] = (background, TC = 2|3, WT = 1|2|3, ET = 3)."""

    def __init__(self, keys):
        self.keys = [keys] if isinstance(keys, str) else list(keys)

    def __call__(self, data: Dict):
        d = dict(data)
        for key in self.keys:
            lab = d[key]
            if lab.dim() == 3:
                lab = lab[None]
            lab = _check(lab, key)
            if lab.shape[0] != 1:
                raise ValueError("ConvertToMultiChannelBasedOnBratsClassesd takes a single-channel label map")
            zeros = torch.zeros(3, dtype=torch.int32, device=lab.device)
            _, out = _gather(None, lab, zeros, [AxisMap()], tuple(lab.shape[1:]), brats=True)
            d[key] = out[0]
        return d


class _PointTransform(_Randomizable):
    """flip / rot90 / shift: `draw()` consumes the transform's random stream the way MONAI's randomize() does and returns a function
    AxisMap -> AxisMap; `image_only`: the map's shift applies to the image key only (always true in the kernel)."""

    def __init__(self, keys, prob):
        super().__init__()
        self.keys = [keys] if isinstance(keys, str) else list(keys)
        self.prob = float(prob)

    def __call__(self, data: Dict):
        fn = self.draw()
        d = dict(data)
        first = _check(d[self.keys[0]], self.keys[0])
        m = fn(AxisMap())
        zeros = torch.zeros(3, dtype=torch.int32, device=first.device)
        for key in self.keys:
            t = _check(d[key], key)
            mk = m
            if key != self.keys[0] and m.shift:
                mk = AxisMap(); mk.src, mk.rev = m.src, m.rev      # noqa: E702  (shift applies to the first key = the image)
            out, _ = _gather(t, None, zeros, [mk], tuple(t.shape[1:]))
            d[key] = out[0]
        return d


class RandFlipd(_PointTransform):
    def __init__(self, keys, prob: float = 0.1, spatial_axis=None):
        super().__init__(keys, prob)
        axes = [0, 1, 2] if spatial_axis is None else ([spatial_axis] if isinstance(spatial_axis, int) else list(spatial_axis))
        self.axes = [int(a) for a in axes]

    def draw(self):
        do = self.R.rand() < self.prob                                 # RandomizableTransform.randomize
        axes = self.axes

        def fn(m: AxisMap):
            if do:
                for a in axes:
                    m = m.flip(a)
            return m
        return fn


class RandRotate90d(_PointTransform):
    def __init__(self, keys, prob: float = 0.1, max_k: int = 3, spatial_axes=(0, 1)):
        super().__init__(keys, prob)
        self.max_k, self.axes = int(max_k), tuple(spatial_axes)

    def draw(self):
        k = self.R.randint(self.max_k) + 1                             # RandRotate90d.randomize: k first, then the coin
        do = self.R.rand() < self.prob
        axes = self.axes
        return lambda m: m.rot90(k, axes) if do else m


class RandShiftIntensityd(_PointTransform):
    def __init__(self, keys, offsets, prob: float = 0.1):
        super().__init__(keys, prob)
        self.offsets = (-abs(float(offsets)), abs(float(offsets))) if np.isscalar(offsets) else (min(offsets), max(offsets))

    def draw(self):
        off = self.R.uniform(low=self.offsets[0], high=self.offsets[1])  # RandShiftIntensityd.randomize: offset first, then the coin
        do = self.R.rand() < self.prob

        def fn(m: AxisMap):
            if do:
                m = m.then([0, 1, 2], [False] * 3)
                m.shift = float(np.float32(np.float32(m.shift) + np.float32(off)))
            return m
        return fn


class _CropSampler(_Randomizable):
    num_samples = 1

    def starts(self, data) -> torch.Tensor:  # device int32 [n, 3]
        raise NotImplementedError

    def __call__(self, data: Dict) -> List[Dict]:
        return _fused_crop(self, [], data)


class RandSpatialCropSamplesd(_CropSampler):
    """rank:365-369: `num_samples` crops of `roi_size` at uniformly random corners (random_size=False)."""

    def __init__(self, keys, roi_size, num_samples: int, random_center: bool = True, random_size: bool = False):
        super().__init__()
        if random_size or not random_center:
            raise NotImplementedError("RandSpatialCropSamplesd is implemented as the reference uses it: random_center=True, random_size=False")
        self.keys = [keys] if isinstance(keys, str) else list(keys)
        self.roi = tuple(int(r) for r in (roi_size if isinstance(roi_size, Sequence) else (roi_size,) * 3))
        self.num_samples = int(num_samples)
        self.image_key, self.label_key = self.keys[0], (self.keys[1] if len(self.keys) > 1 else None)

    def roi_for(self, shape):
        return tuple(r if r > 0 else s for r, s in zip(self.roi, shape))          # fall_back_tuple

    def starts(self, data):
        img = data[self.image_key]
        shape = tuple(img.shape[1:])
        roi = tuple(min(r, s) for r, s in zip(self.roi_for(shape), shape))        # get_valid_patch_size
        out = []
        for _ in range(self.num_samples):                                          # get_random_patch: one randint per axis that has room
            out.append([int(self.R.randint(0, ms - ps + 1)) if ms > ps else 0 for ms, ps in zip(shape, roi)])
        return torch.tensor(out, dtype=torch.int32).to(img.device, non_blocking=True), roi


class RandCropByPosNegLabeld(_CropSampler):
    """seg:341-350.  The foreground / background sets of `map_binary_to_indices` are indexed once per (label, image) pair and cached
    on the sampler (they do not change between epochs); each call then costs the host 2 draws per crop and the GPU two launches."""

    def __init__(self, keys, label_key, spatial_size, pos: float = 1.0, neg: float = 1.0, num_samples: int = 1, image_key=None,
                 image_threshold: float = 0.0):
        super().__init__()
        if pos < 0 or neg < 0:
            raise ValueError(f"pos and neg must be nonnegative, got pos={pos} neg={neg}.")
        if pos + neg == 0:
            raise ValueError("Incompatible values: pos=0 and neg=0.")
        self.keys = [keys] if isinstance(keys, str) else list(keys)
        self.label_key, self.image_key = label_key, image_key
        self.roi = tuple(int(r) for r in (spatial_size if isinstance(spatial_size, Sequence) else (spatial_size,) * 3))
        self.pos_ratio = pos / (pos + neg)
        self.num_samples, self.thr = int(num_samples), float(image_threshold)
        self._index = {}

    def roi_for(self, shape):
        return tuple(r if r > 0 else s for r, s in zip(self.roi, shape))

    def _indexed(self, label, image):
        key = (label.data_ptr(), label._version, tuple(label.shape), image.data_ptr() if image is not None else 0,
               image._version if image is not None else 0)
        ent = self._index.get(key)
        if ent is None:
            lib = _lib.load()
            v = label[0].numel()
            prefix = torch.empty((2, lib.b200_aug_blocks(v)), dtype=torch.int32, device=label.device)
            totals = torch.empty(2, dtype=torch.int64, device=label.device)
            _lib.check(lib.b200_aug_index(_lib.ptr(label), label.shape[0], _lib.ptr(image), image.shape[0] if image is not None else 0,
                                          self.thr, v, _lib.ptr(prefix), _lib.ptr(totals), _lib.stream_ptr()), "b200_aug_index")
            nfg, nbg = (int(t) for t in totals.tolist())                     # the one host round trip per volume
            if len(self._index) >= 64:
                self._index.clear()
            ent = self._index[key] = (prefix, nfg, nbg)
        return ent

    def starts(self, data):
        lib = _lib.load()
        label = _check(data[self.label_key], self.label_key)
        image = _check(data[self.image_key], self.image_key) if self.image_key else None
        prefix, nfg, nbg = self._indexed(label, image)
        roi = self.roi_for(tuple(label.shape[1:]))
        if any(r > s for r, s in zip(roi, label.shape[1:])):
            raise ValueError("The size of the proposed random crop ROI is larger than the image size.")
        if nfg == 0 and nbg == 0:
            raise ValueError("No sampling location available.")
        pos_ratio = self.pos_ratio if (nfg and nbg) else (0.0 if nfg == 0 else 1.0)
        picks = (ctypes.c_int64 * (2 * self.num_samples))()
        for i in range(self.num_samples):                                     # generate_pos_neg_label_crop_centers: coin, then index
            use_fg = self.R.rand() < pos_ratio
            picks[2 * i] = int(use_fg)
            picks[2 * i + 1] = int(self.R.randint(nfg if use_fg else nbg))
        starts = torch.empty((self.num_samples, 3), dtype=torch.int32, device=label.device)
        for i0 in range(0, self.num_samples, 16):
            n = min(16, self.num_samples - i0)
            sub = (ctypes.c_int64 * (2 * n))(*picks[2 * i0:2 * (i0 + n)])
            _lib.check(lib.b200_aug_pick_centers(_lib.ptr(label), label.shape[0], _lib.ptr(image), image.shape[0] if image is not None else 0,
                                                 self.thr, *label.shape[1:], _lib.ptr(prefix), sub, n, *roi,
                                                 ctypes.c_void_p(starts.data_ptr() + 12 * i0), _lib.stream_ptr()), "b200_aug_pick_centers")
        return starts, roi


def _fused_crop(sampler: _CropSampler, points: List[_PointTransform], data: Dict, brats_key=None) -> List[Dict]:
    """crop sampler + following flips / rot90 / shift as ONE gather launch per 16 crops.  Draw order = MONAI's: the sampler first, then
    each point transform over the samples 0..n-1 in turn (Compose maps a transform over the list before moving to the next)."""
    starts, roi = sampler.starts(data)
    n = sampler.num_samples
    maps = [AxisMap() for _ in range(n)]
    for t in points:
        for i in range(n):
            maps[i] = t.draw()(maps[i])
    if any(m.src != [0, 1, 2] for m in maps) and len({roi[a] for a in (0, 1)}) != 1:
        raise NotImplementedError("rot90 of a crop whose rotated plane is not square")
    image_key = getattr(sampler, "image_key", None) or sampler.keys[0]
    label_key = getattr(sampler, "label_key", None)
    image = _check(data[image_key], image_key)
    label = _check(data[label_key], label_key) if label_key and label_key in data else None
    outs_i, outs_l = [], []
    for i0 in range(0, n, 16):
        m = min(16, n - i0)
        oi, ol = _gather(image, label, starts[i0:i0 + m].contiguous(), maps[i0:i0 + m], roi, brats=brats_key is not None and brats_key == label_key)
        outs_i.append(oi); outs_l.append(ol)      # noqa: E702
    oi = torch.cat(outs_i) if len(outs_i) > 1 else outs_i[0]
    ol = (torch.cat(outs_l) if len(outs_l) > 1 else outs_l[0]) if label is not None else None
    out = []
    for i in range(n):
        d = dict(data)
        d[image_key] = oi[i]
        if ol is not None:
            d[label_key] = ol[i]
        out.append(d)
    return out


class Compose(_Randomizable):
    """monai.transforms.Compose for the GPU-side transforms above.  A crop sampler followed by point transforms (and an optional
    ConvertToMultiChannelBasedOnBratsClassesd on the label right before the sampler) runs as one fused gather; other callables are
    applied in order, mapped over lists."""

    def __init__(self, transforms):
        super().__init__()
        self.transforms = list(transforms)

    def set_random_state(self, seed=None, state=None):
        super().set_random_state(seed, state)
        for t in self.transforms:
            if isinstance(t, _Randomizable):
                t.set_random_state(seed=self.R.randint(MAX_SEED, dtype="uint32"))
        return self

    def __call__(self, data):
        ts = self.transforms
        i = 0
        while i < len(ts):
            t = ts[i]
            brats_key, j = None, i
            if isinstance(t, ConvertToMultiChannelBasedOnBratsClassesd) and i + 1 < len(ts) and isinstance(ts[i + 1], RandSpatialCropSamplesd) \
                    and not isinstance(data, list) and len(t.keys) == 1 and t.keys[0] == ts[i + 1].label_key \
                    and data[t.keys[0]].dim() == 4 and data[t.keys[0]].shape[0] == 1:
                brats_key, j = t.keys[0], i + 1       # the conversion commutes with a label-independent crop: fuse it into the gather
            if isinstance(ts[j], _CropSampler) and not isinstance(data, list):
                k = j + 1
                while k < len(ts) and isinstance(ts[k], _PointTransform) and _compatible(ts[j], ts[k]):
                    k += 1
                data = _fused_crop(ts[j], ts[j + 1:k], data, brats_key)
                i = k
                continue
            data = [t(d) for d in data] if isinstance(data, list) else t(data)
            i += 1
        return data


def _compatible(sampler, point) -> bool:
    """a point transform can ride on the sampler's gather when it acts on the sampler's keys: geometry on all of them, shift on the image"""
    image_key = getattr(sampler, "image_key", None) or sampler.keys[0]
    label_key = getattr(sampler, "label_key", None)
    keys = set(k for k in (image_key, label_key) if k)
    if isinstance(point, RandShiftIntensityd):
        return point.keys == [image_key]
    return set(point.keys) == keys
