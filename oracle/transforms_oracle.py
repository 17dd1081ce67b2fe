"""TEST INFRASTRUCTURE -- numpy restatement of the MONAI 0.6.0 dictionary transforms the reference's training pipelines end with
(SURVEY 8f N4): RandCropByPosNegLabeld / RandFlipd / RandRotate90d / RandShiftIntensityd (unetr_segmentation_3d.py:341-375),
RandSpatialCropSamplesd (unetr_ranking_pretraining_3d.py:365-369) and ConvertToMultiChannelBasedOnBratsClassesd
(unetr_segmentation_3d.py:65-93, which is in the reference repository itself and restated line by line).

PARITY UNPINNED for the MONAI classes: monai==0.6.0 is not vendored and cannot be installed here (no network); the restatement
follows monai/transforms/{croppad,spatial,intensity}/dictionary.py and monai/transforms/utils.py of that release: the order of the
random draws inside each `randomize()`, `map_binary_to_indices`, `generate_pos_neg_label_crop_centers`, `correct_crop_centers`,
`get_random_patch`, and Compose's per-transform seeding.  Only tests/ may import this module; the product (transforms.py + csrc/augment.cuh)
never does.
"""
import numpy as np

MAX_SEED = np.iinfo(np.uint32).max + 1


class Randomizable:
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


def map_binary_to_indices(label, image=None, image_threshold=0.0):
    label_flat = np.any(label, axis=0).ravel()
    fg = np.nonzero(label_flat)[0]
    if image is not None:
        img_flat = np.any(image > image_threshold, axis=0).ravel()
        bg = np.nonzero(np.logical_and(img_flat, ~label_flat))[0]
    else:
        bg = np.nonzero(~label_flat)[0]
    return fg, bg


def correct_crop_centers(centers, spatial_size, label_spatial_shape):
    spatial_size = np.asarray(spatial_size)
    if not (np.subtract(label_spatial_shape, spatial_size) >= 0).all():
        raise ValueError("The size of the proposed random crop ROI is larger than the image size.")
    valid_start = np.floor_divide(spatial_size, 2)
    valid_end = np.subtract(np.asarray(label_spatial_shape) + np.array(1), spatial_size / np.array(2)).astype(np.uint16)
    for i, valid_s in enumerate(valid_start):
        if valid_s == valid_end[i]:
            valid_end[i] += 1
    for i, c in enumerate(centers):
        center_i = c
        if c < valid_start[i]:
            center_i = valid_start[i]
        if c >= valid_end[i]:
            center_i = valid_end[i] - 1
        centers[i] = center_i
    return centers


def generate_pos_neg_label_crop_centers(spatial_size, num_samples, pos_ratio, label_spatial_shape, fg_indices, bg_indices, rand_state):
    centers = []
    if fg_indices.size == 0 and bg_indices.size == 0:
        raise ValueError("No sampling location available.")
    if fg_indices.size == 0 or bg_indices.size == 0:
        pos_ratio = 0 if fg_indices.size == 0 else 1
    for _ in range(num_samples):
        indices_to_use = fg_indices if rand_state.rand() < pos_ratio else bg_indices
        random_int = rand_state.randint(len(indices_to_use))
        center = np.unravel_index(indices_to_use[random_int], label_spatial_shape)
        centers.append(correct_crop_centers(list(center), spatial_size, label_spatial_shape))
    return centers


def spatial_crop_center(img, center, size):
    """SpatialCrop(roi_center, roi_size)"""
    start = np.maximum(np.asarray(center) - np.floor_divide(np.asarray(size), 2), 0)
    end = np.maximum(start + np.asarray(size), start)
    sl = (slice(None),) + tuple(slice(int(s), int(e)) for s, e in zip(start, end))
    return img[sl]


class RandCropByPosNegLabeld(Randomizable):
    def __init__(self, keys, label_key, spatial_size, pos=1.0, neg=1.0, num_samples=1, image_key=None, image_threshold=0.0):
        super().__init__()
        self.keys, self.label_key, self.image_key = list(keys), label_key, image_key
        self.spatial_size, self.pos_ratio, self.num_samples, self.thr = tuple(spatial_size), pos / (pos + neg), num_samples, image_threshold

    def __call__(self, data):
        d = dict(data)
        label = d[self.label_key]
        image = d[self.image_key] if self.image_key else None
        fg, bg = map_binary_to_indices(label, image, self.thr)
        centers = generate_pos_neg_label_crop_centers(self.spatial_size, self.num_samples, self.pos_ratio, label.shape[1:], fg, bg, self.R)
        out = []
        for c in centers:
            r = dict(d)
            for k in self.keys:
                r[k] = spatial_crop_center(d[k], c, self.spatial_size)
            out.append(r)
        return out


class RandSpatialCropSamplesd(Randomizable):
    def __init__(self, keys, roi_size, num_samples, random_size=False):
        super().__init__()
        assert not random_size
        self.keys, self.roi, self.num_samples = list(keys), tuple(roi_size), num_samples

    def __call__(self, data):
        d = dict(data)
        out = []
        for _ in range(self.num_samples):
            shape = d[self.keys[0]].shape[1:]
            size = tuple(min(r, s) if r > 0 else s for r, s in zip(self.roi, shape))
            corner = tuple(self.R.randint(0, ms - ps + 1) if ms > ps else 0 for ms, ps in zip(shape, size))       # get_random_patch
            sl = (slice(None),) + tuple(slice(c, c + p) for c, p in zip(corner, size))
            r = dict(d)
            for k in self.keys:
                r[k] = d[k][sl]
            out.append(r)
        return out


class RandFlipd(Randomizable):
    def __init__(self, keys, prob=0.1, spatial_axis=None):
        super().__init__()
        self.keys, self.prob, self.axis = list(keys), prob, spatial_axis

    def __call__(self, data):
        d = dict(data)
        do = self.R.rand() < self.prob
        if do:
            axes = [0, 1, 2] if self.axis is None else ([self.axis] if isinstance(self.axis, int) else list(self.axis))
            for k in self.keys:
                d[k] = np.ascontiguousarray(np.flip(d[k], [a + 1 for a in axes]))
        return d


class RandRotate90d(Randomizable):
    def __init__(self, keys, prob=0.1, max_k=3, spatial_axes=(0, 1)):
        super().__init__()
        self.keys, self.prob, self.max_k, self.axes = list(keys), prob, max_k, spatial_axes

    def __call__(self, data):
        d = dict(data)
        k = self.R.randint(self.max_k) + 1
        do = self.R.rand() < self.prob
        if do:
            for key in self.keys:
                d[key] = np.ascontiguousarray(np.rot90(d[key], k, [a + 1 for a in self.axes]))
        return d


class RandShiftIntensityd(Randomizable):
    def __init__(self, keys, offsets, prob=0.1):
        super().__init__()
        self.keys, self.prob = list(keys), prob
        self.offsets = (-abs(offsets), abs(offsets)) if np.isscalar(offsets) else (min(offsets), max(offsets))

    def __call__(self, data):
        d = dict(data)
        offset = self.R.uniform(low=self.offsets[0], high=self.offsets[1])
        do = self.R.rand() < self.prob
        if do:
            for k in self.keys:
                d[k] = np.asarray(d[k] + np.float32(offset), dtype=d[k].dtype)       # ShiftIntensity: img + offset in the image dtype
        return d


class ConvertToMultiChannelBasedOnBratsClassesd:
    """unetr_segmentation_3d.py:65-93"""

    def __init__(self, keys):
        self.keys = list(keys)

    def __call__(self, data):
        d = dict(data)
        for key in self.keys:
            x = d[key]
            if x.ndim == 4:
                x = x[0]
            result = [x == 0, np.logical_or(x == 2, x == 3), np.logical_or(np.logical_or(x == 2, x == 3), x == 1), x == 3]
            d[key] = np.stack(result, axis=0).astype(np.float32)
        return d


class Compose(Randomizable):
    def __init__(self, transforms):
        super().__init__()
        self.transforms = list(transforms)

    def set_random_state(self, seed=None, state=None):
        super().set_random_state(seed, state)
        for t in self.transforms:
            if isinstance(t, Randomizable):
                t.set_random_state(seed=self.R.randint(MAX_SEED, dtype="uint32"))
        return self

    def __call__(self, data):
        for t in self.transforms:
            data = [t(d) for d in data] if isinstance(data, list) else t(data)
        return data
