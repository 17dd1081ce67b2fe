"""SURVEY 8f N4 -- GPU-side crop sampling / augmentation.  CPU part: the host logic (axis-map composition, draw order) against numpy
and closed-form facts about the restated MONAI helpers.  GPU part (-m gpu): seeded pipelines through the C ABI against the numpy
oracle restatement, bit-exact (pure data movement + one fp32 add)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import transforms_oracle as TO


@pytest.fixture(scope="module")
def T():
    return importlib.import_module("3dmedicalimagesegmentation_b200.transforms")


def test_axis_map_composition_equals_numpy_chain(T):
    rng = np.random.RandomState(0)
    crop = rng.rand(2, 6, 6, 6).astype(np.float32)
    for trial in range(200):
        m, want = T.AxisMap(), crop
        for _ in range(rng.randint(1, 6)):
            if rng.rand() < 0.5:
                a = int(rng.randint(3))
                m, want = m.flip(a), np.flip(want, a + 1)
            else:
                k = int(rng.randint(1, 4))
                axes = [(0, 1), (1, 2), (0, 2)][rng.randint(3)]
                m, want = m.rot90(k, axes), np.rot90(want, k, [axes[0] + 1, axes[1] + 1])
        assert np.array_equal(m.apply_numpy(crop), want), trial


def test_oracle_known_answers():
    # correct_crop_centers: centres are clamped so that the crop fits; even and odd roi
    assert TO.correct_crop_centers([0, 199, 100], (96, 96, 96), (200, 200, 200)) == [48, 152, 100]
    assert TO.correct_crop_centers([0, 199, 100], (97, 97, 97), (200, 200, 200)) == [48, 151, 100]
    assert TO.correct_crop_centers([5, 5, 5], (16, 16, 16), (16, 16, 16)) == [8, 8, 8]
    # map_binary_to_indices: background = image above threshold and not labelled
    lab = np.zeros((1, 2, 2, 2), np.float32); lab[0, 0, 0, 1] = 2
    img = np.zeros((1, 2, 2, 2), np.float32); img[0, 1] = 1; img[0, 0, 0, 1] = 1
    fg, bg = TO.map_binary_to_indices(lab, img, 0.0)
    assert fg.tolist() == [1] and bg.tolist() == [4, 5, 6, 7]
    # BraTS conversion (seg:65-93)
    x = np.array([[[0, 1], [2, 3]]], np.float32)[None]
    out = TO.ConvertToMultiChannelBasedOnBratsClassesd(["label"])({"label": x})["label"]
    assert out.shape == (4, 1, 2, 2)
    assert out[:, 0].reshape(4, 4).tolist() == [[1, 0, 0, 0], [0, 0, 1, 1], [0, 1, 1, 1], [0, 0, 0, 1]]


def test_compose_seeding_matches_between_product_and_oracle(T):
    """both Compose classes hand the same derived seeds to their randomizable children, and the children draw in the same order"""
    mk = lambda M: M.Compose([M.RandSpatialCropSamplesd(["image", "label"], (8, 8, 8), num_samples=3, random_size=False),
                              M.RandFlipd(["image", "label"], prob=0.5, spatial_axis=[0]), M.RandRotate90d(["image", "label"], prob=0.5, max_k=3),
                              M.RandShiftIntensityd(["image"], offsets=0.1, prob=0.5)]).set_random_state(seed=123)
    a, b = mk(T), mk(TO)
    for ta, tb in zip(a.transforms, b.transforms):
        assert ta.R.randint(1 << 30) == tb.R.randint(1 << 30)


# ------------------------------------------------------------------------------------------------------------------ GPU
DEV = "cuda:0"


def volume(seed, shape=(40, 36, 44), channels=1, nlab=4):
    rng = np.random.RandomState(seed)
    img = rng.rand(channels, *shape).astype(np.float32) - 0.3            # ~30 % of the voxels below the image threshold 0
    lab = np.zeros((1, *shape), np.float32)
    lab[0, 5:17, 8:20, 10:30] = rng.randint(1, nlab, (12, 12, 20))      # a labelled blob
    lab[0, 30:33, 2:5, 40:44] = 1                                         # and a sliver at the border (crop centres get clamped)
    return img, lab


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1, 7])
@pytest.mark.parametrize("roi", [(16, 16, 16), (17, 17, 17), (40, 36, 20)])
def test_segmentation_pipeline_matches_oracle(T, seed, roi):
    """RandCropByPosNegLabeld(num_samples=4) -> RandFlipd x3 -> RandRotate90d -> RandShiftIntensityd (seg:341-375), probabilities raised
    so that every branch fires: the fused GPU gather must equal the numpy restatement bit for bit under the same seed."""
    img, lab = volume(seed)
    rot = roi[0] == roi[1]

    def mk(M):
        ts = [M.RandCropByPosNegLabeld(keys=["image", "label"], label_key="label", spatial_size=roi, pos=1, neg=1, num_samples=4,
                                       image_key="image", image_threshold=0),
              M.RandFlipd(keys=["image", "label"], spatial_axis=[0], prob=0.5), M.RandFlipd(keys=["image", "label"], spatial_axis=[1], prob=0.5),
              M.RandFlipd(keys=["image", "label"], spatial_axis=[2], prob=0.5)]
        if rot:
            ts.append(M.RandRotate90d(keys=["image", "label"], prob=0.6, max_k=3))
        ts.append(M.RandShiftIntensityd(keys=["image"], offsets=0.10, prob=0.5))
        return M.Compose(ts).set_random_state(seed=100 + seed)
    want = mk(TO)({"image": img, "label": lab})
    pipe = mk(T)
    data = {"image": torch.from_numpy(img).to(DEV), "label": torch.from_numpy(lab).to(DEV)}
    for rep in range(2):                       # second call: cached index, fresh draws
        got = pipe(data)
        if rep:
            want = mk(TO)
            want = [want({"image": img, "label": lab}), want({"image": img, "label": lab})][1]
        assert len(got) == 4
        for g, w in zip(got, want):
            assert torch.equal(g["image"].cpu(), torch.from_numpy(np.ascontiguousarray(w["image"])))
            assert torch.equal(g["label"].cpu(), torch.from_numpy(np.ascontiguousarray(w["label"])))


@pytest.mark.gpu
def test_ranking_pipeline_and_brats_conversion_match_oracle(T):
    """ConvertToMultiChannelBasedOnBratsClassesd -> RandSpatialCropSamplesd(num_samples=2) -> flips -> rot90 -> shift (rank:365-369 with
    the Task01 label conversion of seg:65-93 in front): conversion fused into the gather."""
    img, lab = volume(3, channels=4)

    def mk(M):
        return M.Compose([M.ConvertToMultiChannelBasedOnBratsClassesd(keys=["label"]),
                          M.RandSpatialCropSamplesd(keys=["image", "label"], roi_size=(24, 24, 24), random_size=False, num_samples=2),
                          M.RandFlipd(keys=["image", "label"], spatial_axis=[0], prob=0.5), M.RandFlipd(keys=["image", "label"], spatial_axis=[2], prob=0.5),
                          M.RandRotate90d(keys=["image", "label"], prob=0.7, max_k=3),
                          M.RandShiftIntensityd(keys=["image"], offsets=0.10, prob=0.9)]).set_random_state(seed=5)
    want = mk(TO)({"image": img, "label": lab})
    got = mk(T)({"image": torch.from_numpy(img).to(DEV), "label": torch.from_numpy(lab).to(DEV)})
    assert len(got) == 2 and tuple(got[0]["label"].shape) == (4, 24, 24, 24)
    for g, w in zip(got, want):
        assert torch.equal(g["image"].cpu(), torch.from_numpy(np.ascontiguousarray(w["image"])))
        assert torch.equal(g["label"].cpu(), torch.from_numpy(np.ascontiguousarray(w["label"])))


@pytest.mark.gpu
def test_unfused_transforms_and_edge_cases(T):
    img, lab = volume(9)
    data = {"image": torch.from_numpy(img).to(DEV), "label": torch.from_numpy(lab).to(DEV)}
    # single transforms applied on their own (one gather each) equal numpy
    f = T.RandFlipd(keys=["image", "label"], spatial_axis=[1], prob=1.0).set_random_state(seed=1)
    out = f(data)
    assert torch.equal(out["image"].cpu(), torch.from_numpy(np.ascontiguousarray(np.flip(img, 2))))
    assert torch.equal(out["label"].cpu(), torch.from_numpy(np.ascontiguousarray(np.flip(lab, 2))))
    conv = T.ConvertToMultiChannelBasedOnBratsClassesd(keys=["label"])(data)["label"]
    want = TO.ConvertToMultiChannelBasedOnBratsClassesd(["label"])({"label": lab})["label"]
    assert torch.equal(conv.cpu(), torch.from_numpy(want))
    # no foreground at all: every crop comes from the background set; no sampling location: ValueError (MONAI's message)
    empty = {"image": data["image"], "label": torch.zeros_like(data["label"])}
    crop = T.RandCropByPosNegLabeld(keys=["image", "label"], label_key="label", spatial_size=(8, 8, 8), num_samples=3, image_key="image").set_random_state(seed=2)
    ref = TO.RandCropByPosNegLabeld(keys=["image", "label"], label_key="label", spatial_size=(8, 8, 8), num_samples=3, image_key="image").set_random_state(seed=2)
    got, want = crop(empty), ref({"image": img, "label": np.zeros_like(lab)})
    for g, w in zip(got, want):
        assert torch.equal(g["image"].cpu(), torch.from_numpy(np.ascontiguousarray(w["image"])))
    with pytest.raises(ValueError, match="No sampling location"):
        crop({"image": torch.full_like(data["image"], -1.0), "label": torch.zeros_like(data["label"])})
    with pytest.raises(ValueError, match="larger than the image"):
        T.RandCropByPosNegLabeld(keys=["image", "label"], label_key="label", spatial_size=(64, 8, 8), image_key="image")(data)
    with pytest.raises(TypeError):
        crop({"image": torch.from_numpy(img), "label": torch.from_numpy(lab)})          # CPU tensors: no CPU path
