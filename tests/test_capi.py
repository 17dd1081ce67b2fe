"""CPU-side checks of the product package: the C-ABI library loads and exports everything the header declares, the
module keeps the reference's constructor / state-dict / error contract, and the host-side index logic (window
enumeration, sharding, triplet ids, gradient reach) matches the oracle.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import unetr_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "unetr_b200.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 16
    lib = pkg._lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/unetr_b200.h but not exported"
    assert declared == set(pkg._lib.SIGNATURES), "ctypes signature table out of sync with the header"


def test_create_validates_config_without_gpu(pkg):
    lib = pkg._lib.load()
    bad = pkg._lib.UnetrConfig(1, 1, 14, 96, 96, 90, 16, 768, 3072, 12, 0, 1)
    assert not lib.b200_unetr_create(ctypes.byref(bad))
    assert "16" in pkg._lib.last_error()
    ok = pkg._lib.UnetrConfig(2, 1, 14, 96, 96, 96, 16, 768, 3072, 12, 0, 1)
    h = lib.b200_unetr_create(ctypes.byref(ok))
    assert h
    train, infer = lib.b200_unetr_workspace_bytes(h, 1), lib.b200_unetr_workspace_bytes(h, 0)
    assert train > infer > 0
    lib.b200_unetr_destroy(h)


def test_state_dict_contract_matches_reference_layout(pkg):
    kw = dict(in_channels=1, out_channels=14, img_size=(96, 96, 96), feature_size=16, hidden_size=768, mlp_dim=3072,
              num_heads=12, pos_embed="perceptron", norm_name="instance", res_block=True)
    torch.manual_seed(0)
    mine = pkg.UNETR(**kw)
    torch.manual_seed(0)
    ref = O.UNETR(**kw)
    a, b = mine.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys()) and len(a) == 165
    for k in a:
        assert a[k].shape == b[k].shape, k
        assert torch.equal(a[k], b[k]), f"default initialisation differs for {k}"
    mine.load_state_dict(ref.state_dict(), strict=True)
    assert sum(p.numel() for p in mine.parameters()) == 92_453_038
    assert len(mine._ordered_params()) == pkg._lib.PARAM_COUNT
    assert len({id(p) for p in mine._ordered_params()}) == 164


def test_conv_patch_embedding_keys(pkg):
    m = pkg.UNETR(4, 3, (32, 32, 32), 8, 64, 128, 4, "conv", "instance", res_block=True)
    r = O.UNETR(4, 3, (32, 32, 32), 8, 64, 128, 4, "conv", "instance", res_block=True)
    assert list(m.state_dict().keys()) == list(r.state_dict().keys())
    assert m.state_dict()["vit.patch_embedding.patch_embeddings.weight"].shape == (64, 4, 16, 16, 16)


def test_constructor_errors(pkg):
    kw = dict(in_channels=1, out_channels=2, img_size=(32,) * 3, feature_size=8, hidden_size=64, mlp_dim=128,
              num_heads=4, pos_embed="perceptron", norm_name="instance", res_block=True)
    with pytest.raises(AssertionError):
        pkg.UNETR(**{**kw, "dropout_rate": 1.5})
    with pytest.raises(AssertionError):
        pkg.UNETR(**{**kw, "num_heads": 5})
    with pytest.raises(KeyError):
        pkg.UNETR(**{**kw, "pos_embed": "sincos"})
    with pytest.raises(NotImplementedError):   # unsupported variants raise instead of falling back
        pkg.UNETR(**{**kw, "norm_name": "batch"})
    with pytest.raises(NotImplementedError):
        pkg.UNETR(**{**kw, "res_block": False})
    assert pkg.DiceCELoss(to_onehot_y=False, sigmoid=True).variant == "sigmoid"       # seg:480
    assert pkg.DiceCELoss(to_onehot_y=True, softmax=True).variant == "softmax"         # seg:404
    for bad in (dict(to_onehot_y=True, sigmoid=True), dict(softmax=True, sigmoid=True, to_onehot_y=True), dict(),
                dict(to_onehot_y=True, softmax=True, squared_pred=True)):
        with pytest.raises(NotImplementedError):
            pkg.DiceCELoss(**bad)
    with pytest.raises(NotImplementedError):
        pkg.DiceMetric(include_background=False)
    with pytest.raises(NotImplementedError):
        pkg.ConfusionMatrixMetric(metric_name="f1 score")
    with pytest.raises(RuntimeError):           # metrics have no CPU path either
        pkg.DiceMetric()(y_pred=[torch.zeros(2, 4, 4, 4)], y=[torch.zeros(2, 4, 4, 4)])


def test_no_cpu_fallback(pkg):
    m = pkg.UNETR(1, 2, (32,) * 3, 8, 64, 128, 4, "perceptron", "instance", res_block=True)
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        m(torch.zeros(1, 1, 32, 32, 32))
    with pytest.raises(RuntimeError):
        pkg.DiceCELoss(to_onehot_y=True, softmax=True)(torch.zeros(1, 2, 4, 4, 4), torch.zeros(1, 1, 4, 4, 4))
    with pytest.raises(RuntimeError):
        pkg.sliding_window_inference(torch.zeros(1, 1, 20, 20, 20), (16,) * 3, 4, lambda w: w)


@pytest.mark.parametrize("size,roi,overlap", [((512, 512, 256), (96,) * 3, 0.5), ((40, 33, 50), (16,) * 3, 0.25),
                                              ((96, 96, 96), (96,) * 3, 0.25), ((20, 16, 16), (16,) * 3, 0.8)])
def test_window_enumeration_matches_oracle(pkg, size, roi, overlap):
    per_axis, flat = pkg.window_starts(size, roi, overlap)
    assert flat == O.dense_window_starts(size, roi, O.scan_intervals(size, roi, overlap))
    assert len(flat) == len(per_axis[0]) * len(per_axis[1]) * len(per_axis[2])


def test_window_sharding_is_a_partition(pkg):
    for n, ws in [(500, 8), (18, 4), (5, 8), (196, 3)]:
        seen = []
        for r in range(ws):
            seen += list(pkg.shard_windows(n, r, ws))
        assert seen == list(range(n))


def test_triplet_ids_match_reference_enumeration(pkg):
    f = torch.zeros(2, 3, 8, 8, 8)
    np.random.seed(3)
    ref, sim, dis = pkg.extract_triplets_more_partitions(f, f, 3)
    assert len(ref) == len(sim) == len(dis) == 576
    assert list(zip(ref, sim, dis)) == O.triplet_ids()
    np.random.seed(3)
    assert ref.plan[3] == O.slice_indices(8)
    with pytest.raises(TypeError):
        pkg.BTLoss([f], [f], [f], None)


def test_gradient_reach_table(pkg):
    reach = pkg.UNETR._grad_reach
    seg = reach(True, True, True)
    assert all(seg)
    feat = reach(False, True, True)          # loss on enc4 only (rank:259-260)
    assert feat[0] and feat[3 + 9 * 11 + 10] and not feat[3 + 10 * 11] and not feat[135] and feat[145] and not feat[146]
    recon = reach(True, False, True)         # freeze_encoder=True (rank:261-262)
    assert not any(recon[:146]) and all(recon[146:])


@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("size,roi,overlap", [((40, 24, 24), (16, 16, 16), 0.5), ((512, 512, 256), (96, 96, 96), 0.5), ((33, 16, 20), (16, 16, 16), 0.25)])
def test_slab_plan_reproduces_the_sequential_sum(pkg, world, size, roi, overlap):
    """Host logic of the slab-owned sliding window (inferers.slab_plan), emulated with numpy along the sharded axis only: every padded
    row has exactly one owner, and adding each slab's pieces in plan order gives, bit for bit, the
    float32 sums of the sequential window loop (MONAI's order)."""
    import importlib
    import numpy as np
    inf = importlib.import_module("3dmedicalimagesegmentation_b200.inferers")
    per_axis, flat = inf.window_starts(size, roi, overlap)
    if len(flat) < world:
        pytest.skip("fewer windows than ranks")
    chunks, bounds, pieces = inf.slab_plan(flat, roi[0], size[0], world)
    assert bounds[0] == 0 and bounds[-1] == size[0] and all(a <= b for a, b in zip(bounds, bounds[1:]))
    assert sorted(w for c in chunks for w in c) == list(range(len(flat)))
    rng = np.random.default_rng(0)
    # one value per (window, row): the y/z extent does not matter for ownership, so the emulation keeps the first axis only,
    # but distinguishes windows that share an x-start by giving every (window, row) its own value and its own (y, z) column
    vals = rng.standard_normal((len(flat), roi[0])).astype(np.float32) * 100
    cols = {}
    seq = {}
    for w, (xs, ys, zs) in enumerate(flat):
        for i in range(roi[0]):
            key = (xs + i, ys, zs)          # a representative voxel of this window row
            seq[key] = np.float32(seq.get(key, np.float32(0)) + vals[w, i])
    # windows with different (ys, zs) do not meet at their representative voxel; windows with equal (ys, zs) and different xs do
    got = {}
    for d in range(world):
        last = -1
        for (w, src, lo, hi) in pieces[d]:
            assert w in chunks[src] and w > last - 1 and bounds[d] <= lo < hi <= bounds[d + 1]
            last = w
            xs, ys, zs = flat[w]
            for x in range(lo, hi):
                key = (x, ys, zs)
                got[key] = np.float32(got.get(key, np.float32(0)) + vals[w, x - xs])
    assert got.keys() == seq.keys()
    assert all(got[k] == seq[k] for k in seq)
