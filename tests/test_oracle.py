"""Pins the CPU oracle: golden vectors produced by the reference's own ranking functions, and the
closed-form known-answer tests T1-T7 of SURVEY.md section 4 for the MONAI-side restatement."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import unetr_oracle as O


@pytest.mark.parametrize("name", ["feat", "recon", "feat_T05"])
@pytest.mark.parametrize("sd", [2, 3, 4])
def test_ranking_matches_reference_golden(golden_dir, name, sd):
    g = np.load(os.path.join(golden_dir, f"ranking_{name}.npz"))
    feat = torch.from_numpy(g["feat"]).requires_grad_(True)
    temp = float(g["temperature"])
    # the index draw consumes the numpy RNG exactly like rank:75
    np.random.seed(int(g[f"npseed_sd{sd}"]))
    idx = O.slice_indices(feat.shape[sd])
    assert idx == g[f"idx_sd{sd}"].tolist()
    f1, f2 = torch.split(feat, [2, 2], dim=0)
    loss = O.bt_ranking_loss(f1, f2, sd, idx, temp)
    loss.backward()
    assert abs(loss.item() - float(g[f"loss_sd{sd}"])) <= 2e-4 * abs(float(g[f"loss_sd{sd}"]))
    ref_grad = torch.from_numpy(g[f"grad_sd{sd}"])
    assert torch.allclose(feat.grad, ref_grad, rtol=1e-4, atol=1e-6)
    # Gram formulation (what the CUDA kernel computes) agrees too
    loss2 = O.bt_ranking_loss_gram(f1.detach(), f2.detach(), sd, idx, temp)
    assert abs(loss2.item() - float(g[f"loss_sd{sd}"])) <= 2e-4 * abs(float(g[f"loss_sd{sd}"]))


def test_T2_identical_slices_is_576_ln2(golden_dir):
    g = np.load(os.path.join(golden_dir, "ranking_const.npz"))
    assert len(O.triplet_ids()) == 576
    feat = torch.ones(4, 4, 8, 8, 8)
    loss = O.bt_ranking_loss(feat[:2], feat[2:], 2, [0, 2, 4, 6], 0.1)
    assert abs(loss.item() - 576 * math.log(2)) < 1e-2
    assert abs(float(g["loss"]) - loss.item()) < 1e-3


def test_T3_gradient_sparsity():
    g = torch.Generator().manual_seed(0)
    feat = torch.randn(4, 6, 12, 12, 12, generator=g, requires_grad=True)
    O.bt_ranking_loss(feat[:2], feat[2:], 3, [1, 4, 7, 10], 0.1).backward()
    nz = (feat.grad != 0).float().mean().item()
    assert abs(nz - 4 / 12) < 1e-3


def test_T1_dicece_zero_logits_closed_form():
    b, c, s = 2, 14, 16
    g = torch.Generator().manual_seed(2)
    y = torch.randint(0, c, (b, 1, s, s, s), generator=g).float()
    loss = O.dice_ce_loss(torch.zeros(b, c, s, s, s), y).item()
    n = s ** 3
    counts = torch.stack([(y[i] == k).sum() for i in range(b) for k in range(c)]).double()
    dice = (1 - (2 * counts / c + 1e-5) / (counts + n / c + 1e-5)).mean().item()
    assert abs(loss - (math.log(c) + dice)) < 1e-5


def test_dicece_shape_mismatch_raises():
    with pytest.raises(AssertionError):
        O.dice_ce_loss(torch.zeros(1, 3, 4, 4, 4), torch.zeros(1, 1, 4, 4, 5))


def test_T4_convtranspose_is_gemm_plus_pixel_shuffle():
    torch.manual_seed(0)
    x = torch.randn(2, 6, 3, 3, 3)
    w = torch.randn(6, 4, 2, 2, 2)
    ref = F.conv_transpose3d(x, w, stride=2)
    rows = x.permute(0, 2, 3, 4, 1).reshape(-1, 6) @ w.reshape(6, 32)  # [vox, co*8]
    got = rows.view(2, 3, 3, 3, 4, 2, 2, 2).permute(0, 4, 1, 5, 2, 6, 3, 7).reshape(2, 4, 6, 6, 6)
    assert torch.allclose(ref, got, atol=1e-5)


def test_T5_conv_and_perceptron_patch_embed_agree():
    torch.manual_seed(0)
    a = O.PatchEmbedding(2, (32, 32, 32), (16, 16, 16), 24, "perceptron")
    b = O.PatchEmbedding(2, (32, 32, 32), (16, 16, 16), 24, "conv")
    with torch.no_grad():
        w = a.patch_embeddings[1].weight.view(24, 16, 16, 16, 2).permute(0, 4, 1, 2, 3)
        b.patch_embeddings.weight.copy_(w)
        b.patch_embeddings.bias.copy_(a.patch_embeddings[1].bias)
        b.position_embeddings.copy_(a.position_embeddings)
    x = torch.randn(1, 2, 32, 32, 32)
    assert torch.allclose(a(x), b(x), atol=1e-4)


@pytest.mark.parametrize("overlap,shape", [(0.25, (40, 33, 50)), (0.5, (48, 48, 32)), (0.8, (20, 16, 16))])
def test_T6_sliding_window_identity(overlap, shape):
    x = torch.rand(1, 1, *shape)
    y = O.sliding_window_inference(x, (16, 16, 16), 4, lambda w: w + 1, overlap=overlap)
    assert torch.equal(y, x + 1) or torch.allclose(y, x + 1, atol=1e-6)


def test_sliding_window_window_count_config5():
    starts = O.dense_window_starts((512, 512, 256), (96,) * 3, O.scan_intervals((512, 512, 256), (96,) * 3, 0.5))
    assert len(starts) == 500
    assert starts[0] == (0, 0, 0) and starts[-1] == (416, 416, 160)
    assert starts[1] == (0, 0, 48)  # last spatial dim fastest


def test_T7_param_count_and_state_dict_keys():
    m = O.UNETR(1, 14, (96, 96, 96), 16, 768, 3072, 12, "perceptron", "instance", res_block=True)
    assert sum(p.numel() for p in m.parameters()) == 92_453_038
    sd = m.state_dict()
    assert len(sd) == 165
    for k, shp in {
        "vit.patch_embedding.position_embeddings": (1, 216, 768),
        "vit.patch_embedding.cls_token": (1, 1, 768),
        "vit.patch_embedding.patch_embeddings.1.weight": (768, 4096),
        "vit.blocks.11.attn.qkv.weight": (2304, 768),
        "vit.blocks.0.mlp.linear1.weight": (3072, 768),
        "vit.norm.weight": (768,),
        "encoder1.layer.conv3.conv.weight": (16, 1, 1, 1, 1),
        "encoder2.transp_conv_init.conv.weight": (768, 32, 2, 2, 2),
        "encoder2.blocks.1.conv.weight": (32, 32, 2, 2, 2),
        "decoder5.transp_conv.conv.weight": (768, 128, 2, 2, 2),
        "decoder5.conv_block.conv1.conv.weight": (128, 256, 3, 3, 3),
        "decoder2.conv_block.conv3.conv.weight": (16, 32, 1, 1, 1),
        "out.conv.conv.weight": (14, 16, 1, 1, 1),
        "out.conv.conv.bias": (14,),
    }.items():
        assert tuple(sd[k].shape) == shp, k
    m128 = O.UNETR(1, 14, (128,) * 3, 16, 768, 3072, 12, "perceptron", "instance", res_block=True)
    assert sum(p.numel() for p in m128.parameters()) == 92_680_366


def test_ctor_errors_follow_reference():
    kw = dict(in_channels=1, out_channels=2, img_size=(32,) * 3, feature_size=8, hidden_size=64, mlp_dim=128,
              num_heads=4, pos_embed="perceptron", norm_name="instance", res_block=True)
    with pytest.raises(AssertionError):
        O.UNETR(**{**kw, "dropout_rate": 1.5})
    with pytest.raises(AssertionError):
        O.UNETR(**{**kw, "num_heads": 5})
    with pytest.raises(KeyError):
        O.UNETR(**{**kw, "pos_embed": "sincos"})


def test_tiny_unetr_regression_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "unetr_tiny.npz"))
    torch.manual_seed(0)
    m = O.UNETR(1, 5, (32, 32, 32), 8, 64, 128, 4, "perceptron", "instance", res_block=True)
    with torch.no_grad():
        m.out.conv.conv.weight.mul_(4.0)
        m.out.conv.conv.bias.copy_(torch.linspace(-1, 1, 5))
    x, y = O.make_inputs(batch=2, img=32, n_classes=5, seed=3)
    enc4, logits = m(x)
    assert enc4.shape == (2, 64, 4, 4, 4) and logits.shape == (2, 5, 32, 32, 32)
    assert abs(O.dice_ce_loss(logits, y).item() - float(g["loss"])) < 1e-4
    assert np.allclose(enc4.detach().numpy(), g["enc4"], atol=1e-4)


def test_freeze_encoder_grad_reach():
    torch.manual_seed(0)
    m = O.UNETR(1, 3, (32, 32, 32), 8, 64, 128, 4, "perceptron", "instance", res_block=True)
    x = torch.rand(1, 1, 32, 32, 32)
    _, logits = m(x, freeze_encoder=True)
    logits.sum().backward()
    assert m.vit.blocks[0].attn.qkv.weight.grad is None and m.encoder4.transp_conv_init.conv.weight.grad is None
    assert m.decoder5.transp_conv.conv.weight.grad is not None and m.out.conv.conv.bias.grad is not None
    m.zero_grad(set_to_none=True)
    enc4, _ = m(x)
    enc4.sum().backward()
    assert m.vit.blocks[9].mlp.linear1.weight.grad is not None
    assert m.vit.blocks[10].mlp.linear1.weight.grad is None and m.decoder5.transp_conv.conv.weight.grad is None


# ------------------------------------------------------------------------------- SURVEY 8f N2 / N3 restatements
def test_T8_sigmoid_dicece_zero_logits_closed_form():
    """DiceCELoss(to_onehot_y=False, sigmoid=True) (seg:480): zero logits -> p = 1/2 everywhere, CE = ln C."""
    b, c, s = 2, 4, 8
    g = torch.Generator().manual_seed(4)
    lab = torch.randint(0, 4, (b, 1, s, s, s), generator=g)
    t = O.brats_multichannel(lab)
    assert t.shape == (b, c, s, s, s)
    # BraTS channel rule (seg:79-91): label 2 -> TC and WT; label 3 -> TC, WT and ET; label 1 -> WT only
    assert t[:, 1].sum() == ((lab == 2) | (lab == 3)).sum() and t[:, 2].sum() == (lab > 0).sum() and t[:, 3].sum() == (lab == 3).sum()
    loss, dice, ce = O.dice_ce_loss_sigmoid(torch.zeros(b, c, s, s, s), t, return_terms=True)
    n = s ** 3
    gsum = t.sum((2, 3, 4)).double()
    closed = (1 - (gsum + 1e-5) / (gsum + n / 2 + 1e-5)).mean().item()
    assert abs(dice.item() - closed) < 1e-6 and abs(ce.item() - math.log(c)) < 1e-6
    with pytest.raises(AssertionError):
        O.dice_ce_loss_sigmoid(torch.zeros(1, 4, 4, 4, 4), torch.zeros(1, 1, 4, 4, 4))


def test_T9_dice_metric_nan_rules_and_reductions():
    """DiceMetric / do_metric_reduction (Appendix B.10): NaN when a class is absent from the label; "mean" averages classes
    first (ignoring NaN), then samples; "mean_batch" keeps the class axis."""
    y = torch.zeros(2, 3, 4, 4, 4)
    p = torch.zeros(2, 3, 4, 4, 4)
    y[0, 0] = 1; p[0, 0] = 1                                  # sample 0: class 0 perfect, classes 1,2 absent -> NaN
    y[1, 0, :2] = 1; y[1, 1, 2:] = 1                          # sample 1: class 0 = half the volume, class 1 = other half
    p[1, 0, :1] = 1; p[1, 1, 1:] = 1                          #           prediction shifts the boundary by one plane
    d = O.dice_metric(p, y)
    assert d[0, 0] == 1 and torch.isnan(d[0, 1:]).all() and torch.isnan(d[1, 2])
    assert abs(d[1, 0].item() - 2 * 16 / (32 + 16)) < 1e-6 and abs(d[1, 1].item() - 2 * 32 / (32 + 48)) < 1e-6
    mean, nn_ = O.metric_reduce(d, "mean")
    assert abs(mean.item() - (1.0 + (2 / 3 + 0.8) / 2) / 2) < 1e-6 and nn_.item() == 2
    mb, nb = O.metric_reduce(d, "mean_batch")
    assert torch.allclose(mb, torch.tensor([(1 + 2 / 3) / 2, 0.8, 0.0])) and nb.tolist() == [2, 1, 0]
    cm = O.confusion_matrix(p, y)
    assert cm[1, 0].tolist() == [16, 0, 32, 16] and cm[1, 1].tolist() == [32, 16, 16, 0]
    assert cm.sum(-1).eq(64).all()
    prec = O.confusion_metric(cm, "precision")
    assert prec[1, 0] == 1 and abs(prec[1, 1].item() - 32 / 48) < 1e-6 and torch.isnan(prec[0, 1])
    # compute_sample=False: counts are averaged first (micro average), then the metric is formed
    agg = O.confusion_aggregate(cm, "sensitivity", "mean")
    red = O.metric_reduce(cm, "mean")[0]
    assert abs(agg.item() - (red[0] / (red[0] + red[3])).item()) < 1e-7
