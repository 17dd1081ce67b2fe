"""CPU oracle for the UNETR hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.  The product package never does (it raises if its CUDA
library is missing).

What this restates
------------------
A plain-PyTorch fp32 restatement of the arithmetic the reference delegates to the
un-vendored dependency ``monai==0.6.0`` (pytorch_env.yml:94), reached from

  * unetr.py:69-208            UNETR wiring (`UNETR.__init__`, `proj_feat`, `forward`)
  * unetr_segmentation_3d.py:404,222   DiceCELoss(to_onehot_y=True, softmax=True)
  * unetr_segmentation_3d.py:109,143,694   sliding_window_inference
  * unetr_ranking_pretraining_3d.py:59-133,202-217   triplet extraction + Bradley-Terry loss

PARITY STATUS
-------------
* Ranking path (a14/a15): PINNED.  `oracle/make_golden.py` executes the reference's own
  `extract_triplets_more_partitions` and `BTLoss` function bodies (loaded from
  /root/reference by `ast`, unmodified) and the results are committed under
  `tests/golden/ranking_*.npz`; `tests/test_oracle.py` checks this restatement against them.
* Network / DiceCE / sliding window: PARITY UNPINNED.  MONAI 0.6.0 is not installable
  here and the reference ships no tests or golden vectors, so these follow the behavioural
  spec in SURVEY.md Appendix B and are pinned only by closed-form known-answer tests
  (T1, T4-T7 in SURVEY.md section 4).

State-dict keys and shapes equal the MONAI ones so a checkpoint written
by the reference loads here and in the product module alike.
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# conv decoder blocks  (MONAI dynunet_block / unetr_block semantics)
# --------------------------------------------------------------------------------------


def _same_padding(kernel: int, stride: int) -> int:
    return int((kernel - stride + 1) / 2)


class ConvOnly(nn.Sequential):
    """`get_conv_layer(..., conv_only=True)`: a Sequential whose single child is named `conv`."""

    def __init__(self, cin, cout, kernel, stride, bias=False, transposed=False):
        super().__init__()
        pad = _same_padding(kernel, stride)
        if transposed:
            out_pad = 2 * pad + stride - kernel
            conv = nn.ConvTranspose3d(cin, cout, kernel, stride, padding=pad, output_padding=out_pad, bias=bias)
        else:
            conv = nn.Conv3d(cin, cout, kernel, stride, padding=pad, bias=bias)
        self.add_module("conv", conv)


class ResBlock(nn.Module):
    """UnetResBlock with norm_name='instance' (Appendix B.2)."""

    def __init__(self, cin, cout, kernel=3, stride=1):
        super().__init__()
        self.conv1 = ConvOnly(cin, cout, kernel, stride)
        self.conv2 = ConvOnly(cout, cout, kernel, 1)
        self.conv3 = ConvOnly(cin, cout, 1, stride)
        self.lrelu = nn.LeakyReLU(negative_slope=0.01, inplace=True)
        self.norm1 = nn.InstanceNorm3d(cout)
        self.norm2 = nn.InstanceNorm3d(cout)
        self.norm3 = nn.InstanceNorm3d(cout)
        self.downsample = cin != cout or stride != 1

    def forward(self, inp):
        main = self.lrelu(self.norm1(self.conv1(inp)))
        main = self.norm2(self.conv2(main))
        skip = self.norm3(self.conv3(inp)) if self.downsample else inp
        return self.lrelu(main + skip)


class BasicBlock(nn.Module):  # UnetrBasicBlock(res_block=True), B.3
    def __init__(self, cin, cout):
        super().__init__()
        self.layer = ResBlock(cin, cout)

    def forward(self, x):
        return self.layer(x)


class PrUpBlock(nn.Module):  # UnetrPrUpBlock(conv_block=False), B.4
    def __init__(self, cin, cout, num_layer):
        super().__init__()
        self.transp_conv_init = ConvOnly(cin, cout, 2, 2, transposed=True)
        self.blocks = nn.ModuleList([ConvOnly(cout, cout, 2, 2, transposed=True) for _ in range(num_layer)])

    def forward(self, x):
        x = self.transp_conv_init(x)
        for blk in self.blocks:
            x = blk(x)
        return x


class UpBlock(nn.Module):  # UnetrUpBlock(res_block=True), B.5
    def __init__(self, cin, cout):
        super().__init__()
        self.transp_conv = ConvOnly(cin, cout, 2, 2, transposed=True)
        self.conv_block = ResBlock(2 * cout, cout)

    def forward(self, inp, skip):
        up = self.transp_conv(inp)
        return self.conv_block(torch.cat((up, skip), dim=1))


class OutBlock(nn.Module):  # UnetOutBlock, B.6
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = ConvOnly(cin, cout, 1, 1, bias=True)

    def forward(self, x):
        return self.conv(x)


# --------------------------------------------------------------------------------------
# ViT encoder  (Appendix B.7)
# --------------------------------------------------------------------------------------


class _ToPatchRows(nn.Module):
    """einops 'b c (h p1) (w p2) (d p3) -> b (h w d) (p1 p2 p3 c)' without einops."""

    def __init__(self, patch):
        super().__init__()
        self.patch = patch

    def forward(self, x):
        b, c, hh, ww, dd = x.shape
        p1, p2, p3 = self.patch
        h, w, d = hh // p1, ww // p2, dd // p3
        x = x.view(b, c, h, p1, w, p2, d, p3)
        x = x.permute(0, 2, 4, 6, 3, 5, 7, 1)  # b h w d p1 p2 p3 c
        return x.reshape(b, h * w * d, p1 * p2 * p3 * c)


class PatchEmbedding(nn.Module):
    def __init__(self, in_channels, img_size, patch_size, hidden, pos_embed):
        super().__init__()
        for m, p in zip(img_size, patch_size):
            if m < p:
                raise ValueError("patch_size should be smaller than img_size.")
            if pos_embed == "perceptron" and m % p != 0:
                raise ValueError("patch_size should be divisible by img_size for perceptron.")
        self.n_patches = int(np.prod([i // p for i, p in zip(img_size, patch_size)]))
        self.patch_dim = int(in_channels * np.prod(patch_size))
        self.pos_embed = pos_embed
        if pos_embed == "conv":
            self.patch_embeddings = nn.Conv3d(in_channels, hidden, kernel_size=patch_size, stride=patch_size)
        else:
            self.patch_embeddings = nn.Sequential(_ToPatchRows(patch_size), nn.Linear(self.patch_dim, hidden))
        self.position_embeddings = nn.Parameter(torch.zeros(1, self.n_patches, hidden))
        self.cls_token = nn.Parameter(torch.zeros(1, 1, hidden))
        nn.init.trunc_normal_(self.position_embeddings, mean=0.0, std=0.02, a=-2.0, b=2.0)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, mean=0.0, std=0.02, a=-2.0, b=2.0)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        x = self.patch_embeddings(x)
        if self.pos_embed == "conv":
            x = x.flatten(2).transpose(-1, -2)
        return x + self.position_embeddings


class MLP(nn.Module):
    def __init__(self, hidden, mlp_dim):
        super().__init__()
        self.linear1 = nn.Linear(hidden, mlp_dim)
        self.linear2 = nn.Linear(mlp_dim, hidden)
        self.fn = nn.GELU()

    def forward(self, x):
        return self.linear2(self.fn(self.linear1(x)))


class SelfAttention(nn.Module):
    def __init__(self, hidden, heads):
        super().__init__()
        self.out_proj = nn.Linear(hidden, hidden)
        self.qkv = nn.Linear(hidden, hidden * 3, bias=False)
        self.heads = heads
        self.scale = (hidden // heads) ** -0.5

    def forward(self, x):
        b, n, hid = x.shape
        d = hid // self.heads
        qkv = self.qkv(x).view(b, n, 3, self.heads, d).permute(2, 0, 3, 1, 4)  # qkv b l h d -> [3,b,heads,n,d]
        q, k, v = qkv[0], qkv[1], qkv[2]
        att = (torch.einsum("bhxd,bhyd->bhxy", q, k) * self.scale).softmax(dim=-1)
        y = torch.einsum("bhxy,bhyd->bhxd", att, v)
        y = y.permute(0, 2, 1, 3).reshape(b, n, hid)
        return self.out_proj(y)


class TransformerBlock(nn.Module):
    def __init__(self, hidden, mlp_dim, heads):
        super().__init__()
        self.mlp = MLP(hidden, mlp_dim)
        self.norm1 = nn.LayerNorm(hidden)
        self.attn = SelfAttention(hidden, heads)
        self.norm2 = nn.LayerNorm(hidden)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        return x


class ViT(nn.Module):
    def __init__(self, in_channels, img_size, patch_size, hidden, mlp_dim, num_layers, heads, pos_embed):
        super().__init__()
        self.patch_embedding = PatchEmbedding(in_channels, img_size, patch_size, hidden, pos_embed)
        self.blocks = nn.ModuleList([TransformerBlock(hidden, mlp_dim, heads) for _ in range(num_layers)])
        self.norm = nn.LayerNorm(hidden)

    def forward(self, x):
        x = self.patch_embedding(x)
        hidden_states = []
        for blk in self.blocks:
            x = blk(x)
            hidden_states.append(x)
        return self.norm(x), hidden_states


# --------------------------------------------------------------------------------------
# UNETR  (unetr.py:21-208)
# --------------------------------------------------------------------------------------


class UNETR(nn.Module):
    """Restatement of the reference network.  `tuple_output=True` is the local flavour
    (unetr.py:208 returns `(enc4, logits)`); False is `monai.networks.nets.UNETR` (logits only)."""

    def __init__(self, in_channels, out_channels, img_size, feature_size, hidden_size, mlp_dim, num_heads,
                 pos_embed, norm_name, conv_block=False, res_block=False, dropout_rate=0.0, tuple_output=True):
        super().__init__()
        if not (0 <= dropout_rate <= 1):  # unetr.py:60
            raise AssertionError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:  # unetr.py:63
            raise AssertionError("hidden size should be divisible by num_heads.")
        if pos_embed not in ["conv", "perceptron"]:  # unetr.py:66
            raise KeyError(f"Position embedding layer of type {pos_embed} is not supported.")
        if norm_name != "instance" or not res_block or conv_block or dropout_rate != 0.0:
            raise NotImplementedError("oracle covers the configuration both reference scripts use")
        self.tuple_output = tuple_output
        self.hidden_size = hidden_size
        self.patch_size = (16, 16, 16)  # unetr.py:70
        self.feat_size = tuple(s // p for s, p in zip(img_size, self.patch_size))
        fs = feature_size
        self.vit = ViT(in_channels, img_size, self.patch_size, hidden_size, mlp_dim, 12, num_heads, pos_embed)
        self.encoder1 = BasicBlock(in_channels, fs)
        self.encoder2 = PrUpBlock(hidden_size, fs * 2, num_layer=2)
        self.encoder3 = PrUpBlock(hidden_size, fs * 4, num_layer=1)
        self.encoder4 = PrUpBlock(hidden_size, fs * 8, num_layer=0)
        self.decoder5 = UpBlock(hidden_size, fs * 8)
        self.decoder4 = UpBlock(fs * 8, fs * 4)
        self.decoder3 = UpBlock(fs * 4, fs * 2)
        self.decoder2 = UpBlock(fs * 2, fs)
        self.out = OutBlock(fs, out_channels)

    def proj_feat(self, tokens):  # unetr.py:177-180
        b = tokens.size(0)
        vol = tokens.view(b, *self.feat_size, self.hidden_size)
        return vol.permute(0, 4, 1, 2, 3).contiguous()

    def _encode(self, x_in):
        x, hs = self.vit(x_in)
        enc1 = self.encoder1(x_in)
        enc2 = self.encoder2(self.proj_feat(hs[3]))
        enc3 = self.encoder3(self.proj_feat(hs[6]))
        enc4 = self.encoder4(self.proj_feat(hs[9]))
        return x, enc1, enc2, enc3, enc4

    def forward(self, x_in, freeze_encoder=False, return_intermediates=False):
        if freeze_encoder:  # unetr.py:183-192
            with torch.no_grad():
                x, enc1, enc2, enc3, enc4 = self._encode(x_in)
        else:
            x, enc1, enc2, enc3, enc4 = self._encode(x_in)
        dec4 = self.proj_feat(x)
        dec3 = self.decoder5(dec4, enc4)
        dec2 = self.decoder4(dec3, enc3)
        dec1 = self.decoder3(dec2, enc2)
        out = self.decoder2(dec1, enc1)
        logits = self.out(out)
        if return_intermediates:
            return dict(vit=x, enc1=enc1, enc2=enc2, enc3=enc3, enc4=enc4, dec3=dec3, dec2=dec2, dec1=dec1,
                        out=out, logits=logits)
        return (enc4, logits) if self.tuple_output else logits


# --------------------------------------------------------------------------------------
# DiceCELoss(to_onehot_y=True, softmax=True)   (Appendix B.8; call sites seg:404,222)
# --------------------------------------------------------------------------------------


def dice_ce_loss(logits: torch.Tensor, target: torch.Tensor, smooth_nr=1e-5, smooth_dr=1e-5,
                 return_terms=False):
    n_cls = logits.shape[1]
    if target.shape[1] != 1 or target.shape[0] != logits.shape[0] or target.shape[2:] != logits.shape[2:]:
        raise AssertionError(f"ground truth has differing shape ({target.shape}) from input ({logits.shape})")
    prob = torch.softmax(logits, dim=1)
    onehot = torch.zeros_like(prob).scatter_(1, target.long(), 1.0)
    axes = tuple(range(2, logits.dim()))
    inter = (onehot * prob).sum(axes)
    denom = onehot.sum(axes) + prob.sum(axes)
    dice = (1.0 - (2.0 * inter + smooth_nr) / (denom + smooth_dr)).mean()
    ce = F.cross_entropy(logits, target.squeeze(1).long(), reduction="mean")
    if return_terms:
        return dice + ce, dice, ce
    return dice + ce


def dice_metric(pred_onehot: torch.Tensor, y_onehot: torch.Tensor) -> torch.Tensor:
    """DiceMetric(include_background=True) per (b,c): 2|y&p|/(|y|+|p|), NaN when |y|==0 (Appendix B.10)."""
    axes = tuple(range(2, pred_onehot.dim()))
    inter = (pred_onehot * y_onehot).sum(axes)
    y_o = y_onehot.sum(axes)
    den = y_o + pred_onehot.sum(axes)
    return torch.where(y_o > 0, 2.0 * inter / den, torch.full_like(inter, float("nan")))


def dice_ce_loss_sigmoid(logits: torch.Tensor, target: torch.Tensor, smooth_nr=1e-5, smooth_dr=1e-5, return_terms=False):
    """DiceCELoss(to_onehot_y=False, sigmoid=True) as configured at unetr_segmentation_3d.py:480:
    Dice on sigmoid(logits) against the float multi-hot target [B,C,...]; CE against argmax(target, 1) because the target has
    as many channels as the prediction (MONAI 0.6.0 DiceCELoss.ce)."""
    if target.shape != logits.shape:
        raise AssertionError(f"ground truth has differing shape ({target.shape}) from input ({logits.shape})")
    prob = torch.sigmoid(logits)
    axes = tuple(range(2, logits.dim()))
    inter = (target * prob).sum(axes)
    denom = target.sum(axes) + prob.sum(axes)
    dice = (1.0 - (2.0 * inter + smooth_nr) / (denom + smooth_dr)).mean()
    ce = F.cross_entropy(logits, torch.argmax(target, dim=1).long(), reduction="mean")
    if return_terms:
        return dice + ce, dice, ce
    return dice + ce


def brats_multichannel(label: torch.Tensor) -> torch.Tensor:
    """ConvertToMultiChannelBasedOnBratsClassesd (unetr_segmentation_3d.py:65-93) on a [B,1,...] or [B,...] label map:
    channels (background, TC = 2|3, WT = 1|2|3, ET = 3), float32."""
    if label.dim() == 5:
        label = label[:, 0]
    return torch.stack([label == 0, (label == 2) | (label == 3), (label == 1) | (label == 2) | (label == 3), label == 3],
                       dim=1).float()


def metric_reduce(f: torch.Tensor, reduction: str):
    """monai.metrics.utils.do_metric_reduction (0.6.0, recalled) for "mean" and "mean_batch" on f[N,C,...]: NaN-aware."""
    f = f.clone()
    nans = torch.isnan(f)
    not_nans = (~nans).float()
    f[nans] = 0
    zero = torch.zeros(1, dtype=f.dtype)
    if reduction == "mean":
        not_nans = not_nans.sum(dim=1)
        f = torch.where(not_nans > 0, f.sum(dim=1) / not_nans, zero)
        not_nans = (not_nans > 0).float().sum(dim=0)
        f = torch.where(not_nans > 0, f.sum(dim=0) / not_nans, zero)
    elif reduction == "mean_batch":
        not_nans = not_nans.sum(dim=0)
        f = torch.where(not_nans > 0, f.sum(dim=0) / not_nans, zero)
    else:
        raise ValueError(reduction)
    return f, not_nans


def confusion_matrix(pred_onehot: torch.Tensor, y_onehot: torch.Tensor) -> torch.Tensor:
    """monai.metrics.get_confusion_matrix (include_background=True): [B,C,4] = (tp, fp, tn, fn)."""
    b, c = pred_onehot.shape[:2]
    p = pred_onehot.reshape(b, c, -1).float()
    y = y_onehot.reshape(b, c, -1).float()
    tp = ((p + y) == 2).float().sum(2)
    tn = ((p + y) == 0).float().sum(2)
    pos = y.sum(2)
    neg = y.shape[-1] - pos
    return torch.stack([tp, neg - tn, tn, pos - tp], dim=-1)


def confusion_metric(cm: torch.Tensor, metric_name: str) -> torch.Tensor:
    """compute_confusion_matrix_metric for "precision" (tp/(tp+fp)) and "sensitivity" (tp/(tp+fn)); NaN on a zero denominator."""
    tp, fp, fn = cm[..., 0], cm[..., 1], cm[..., 3]
    den = tp + fp if metric_name == "precision" else tp + fn
    return torch.where(den != 0, tp / den, torch.full_like(den, float("nan")))


def confusion_aggregate(cm: torch.Tensor, metric_name: str, reduction: str, compute_sample: bool = False) -> torch.Tensor:
    """ConfusionMatrixMetric.aggregate (0.6.0, recalled): compute_sample=False (the default the reference uses, seg:487-494)
    reduces the confusion matrix first and evaluates the metric on the reduced counts."""
    if compute_sample:
        return metric_reduce(confusion_metric(cm, metric_name), reduction)[0]
    return confusion_metric(metric_reduce(cm, reduction)[0], metric_name)


# --------------------------------------------------------------------------------------
# sliding_window_inference   (Appendix B.9; call sites seg:109,143,694)
# --------------------------------------------------------------------------------------


def scan_intervals(image_size, roi, overlap):
    return tuple(int(r) if r == s else max(int(r * (1 - overlap)), 1) for r, s in zip(roi, image_size))


def dense_window_starts(image_size, roi, interval) -> List[Tuple[int, ...]]:
    """Window start coordinates in the reference's order (first spatial dim slowest)."""
    per_dim = []
    for size, r, step in zip(image_size, roi, interval):
        num = int(math.ceil(float(size) / step))
        scan = -1
        for d in range(num):
            if d * step + r >= size:
                scan = d
                break
        n = scan + 1 if scan != -1 else 1
        per_dim.append([i * step - max(i * step + r - size, 0) for i in range(n)])
    return list(itertools.product(*per_dim))


def sliding_window_inference(inputs: torch.Tensor, roi_size, sw_batch_size: int, predictor: Callable,
                             overlap: float = 0.25, mode: str = "constant", cval: float = 0.0):
    if mode != "constant":
        raise NotImplementedError("the reference scripts only use constant blending")
    nd = inputs.dim() - 2
    roi = tuple(roi_size) if isinstance(roi_size, (tuple, list)) else (roi_size,) * nd
    batch = inputs.shape[0]
    orig = tuple(inputs.shape[2:])
    size = tuple(max(o, r) for o, r in zip(orig, roi))
    pads = []
    for k in range(nd - 1, -1, -1):
        diff = max(roi[k] - orig[k], 0)
        half = diff // 2
        pads.extend([half, diff - half])
    x = F.pad(inputs, pads, mode="constant", value=cval)
    starts = dense_window_starts(size, roi, scan_intervals(size, roi, overlap))
    num_win = len(starts)
    total = num_win * batch
    out = cnt = None
    for g in range(0, total, sw_batch_size):
        ids = range(g, min(g + sw_batch_size, total))
        where = [(i // num_win, starts[i % num_win]) for i in ids]
        data = torch.cat([x[b:b + 1, :, s[0]:s[0] + roi[0], s[1]:s[1] + roi[1], s[2]:s[2] + roi[2]] for b, s in where])
        pred = predictor(data)
        if out is None:
            out = torch.zeros((batch, pred.shape[1]) + size, dtype=torch.float32)
            cnt = torch.zeros_like(out)
        for k, (b, s) in enumerate(where):
            sl = (b, slice(None), slice(s[0], s[0] + roi[0]), slice(s[1], s[1] + roi[1]), slice(s[2], s[2] + roi[2]))
            out[sl] += pred[k]
            cnt[sl] += 1.0
    out = out / cnt
    crop = [slice(None), slice(None)]
    for k in range(nd):
        half = max(roi[k] - orig[k], 0) // 2
        crop.append(slice(half, half + orig[k]))
    return out[tuple(crop)]


# --------------------------------------------------------------------------------------
# Ranking pre-training loss   (rank:59-133 triplets, rank:202-217 BTLoss)
# --------------------------------------------------------------------------------------

NUM_PARTITIONS = 4  # rank:330


def slice_indices(dim_size: int, rng=np.random, num_partitions: int = NUM_PARTITIONS) -> List[int]:
    """rank:73-76 -- one draw from the global numpy RNG, then one index per partition."""
    part = int(dim_size / num_partitions)
    first = rng.choice(np.arange(0, part))
    return [int(first + p * part) for p in range(num_partitions)]


def gather_slices(batch1: torch.Tensor, batch2: torch.Tensor, slice_dimension: int, idx: Sequence[int]):
    """rank:77-118 -- 16 `[C, F]` slices ordered (partition, [b1[0], b1[1], b2[0], b2[1]])."""
    c = batch1.shape[1]
    rows = []
    for i in idx:
        for vol in (batch1[0], batch1[1], batch2[0], batch2[1]):
            rows.append(vol.select(slice_dimension - 1, i).reshape(c, -1))
    return rows


def triplet_ids(num_partitions: int = NUM_PARTITIONS, per_partition: int = 4):
    """rank:120-132 -- (ref, sim, dissim) slice ids in the reference's enumeration order."""
    out = []
    for p in range(num_partitions):
        mine = [p * per_partition + j for j in range(per_partition)]
        others = [q * per_partition + j for q in range(num_partitions) if q != p for j in range(per_partition)]
        for (r, s), d in itertools.product(itertools.permutations(mine, 2), others):
            out.append((r, s, d))
    return out


def bt_ranking_loss(batch1, batch2, slice_dimension, idx, temperature, eps=1e-6):
    """rank:202-212 -- sum over triplets of mean_c log(1+exp(-(cos(r,s)-cos(r,d))/T)).
    cos = torch.nn.CosineSimilarity(dim=-1, eps=1e-6) (rank:467)."""
    rows = gather_slices(batch1, batch2, slice_dimension, idx)
    loss = 0
    for r, s, d in triplet_ids():
        sim = F.cosine_similarity(rows[r], rows[s], dim=-1, eps=eps) / temperature
        dis = F.cosine_similarity(rows[r], rows[d], dim=-1, eps=eps) / temperature
        loss = loss + torch.mean(torch.log(1 + torch.exp(-(sim - dis))))
    return loss


def bt_ranking_loss_gram(batch1, batch2, slice_dimension, idx, temperature, eps=1e-6):
    """Same value through the 16x16 per-channel cosine Gram (the form the CUDA kernel uses)."""
    rows = torch.stack(gather_slices(batch1, batch2, slice_dimension, idx))  # [16,C,F]
    gram = torch.einsum("icf,jcf->cij", rows, rows)
    nrm = rows.norm(dim=-1).clamp_min(eps).transpose(0, 1)  # [C,16]
    cosm = gram / (nrm[:, :, None] * nrm[:, None, :])
    t = torch.tensor(triplet_ids())
    z = (cosm[:, t[:, 0], t[:, 1]] - cosm[:, t[:, 0], t[:, 2]]) / temperature
    return F.softplus(-z).mean(0).sum()


# --------------------------------------------------------------------------------------
# synthetic inputs / weights shared by tests and benchmarks
# --------------------------------------------------------------------------------------


def make_model(img=96, in_channels=1, out_channels=14, feature_size=16, hidden=768, mlp=3072, heads=12,
               pos_embed="perceptron", seed=0, tuple_output=True):
    torch.manual_seed(seed)
    m = UNETR(in_channels, out_channels, (img,) * 3, feature_size, hidden, mlp, heads, pos_embed, "instance",
              res_block=True, tuple_output=tuple_output)
    with torch.no_grad():  # widen the top-2 logit margin so argmax parity is meaningful
        m.out.conv.conv.weight.mul_(4.0)
        m.out.conv.conv.bias.copy_(torch.linspace(-1, 1, out_channels))
    return m


def make_inputs(batch=1, img=96, in_channels=1, n_classes=14, seed=1):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, in_channels, img, img, img, generator=g)
    g2 = torch.Generator().manual_seed(seed + 1)
    y = torch.randint(0, n_classes, (batch, 1, img, img, img), generator=g2).float()
    return x, y
