"""GPU parity tests (B200): the CUDA path, called through the C ABI, against the CPU oracle on identical seeded
inputs, the reference-generated golden vectors, and size-independent properties.

Tolerances (BASELINE.json north_star): logits rel-err <= 1e-4 in fp32 mode, <= 1e-2 in bf16 mode
(rel-err = max|a-b| / max|b|), Dice/loss abs diff <= 1e-3, argmax masks identical in fp32 mode."""
import copy
import math
import os

import numpy as np
import pytest
import torch

from oracle import unetr_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def cosine(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()


TINY = dict(in_channels=1, out_channels=5, img_size=(32, 32, 32), feature_size=8, hidden_size=64, mlp_dim=128,
            num_heads=4, pos_embed="perceptron", norm_name="instance", res_block=True)


def l2err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def grad_check(mine, ref, ref64, floor, cos_min=None):
    """Gradients against the fp64 oracle in relative L2 norm.  (Max-norm is not meaningful here: a voxel whose
    pre-activation sits within fp32 rounding of the LeakyReLU kink flips its slope 1 <-> 0.01 between any two fp32
    implementations -- observed: 1 element of 65 536 -- and that single voxel dominates a max-norm.)  The fp32 CPU
    oracle's own distance to fp64 sets the scale."""
    worst = ("", 0.0, 0.0)
    for (k, p), (_, q), (_, q64) in zip(mine.named_parameters(), ref.named_parameters(), ref64.named_parameters()):
        assert (p.grad is None) == (q.grad is None), k
        if q.grad is None:
            continue
        e_mine, e_ref = l2err(p.grad, q64.grad), l2err(q.grad, q64.grad)
        if cos_min is not None:
            assert cosine(p.grad, q64.grad) >= cos_min, (k, cosine(p.grad, q64.grad))
        if e_mine > worst[1]:
            worst = (k, e_mine, e_ref)
        assert e_mine <= max(floor, 8 * e_ref), (k, e_mine, e_ref)
    return worst


def to64(ref):
    import copy
    return copy.deepcopy(ref).double()


def build_pair(pkg, mode, **kw):
    cfg = {**TINY, **kw}
    torch.manual_seed(0)
    ref = O.UNETR(**cfg)
    with torch.no_grad():
        ref.out.conv.conv.weight.mul_(4.0)
        ref.out.conv.conv.bias.copy_(torch.linspace(-1, 1, cfg["out_channels"]))
    mine = pkg.UNETR(**cfg)
    mine.load_state_dict(ref.state_dict())
    return ref, mine.to(DEV).set_mode(mode)


# ------------------------------------------------------------------------------------------- full network
@pytest.mark.parametrize("mode,tol_logits,tol_grad,cos_min", [("fp32", 1e-4, 1e-2, 0.999), ("bf16", 2e-2, 0.5, 0.9)])
def test_tiny_unetr_forward_backward_matches_oracle(pkg, mode, tol_logits, tol_grad, cos_min):
    ref, mine = build_pair(pkg, mode)
    ref64 = to64(ref)
    x, y = O.make_inputs(batch=2, img=32, n_classes=5, seed=3)
    enc4_r, logits_r = ref(x)
    loss_r, dice_r, _ = O.dice_ce_loss(logits_r, y, return_terms=True)
    loss_r.backward()
    O.dice_ce_loss(ref64(x.double())[1], y.double()).backward()
    enc4, logits = mine(x.to(DEV))
    assert relerr(enc4, enc4_r) <= tol_logits * 2, ("enc4", relerr(enc4, enc4_r))
    assert relerr(logits, logits_r) <= tol_logits, ("logits", relerr(logits, logits_r))
    loss = pkg.DiceCELoss(to_onehot_y=True, softmax=True)(logits, y.to(DEV))
    assert abs(loss.item() - loss_r.item()) <= 1e-3
    loss.backward()
    if mode == "fp32":
        assert_argmax_parity(logits, logits_r)
    print(f"[tiny {mode}] worst grad (name, err vs fp64, oracle32 err vs fp64):", grad_check(mine, ref, ref64, tol_grad, cos_min))


def assert_argmax_parity(logits, logits_r, tie=1e-5):
    """Masks must agree wherever the reference's own top-2 margin is above fp32 rounding of the logits."""
    a, b = logits.argmax(1).cpu(), logits_r.argmax(1)
    top2 = logits_r.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    decided = margin > tie * logits_r.abs().max()
    bad = ((a != b) & decided).sum().item()
    assert bad == 0, f"{bad} argmax mismatches at voxels with a decided margin"
    return ((a != b) & ~decided).sum().item()


def test_golden_fixture_of_tiny_network(pkg, golden_dir):
    g = np.load(os.path.join(golden_dir, "unetr_tiny.npz"))
    _, mine = build_pair(pkg, "fp32")
    x, y = O.make_inputs(batch=2, img=32, n_classes=5, seed=3)
    enc4, logits = mine(x.to(DEV))
    loss = pkg.DiceCELoss(to_onehot_y=True, softmax=True)(logits, y.to(DEV))
    assert abs(loss.item() - float(g["loss"])) <= 1e-4
    assert np.allclose(enc4.detach().cpu().numpy(), g["enc4"], atol=2e-4)
    assert np.allclose(logits[:, :, :4, :4, :4].detach().cpu().numpy(), g["logits_corner"], atol=2e-4)
    assert (np.bincount(logits.argmax(1).flatten().cpu().numpy(), minlength=5) == g["argmax_hist"]).all()


@pytest.mark.parametrize("kw", [dict(pos_embed="conv", in_channels=2), dict(img_size=(32, 48, 16), out_channels=3)])
def test_variants_fp32(pkg, kw):
    ref, mine = build_pair(pkg, "fp32", **kw)
    cfg = {**TINY, **kw}
    g = torch.Generator().manual_seed(5)
    x = torch.rand(1, cfg["in_channels"], *cfg["img_size"], generator=g)
    ref64 = to64(ref)
    enc4_r, logits_r = ref(x)
    (logits_r.square().mean() + enc4_r.square().mean()).backward()
    e64, l64 = ref64(x.double())
    (l64.square().mean() + e64.square().mean()).backward()
    enc4, logits = mine(x.to(DEV))
    assert relerr(logits, logits_r) <= 1e-4 and relerr(enc4, enc4_r) <= 1e-4
    (logits.square().mean() + enc4.square().mean()).backward()
    grad_check(mine, ref, ref64, 1e-2)


def test_ranking_stage_gradient_reach(pkg):
    """grad=None (not zeros) for parameters the loss does not reach (rank:259-262)."""
    ref, mine = build_pair(pkg, "fp32")
    x = torch.rand(2, 1, 32, 32, 32, generator=torch.Generator().manual_seed(9))
    ref64 = to64(ref)
    ref(x)[0].square().sum().backward()
    ref64(x.double())[0].square().sum().backward()
    enc4, _ = mine(x.to(DEV))
    enc4.square().sum().backward()
    grad_check(mine, ref, ref64, 1e-2)
    assert mine.vit.blocks[10].mlp.linear1.weight.grad is None and mine.decoder5.transp_conv.conv.weight.grad is None
    for m in (ref, ref64, mine):
        m.zero_grad(set_to_none=True)
    ref(x, freeze_encoder=True)[1].square().mean().backward()
    ref64(x.double(), freeze_encoder=True)[1].square().mean().backward()
    _, logits = mine(x.to(DEV), freeze_encoder=True)
    logits.square().mean().backward()
    grad_check(mine, ref, ref64, 1e-2)
    assert mine.vit.blocks[0].attn.qkv.weight.grad is None and mine.out.conv.conv.bias.grad is not None


def test_monai_flavour_and_eval_no_grad(pkg):
    cfg = dict(TINY)
    torch.manual_seed(0)
    m = pkg.MonaiUNETR(**cfg).to(DEV).set_mode("fp32").eval()
    with torch.no_grad():
        out = m(torch.rand(3, 1, 32, 32, 32, device=DEV))
    assert isinstance(out, torch.Tensor) and out.shape == (3, 5, 32, 32, 32) and torch.isfinite(out).all()


# ------------------------------------------------------------------------------------------- config 1 (full ViT-B, 96^3)
@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_config1_full_size_forward_and_dice(pkg, mode, tol):
    """BASELINE.json configs[0]: logits within 1e-4 (fp32 mode) / 1e-2 (bf16 mode) of the fp32 reference arithmetic,
    DiceCE within 1e-3, identical argmax masks in fp32 mode (wherever the reference's own margin is decided)."""
    ref = O.make_model()
    mine = pkg.UNETR(1, 14, (96,) * 3, 16, 768, 3072, 12, "perceptron", "instance", res_block=True)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(DEV).set_mode(mode).eval()
    x, y = O.make_inputs()
    with torch.no_grad():
        enc4_r, logits_r = ref(x)
        loss_r, dice_r, _ = O.dice_ce_loss(logits_r, y, return_terms=True)
        enc4, logits = mine(x.to(DEV))
        lossfn = pkg.DiceCELoss(to_onehot_y=True, softmax=True)
        loss = lossfn(logits, y.to(DEV))
    e = relerr(logits, logits_r)
    top2 = logits_r.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1]).min().item()
    mism = (logits.argmax(1).cpu() != logits_r.argmax(1)).sum().item()
    print(f"[config1 {mode}] logits rel-err {e:.3e}  enc4 rel-err {relerr(enc4, enc4_r):.3e}  loss {loss.item():.6f} vs {loss_r.item():.6f}"
          f"  argmax mismatches {mism}/{logits_r[:, 0].numel()}  min top-2 margin {margin:.3e}")
    assert e <= tol
    assert abs(loss.item() - loss_r.item()) <= 1e-3
    if mode == "fp32":
        assert assert_argmax_parity(logits, logits_r) <= 8
        # arbitration by the fp64 oracle: wherever the mask differs from the fp64 mask, the fp64 logits of the two classes involved must
        # be closer than the fp32 rounding of the logits themselves (measured error of both fp32 implementations: ~2e-6 relative);
        # the fp32 CPU reference is held to the same rule, and the smallest margin either one decided correctly is printed
        with torch.no_grad():
            l64 = to64(ref)(x.double())[1]
        a64 = l64.argmax(1)
        scale = l64.abs().max().item()
        t2 = l64.topk(2, dim=1).values
        m64 = (t2[:, 0] - t2[:, 1])
        for tag, lg in (("b200 fp32 mode", logits.cpu()), ("fp32 CPU reference", logits_r)):
            wrong = lg.argmax(1) != a64
            worst_wrong = m64[wrong].max().item() if wrong.any() else 0.0
            min_right = m64[~wrong].min().item()
            print(f"[config1 argmax vs fp64] {tag}: {int(wrong.sum())} of {a64.numel()} voxels differ; largest fp64 margin among them "
                  f"{worst_wrong:.3e}; smallest fp64 margin decided correctly {min_right:.3e}  (logit scale {scale:.3f})")
            assert worst_wrong <= 8e-6 * scale, (tag, worst_wrong, scale)


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_config4_128_cube_forward_and_training_step(pkg, mode, tol):
    """BASELINE.json configs[3] geometry: 128^3 crops (L = 512 tokens: the attention runs as batched GEMMs + softmax, the decoder
    stages are 8..128 voxels wide).  Forward of one crop against the fp32 oracle at the north_star tolerances; then the per-GPU
    batch of the configuration (4 crops) runs a full training step in bf16 mode: finite loss and gradients, loss of the first crop's
    logits consistent with the batch-1 forward (samples are independent: InstanceNorm is per sample)."""
    ref = O.make_model(img=128)
    mine = pkg.UNETR(1, 14, (128,) * 3, 16, 768, 3072, 12, "perceptron", "instance", res_block=True)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(DEV).set_mode(mode)
    x, y = O.make_inputs(img=128)
    with torch.no_grad():
        enc4_r, logits_r = ref(x)
        loss_r = O.dice_ce_loss(logits_r, y)
        enc4, logits = mine.eval()(x.to(DEV))
        loss = pkg.DiceCELoss(to_onehot_y=True, softmax=True)(logits, y.to(DEV))
    e = relerr(logits, logits_r)
    print(f"[config4 {mode}] 128^3 logits rel-err {e:.3e}  enc4 rel-err {relerr(enc4, enc4_r):.3e}  loss {loss.item():.6f} vs {loss_r.item():.6f}")
    assert enc4.shape == (1, 128, 16, 16, 16) and logits.shape == (1, 14, 128, 128, 128)
    assert e <= tol and abs(loss.item() - loss_r.item()) <= 1e-3
    if mode == "bf16":
        x4, y4 = O.make_inputs(batch=4, img=128, seed=7)
        x4[0], y4[0] = x[0], y[0]
        mine.train()
        _, lg4 = mine(x4.to(DEV))
        l4 = pkg.DiceCELoss(to_onehot_y=True, softmax=True)(lg4, y4.to(DEV))
        l4.backward()
        torch.cuda.synchronize()
        assert torch.isfinite(l4).item() and all(torch.isfinite(p.grad).all().item() for p in mine.parameters() if p.grad is not None)
        assert all(p.grad is not None for n, p in mine.named_parameters() if "cls_token" not in n)
        # a batch of 4 is 4 independent crops; tile shapes / split-K depend on the batch, so bf16 roundings differ: same budget
        assert relerr(lg4[:1], logits) <= 1e-2


@pytest.mark.parametrize("mode,cos_min,cos_med", [("fp32", 0.9995, 0.99999), ("bf16", 0.97, 0.985)])
def test_config2_training_step_gradients(pkg, mode, cos_min, cos_med):
    """BASELINE.json configs[1]: batch 2, 96^3 -- fwd + DiceCE + bwd; per-tensor gradient cosine against the fp32 oracle.

    fp32 mode pins the backward kernels at full size (cosine ~1).  In bf16 mode the bound is set by the network, not by the
    kernels: this randomly initialised UNETR is chaotic in its gradients (LeakyReLU kinks + InstanceNorm) -- the fp32 ORACLE
    ITSELF, run with nothing but its weights and input rounded to bf16, lands at cosine 0.983-0.989 / relative L2 0.17 on
    every tensor upstream of the decoder (and 1.000 next to the loss).  The test measures that figure in the same run
    (`pert`) and requires the CUDA path to be no worse than it by more than 0.005 in median and 0.01 in minimum."""
    ref = O.make_model()
    mine = pkg.UNETR(1, 14, (96,) * 3, 16, 768, 3072, 12, "perceptron", "instance", res_block=True)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(DEV).set_mode(mode)
    x, y = O.make_inputs(batch=2)
    O.dice_ce_loss(ref(x)[1], y).backward()
    loss = pkg.DiceCELoss(to_onehot_y=True, softmax=True)(mine(x.to(DEV))[1], y.to(DEV))
    loss.backward()
    rows = []
    for (k, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
        assert (p.grad is None) == (q.grad is None), k
        if q.grad is not None:
            rows.append((cosine(p.grad, q.grad), l2err(p.grad, q.grad), q.grad.numel(), k))
    if os.environ.get("B200_DUMP_GRADS"):
        with open(os.environ["B200_DUMP_GRADS"] + "." + mode, "w") as f:
            for c, e, n, k in rows:
                f.write(f"{c:.5f} {e:.4f} {n:9d} {k}\n")
    rows.sort()
    big = [r for r in rows if r[2] >= 4096]
    med = sorted(r[0] for r in big)[len(big) // 2]
    if mode == "bf16":   # the unavoidable part: fp32 oracle with bf16-rounded weights and input
        pert = copy.deepcopy(ref)
        with torch.no_grad():
            for q in pert.parameters():
                q.grad = None
                q.copy_(q.to(torch.bfloat16).float())
        O.dice_ce_loss(pert(x.to(torch.bfloat16).float())[1], y).backward()
        pc = sorted(cosine(a.grad, b.grad) for a, b in zip(pert.parameters(), ref.parameters()) if b.grad is not None and b.grad.numel() >= 4096)
        print(f"[config2 bf16] oracle with bf16-rounded weights/input vs fp32 oracle: min cosine {pc[0]:.5f}, median {pc[len(pc) // 2]:.5f}")
        cos_min, cos_med = min(cos_min, pc[0] - 0.01), min(cos_med, pc[len(pc) // 2] - 0.005)
        assert min(r[0] for r in big) >= pc[0] - 0.01 and med >= pc[len(pc) // 2] - 0.005, (rows[0], pc[0], pc[len(pc) // 2])
    print(f"[config2 {mode}] worst cosines:", [(round(c, 5), round(e, 4), k) for c, e, _, k in rows[:4]])
    print(f"[config2 {mode}] weight tensors (>=4096 elems): min cosine %.5f, median %.5f, max relL2 %.4f" %
          (min(r[0] for r in big), med, max(r[1] for r in big)))
    assert min(r[0] for r in big) >= cos_min, rows[0]
    assert med >= cos_med
    near_loss = [r for r in rows if r[3].startswith(("out.conv", "decoder2.conv_block.conv2", "decoder2.conv_block.conv3"))]
    assert min(r[0] for r in near_loss) >= (0.9999 if mode == "fp32" else 0.995), near_loss


def test_inference_reuses_packed_weights_until_they_change(pkg):
    """no_grad calls keep one workspace and skip the fp32->bf16 weight re-pack while (storage, version) of every parameter
    is unchanged; an in-place update (optimizer step, load_state_dict) must invalidate the cache."""
    cfg = dict(in_channels=1, out_channels=3, img_size=(32, 32, 32), feature_size=16, hidden_size=64, mlp_dim=128, num_heads=4,
               pos_embed="perceptron", norm_name="instance", res_block=True)
    torch.manual_seed(3)
    net = pkg.MonaiUNETR(**cfg).to(DEV).set_mode("bf16").eval()
    x = torch.rand(2, 1, 32, 32, 32, device=DEV)
    with torch.no_grad():
        a = net(x).clone()
        b = net(x).clone()           # second call: packed weights reused
        assert torch.equal(a, b)
        for p in net.parameters():   # in-place change of every weight -> cache must be refreshed
            p.mul_(1.5)
        c = net(x).clone()
        fresh = pkg.MonaiUNETR(**cfg).to(DEV).set_mode("bf16").eval()
        fresh.load_state_dict(net.state_dict())
        d = fresh(x)
    assert not torch.equal(a, c)
    assert torch.equal(c, d)


def test_fused_adamw_matches_torch(pkg):
    """one-launch AdamW == torch.optim.AdamW (decoupled decay, bias correction), including parameters without a
    gradient (state untouched), odd sizes and gradients that are unaligned views of a flat buffer."""
    g = torch.Generator().manual_seed(0)
    shapes = [(3, 5, 7), (1031,), (64, 64), (14,), (40000,), (2, 3)]
    ref = [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes]
    mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    kw = dict(lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.05)
    o_ref, o_mine = torch.optim.AdamW(ref, **kw), pkg.FusedAdamW(mine, **kw)
    for it in range(5):
        flat = torch.randn(sum(p.numel() for p in ref) + 1, generator=g).to(DEV)
        off = 1                                              # odd offset -> 4-byte aligned views only
        for i, (a, b) in enumerate(zip(ref, mine)):
            if i == 5 and it < 2:                            # a parameter that wakes up at the third step
                a.grad = b.grad = None
            else:
                a.grad = flat[off:off + a.numel()].view_as(a).clone()
                b.grad = flat[off:off + a.numel()].view_as(a)
            off += a.numel()
        o_ref.step(); o_mine.step()
    for a, b in zip(ref, mine):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (a - b).abs().max()
    assert o_mine.state[mine[5]]["step"] == 3 and o_mine.state[mine[0]]["step"] == 5
    v0 = mine[0]._version
    o_mine.step()
    assert mine[0]._version > v0          # in-place update is visible to version-keyed caches (UNETR inference workspace)


# ------------------------------------------------------------------------------------------- losses
def test_dicece_matches_oracle_and_closed_form(pkg):
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(2, 14, 24, 20, 16, generator=g)
    y = torch.randint(0, 14, (2, 1, 24, 20, 16), generator=g).float()
    lr = logits.clone().requires_grad_(True)
    loss_r = O.dice_ce_loss(lr, y)
    (loss_r * 1.7).backward()
    lg = logits.to(DEV).requires_grad_(True)
    loss = pkg.DiceCELoss(to_onehot_y=True, softmax=True)(lg, y.to(DEV))
    (loss * 1.7).backward()
    assert abs(loss.item() - loss_r.item()) <= 1e-5
    assert relerr(lg.grad, lr.grad) <= 1e-4
    # T1: all-zero logits
    z = pkg.DiceCELoss(to_onehot_y=True, softmax=True)(torch.zeros(2, 14, 24, 20, 16, device=DEV), y.to(DEV)).item()
    n = 24 * 20 * 16
    counts = torch.stack([(y[i] == k).sum() for i in range(2) for k in range(14)]).double()
    closed = math.log(14) + (1 - (2 * counts / 14 + 1e-5) / (counts + n / 14 + 1e-5)).mean().item()
    assert abs(z - closed) <= 1e-5
    with pytest.raises(AssertionError):
        pkg.DiceCELoss(to_onehot_y=True, softmax=True)(lg, y.to(DEV)[:, :, :-1])


@pytest.mark.parametrize("name", ["feat", "recon", "feat_T05"])
@pytest.mark.parametrize("sd", [2, 3, 4])
def test_ranking_loss_matches_reference_golden(pkg, golden_dir, name, sd):
    """Golden vectors were produced by the reference's own extract_triplets_more_partitions + BTLoss."""
    g = np.load(os.path.join(golden_dir, f"ranking_{name}.npz"))
    feat = torch.from_numpy(g["feat"]).to(DEV).requires_grad_(True)
    pkg.configure_ranking(temperature=float(g["temperature"]))
    f1, f2 = torch.split(feat, [2, 2], dim=0)
    np.random.seed(int(g[f"npseed_sd{sd}"]))
    ref, sim, dis = pkg.extract_triplets_more_partitions(f1, f2, sd)
    assert ref.plan[3] == g[f"idx_sd{sd}"].tolist()

    class Opt:
        steps = 0
        def step(self): Opt.steps += 1
        def zero_grad(self): pass
    val = pkg.BTLoss(ref, sim, dis, Opt())
    assert isinstance(val, float) and Opt.steps == 1
    want = float(g[f"loss_sd{sd}"])
    assert abs(val - want) <= 2e-4 * abs(want), (val, want)
    assert relerr(feat.grad, torch.from_numpy(g[f"grad_sd{sd}"])) <= 1e-3
    nz = (feat.grad != 0).float().mean().item()
    assert abs(nz - 4 / feat.shape[sd]) < 1e-6     # T3: only the 4 selected planes receive gradient
    pkg.configure_ranking(temperature=0.1)


def test_ranking_identical_slices_is_576_ln2(pkg):
    feat = torch.ones(4, 4, 8, 8, 8, device=DEV)
    v = pkg.ranking_loss(feat[:2], feat[2:], 2, [0, 2, 4, 6], 0.1).item()
    assert abs(v - 576 * math.log(2)) < 1e-2


# ------------------------------------------------------------------------------------------- sliding window
@pytest.mark.parametrize("overlap,shape,roi", [(0.25, (40, 33, 50), (16,) * 3), (0.5, (48, 48, 32), (16,) * 3),
                                               (0.8, (20, 16, 16), (16,) * 3), (0.25, (10, 20, 12), (16,) * 3)])
def test_sliding_window_identity_and_oracle(pkg, overlap, shape, roi):
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 1, *shape, generator=g)
    y = pkg.sliding_window_inference(x.to(DEV), roi, 4, lambda w: w + 1, overlap=overlap)
    assert torch.allclose(y.cpu(), x + 1, atol=1e-6)          # T6
    w = torch.randn(3, 1, 3, 3, 3, generator=g)
    f_cpu = lambda t: torch.nn.functional.conv3d(t, w, padding=1)
    f_gpu = lambda t: torch.nn.functional.conv3d(t, w.to(DEV), padding=1)
    want = O.sliding_window_inference(x, roi, 4, f_cpu, overlap=overlap)
    got, mask = pkg.sliding_window_inference(x.to(DEV), roi, 4, f_gpu, overlap=overlap, return_argmax=True)
    assert got.shape == want.shape and torch.allclose(got.cpu(), want, atol=1e-5)
    assert (mask[:, 0].cpu().long() == got.argmax(1).cpu()).all()


@pytest.mark.parametrize("rounds", [False, True])
@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("shape,roi,overlap", [((48, 40, 32), (16, 16, 16), 0.5), ((40, 20, 24), (16, 16, 16), 0.25), ((12, 40, 40), (16, 16, 16), 0.5)])
def test_sliding_window_slabs_equal_single_bitwise(pkg, world, shape, roi, overlap, rounds):
    """The slab-owned multi-GPU path, every rank run in turn on one GPU with the NCCL exchange replaced by a hand-over of the packed
    halo buffers: each rank predicts its window chunk, packs the row-clipped pieces for the higher ranks, accumulates its slab in
    global window order and normalises it.  Stitched together, logits and argmax mask must equal the single-GPU call BIT FOR BIT
    (same per-voxel addition order), including a volume smaller than the window along the sharded axis (padding rows)."""
    import importlib
    inf = importlib.import_module("3dmedicalimagesegmentation_b200.inferers")
    lib = pkg._lib.load()
    g = torch.Generator().manual_seed(sum(shape) + world)
    x = torch.randn(1, 2, *shape, generator=g).to(DEV)
    wts = torch.tensor([[0.7, -1.3], [1.9, 0.4], [-0.6, 0.8]], device=DEV)

    def f(t):      # 2 -> 3 channels, not window-position independent after fp32 rounding of the overlap sums
        return torch.einsum("kc,bcxyz->bkxyz", wts, t) + 0.1 * t[:, :1] * t[:, 1:2]

    lab = torch.randint(0, 3, (1, 1, *shape), generator=g).float().to(DEV)
    want, want_mask, want_counts = pkg.sliding_window_inference(x, roi, 4, f, overlap=overlap, return_argmax=True, labels=lab)
    size = tuple(max(o, r) for o, r in zip(shape, roi))
    pad = tuple((s_ - o) // 2 for s_, o in zip(size, shape))
    per_axis, flat = inf.window_starts(size, roi, overlap)
    if len(flat) < world:
        pytest.skip("fewer windows than ranks")
    ranks = [inf._SlabItem(lib, x, 0, flat, roi, size, shape, pad, 2, r, world, 0.0) for r in range(world)]
    sends = []
    for it in ranks:
        it.predict(f, 4, (), {})
        # rounds: the pipelined form of the exchange (pieces leave in ROUNDS parts while the prediction loop runs)
        sends.append([it.pack_sends(j) for j in range(it.ROUNDS)] if rounds else it.pack_sends())
    cout = ranks[0].cout
    got = torch.full_like(want, float("nan"))
    mask = torch.zeros_like(want_mask)
    counts = torch.zeros_like(want_counts)
    covered = 0
    for r, it in enumerate(ranks):
        if rounds:
            recv = {}
            for j in range(it.ROUNDS):
                for src, n in it.recv_sizes(cout, j).items():
                    recv[(src, j)] = sends[src][j][r]
                    assert recv[(src, j)].numel() == n
        else:
            sizes = it.recv_sizes(cout)
            recv = {src: sends[src][r] for src in sizes}
            assert all(recv[src].numel() == n for src, n in sizes.items())
        acc = it.accumulate(recv)
        d0, d1 = it.rows()
        covered += d1 - d0
        it.finalize(acc, per_axis, got[0], 0, shape[0], mask, lab, counts)
    torch.cuda.synchronize()
    assert covered == shape[0]
    assert torch.equal(got, want), (got - want).abs().max().item()
    assert torch.equal(mask, want_mask) and torch.equal(counts, want_counts)


def test_unetr_as_sliding_window_predictor(pkg):
    m = pkg.MonaiUNETR(**TINY).to(DEV).set_mode("fp32").eval()
    x = torch.rand(1, 1, 48, 32, 40, device=DEV)
    with torch.no_grad():
        out = pkg.sliding_window_inference(x, (32, 32, 32), 4, m, overlap=0.5)
    assert out.shape == (1, 5, 48, 32, 40) and torch.isfinite(out).all()


# ------------------------------------------------------------------------------------------- sigmoid DiceCE
@pytest.mark.parametrize("shape,chan", [((24, 20, 16), 4), ((9, 7, 5), 3), ((16, 16, 16), 6)])
def test_sigmoid_dicece_matches_oracle_and_closed_form(pkg, shape, chan):
    g = torch.Generator().manual_seed(7)
    logits = torch.randn(2, chan, *shape, generator=g) * 2
    if chan == 4:
        target = O.brats_multichannel(torch.randint(0, 4, (2, 1, *shape), generator=g))     # seg:65-93 multi-hot (ties in argmax)
    else:
        target = (torch.rand(2, chan, *shape, generator=g) > 0.6).float()
    lr = logits.clone().requires_grad_(True)
    loss_r, dice_r, ce_r = O.dice_ce_loss_sigmoid(lr, target, return_terms=True)
    (loss_r * 0.6).backward()
    fn = pkg.DiceCELoss(to_onehot_y=False, sigmoid=True)
    lg = logits.to(DEV).requires_grad_(True)
    loss = fn(lg, target.to(DEV))
    (loss * 0.6).backward()
    assert abs(loss.item() - loss_r.item()) <= 1e-5                      # bar: 1e-3 (north_star), measured ~1e-6
    assert relerr(lg.grad, lr.grad) <= 1e-4
    # T8: zero logits -> Dice term closed form with p = 1/2, CE = ln C
    z = fn(torch.zeros(2, chan, *shape, device=DEV), target.to(DEV)).item()
    n = shape[0] * shape[1] * shape[2]
    gs = target.sum((2, 3, 4)).double()
    closed = math.log(chan) + (1 - (gs + 1e-5) / (gs + n / 2 + 1e-5)).mean().item()
    assert abs(z - closed) <= 1e-5
    with pytest.raises(AssertionError):
        fn(lg, target.to(DEV)[:, :1])
    with pytest.raises(NotImplementedError):
        pkg.DiceCELoss(to_onehot_y=True, sigmoid=True)


def test_sigmoid_dicece_task01_size(pkg):
    """seg:480 configuration at its real size (4 channels, 128^3 crops, batch 2): loss within 1e-3 of the oracle, and the
    size-independent property d(loss)/d(logits) sums: CE part sums to 0 over channels at every voxel."""
    g = torch.Generator().manual_seed(8)
    logits = torch.randn(2, 4, 128, 128, 128, generator=g)
    target = O.brats_multichannel(torch.randint(0, 4, (2, 1, 128, 128, 128), generator=g))
    want = O.dice_ce_loss_sigmoid(logits, target).item()
    lg = logits.to(DEV).requires_grad_(True)
    loss = pkg.DiceCELoss(to_onehot_y=False, sigmoid=True)(lg, target.to(DEV))
    loss.backward()
    assert abs(loss.item() - want) <= 1e-4
    assert torch.isfinite(lg.grad).all()


# ------------------------------------------------------------------------------------------- validation metrics
def _onehot(idx, c):
    return torch.zeros(idx.shape[0], c, *idx.shape[2:]).scatter_(1, idx.long(), 1.0)


def test_dice_and_confusion_metrics_match_oracle(pkg):
    g = torch.Generator().manual_seed(9)
    c, shape = 6, (20, 18, 10)
    lab = torch.randint(0, 4, (3, 1, *shape), generator=g)            # classes 4, 5 never appear in the label -> NaN Dice
    lab[2] = 0                                                         # sample 2: background only
    pred = torch.randint(0, c, (3, 1, *shape), generator=g)
    y, p = _onehot(lab, c), _onehot(pred, c)
    d_want = O.dice_metric(p, y)
    cm_want = O.confusion_matrix(p, y)
    for red in ("mean", "mean_batch"):
        dm = pkg.DiceMetric(include_background=True, reduction=red, get_not_nans=False)
        # the reference's call form (seg:112-121): lists of channel-first one-hot tensors, one volume per call, cumulative buffer
        for i in range(3):
            rows = dm(y_pred=[p[i].to(DEV)], y=[y[i].to(DEV)])
            assert torch.allclose(rows.cpu(), d_want[i:i + 1], atol=1e-6, equal_nan=True)
            agg = dm.aggregate()
            want, _ = O.metric_reduce(d_want[:i + 1], red)
            assert agg.shape == want.shape and torch.allclose(agg.cpu(), want, atol=1e-6), (red, i)
        assert agg.shape == ((1,) if red == "mean" else (c,)) and isinstance(agg.sum().item(), float)
        dm.reset()
        with pytest.raises(ValueError):
            dm.aggregate()
        for name in ("precision", "sensitivity"):
            for cs in (False, True):
                m = pkg.ConfusionMatrixMetric(include_background=True, metric_name=name, reduction=red, compute_sample=cs)
                rows = m(y_pred=p.to(DEV), y=y.to(DEV))
                assert torch.equal(rows.cpu(), cm_want)                # integer counts: exact
                got = m.aggregate()[0]
                want = O.confusion_aggregate(cm_want, name, red, compute_sample=cs)
                assert got.shape == want.shape and torch.allclose(got.cpu(), want, atol=1e-6, equal_nan=True), (name, red, cs)
    # label-map form == one-hot form, bit for bit (integer counts)
    a = pkg.segmentation_counts(p.to(DEV), y.to(DEV))
    b = pkg.segmentation_counts_from_label_maps(pred.to(torch.uint8).to(DEV), lab.float().to(DEV), c)
    assert torch.equal(a, b)
    assert a[..., 1].sum().item() == 3 * 20 * 18 * 10 and a[..., 2].sum().item() == 3 * 20 * 18 * 10


def test_sliding_window_fused_validation_tail(pkg):
    """seg:103-126 in one call: sliding-window logits -> argmax -> Dice against the label, counts formed inside the normalise
    pass; equal to the reference's route (one-hot tensors -> DiceMetric) on the same logits."""
    g = torch.Generator().manual_seed(10)
    x = torch.rand(2, 1, 40, 33, 50, generator=g)
    lab = torch.randint(0, 3, (2, 1, 40, 33, 50), generator=g).float()
    w = torch.randn(5, 1, 3, 3, 3, generator=g)
    f_gpu = lambda t: torch.nn.functional.conv3d(t, w.to(DEV), padding=1)
    logits, mask, counts = pkg.sliding_window_inference(x.to(DEV), (16,) * 3, 4, f_gpu, overlap=0.25, labels=lab.to(DEV))
    plain = pkg.sliding_window_inference(x.to(DEV), (16,) * 3, 4, f_gpu, overlap=0.25)
    assert torch.equal(logits, plain)
    assert (mask[:, 0].long() == logits.argmax(1)).all()
    p, y = _onehot(mask.cpu(), 5), _onehot(lab, 5)
    dm = pkg.DiceMetric(include_background=True, reduction="mean_batch")
    dm.update_from_counts(counts)
    assert torch.allclose(dm.aggregate().cpu(), O.metric_reduce(O.dice_metric(p, y), "mean_batch")[0], atol=1e-6)
    assert torch.equal(counts, pkg.segmentation_counts(p.to(DEV), y.to(DEV)))
    none, mask2, counts2 = pkg.sliding_window_inference(x.to(DEV), (16,) * 3, 4, f_gpu, overlap=0.25, labels=lab.to(DEV),
                                                        return_logits=False)
    assert none is None and torch.equal(mask2, mask) and torch.equal(counts2, counts)


def test_multi_window_accumulate_is_bit_identical_to_window_loop(pkg):
    """b200_sw_accumulate_n (one launch per predictor call) against the one-window-per-launch loop it replaces: identical bits,
    because every voxel adds its windows in window order (MONAI's accumulation order, Appendix B.9)."""
    import ctypes
    L_ = pkg._lib; lib = L_.load()
    C, roi, size = 3, (16, 12, 20), (40, 30, 48)
    g = L_.SwGeom(C, *size, 0, 0, 0, *size, *roi)
    gen = torch.Generator().manual_seed(11)
    for starts in ([(0, 0, 0, 0), (0, 0, 0, 10), (0, 0, 0, 20), (0, 0, 0, 28)],          # a run along z, 50 % overlaps
                   [(0, 8, 18, 28), (0, 16, 0, 0), (0, 16, 0, 10)],                      # wraps to the next row of windows
                   [(0, 24, 18, 28)], [(0, 0, 0, 0), (0, 0, 0, 0)]):                     # single window; the same window twice
        n = len(starts)
        pred = torch.randn(n, C, *roi, generator=gen).to(DEV)
        base = torch.randn(1, C, *size, generator=gen).to(DEV)
        a, b = base.clone(), base.clone()
        for k, it in enumerate(starts):
            s4 = (ctypes.c_int32 * 4)(*it)
            L_.check(lib.b200_sw_accumulate(L_.ptr(a), L_.ptr(pred[k]), ctypes.byref(g), s4, L_.stream_ptr()), "acc")
        sN = (ctypes.c_int32 * (4 * n))(*[v for it in starts for v in it])
        L_.check(lib.b200_sw_accumulate_n(L_.ptr(b), L_.ptr(pred), ctypes.byref(g), sN, n, L_.stream_ptr()), "acc_n")
        assert torch.equal(a, b)
        assert not torch.equal(a, base)


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_task01_configuration_four_channel_128_cube_sigmoid_loss(pkg, mode, tol):
    """The second dataset configuration the segmentation script runs (seg:408-482, 501-513): 4 MR channels in,
    4 multi-hot BraTS channels out, 128^3 crops -- perceptron patch embedding with K = 4 * 4096, encoder1 on 4 input channels,
    DiceCELoss(to_onehot_y=False, sigmoid=True).  Logits and loss against the oracle at the north_star tolerances; one backward."""
    ref = O.make_model(img=128, in_channels=4, out_channels=4)
    mine = pkg.UNETR(4, 4, (128,) * 3, 16, 768, 3072, 12, "perceptron", "instance", res_block=True)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(DEV).set_mode(mode)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(1, 4, 128, 128, 128, generator=g)                       # NormalizeIntensityd output: zero mean, unit variance (seg:476)
    target = O.brats_multichannel(torch.randint(0, 4, (1, 1, 128, 128, 128), generator=g))
    with torch.no_grad():
        _, logits_r = ref(x)
        loss_r = O.dice_ce_loss_sigmoid(logits_r, target)
    _, logits = mine(x.to(DEV))
    loss = pkg.DiceCELoss(to_onehot_y=False, sigmoid=True)(logits, target.to(DEV))
    e = relerr(logits, logits_r)
    print(f"[task01 {mode}] logits rel-err {e:.3e}  sigmoid DiceCE {loss.item():.6f} vs {loss_r.item():.6f}")
    assert e <= tol and abs(loss.item() - loss_r.item()) <= 1e-3
    loss.backward()
    torch.cuda.synchronize()
    assert all(torch.isfinite(p.grad).all().item() for n, p in mine.named_parameters() if "cls_token" not in n)


def test_fused_1x1_weight_gradient_rides_on_the_halo_kernel(pkg):
    """decoder2's residual block: the 1^3 convolution's weight gradient is formed by the 3^3 halo weight-gradient launch (tenth TMEM
    accumulator, centre tap).  bf16-mode gradients of both convolutions against the fp32-mode kernels of the same network: the
    layers sit next to the loss, where bf16 gradients agree to cosine > 0.995; with the fusion switched off the same numbers come out."""
    cfg = dict(in_channels=1, out_channels=5, img_size=(48, 48, 48), feature_size=16, hidden_size=128, mlp_dim=256, num_heads=2,
               pos_embed="perceptron", norm_name="instance", res_block=True)
    torch.manual_seed(3)
    base = pkg.UNETR(**cfg)
    x, y = O.make_inputs(batch=2, img=48, n_classes=5, seed=5)
    grads = {}
    for mode in ("fp32", "bf16"):
        net = pkg.UNETR(**cfg)
        net.load_state_dict(base.state_dict())
        net = net.to(DEV).set_mode(mode)
        _, logits = net(x.to(DEV))
        pkg.DiceCELoss(to_onehot_y=True, softmax=True)(logits, y.to(DEV)).backward()
        grads[mode] = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
    for k in ("decoder2.conv_block.conv1.conv.weight", "decoder2.conv_block.conv3.conv.weight", "decoder2.conv_block.conv2.conv.weight"):
        c = cosine(grads["bf16"][k], grads["fp32"][k])
        n = (grads["bf16"][k].norm() / grads["fp32"][k].norm()).item()
        assert c >= 0.995 and 0.95 <= n <= 1.05, (k, c, n)
