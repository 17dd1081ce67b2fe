"""Runtime behaviour of the module around the kernels (B200): packed bf16 weight mirrors kept by FusedAdamW, persistent flat gradient
buffer under gradient accumulation, deepcopy / pickling of a module that owns C handles."""
import copy
import io

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# a network whose shapes take the tcgen05 engine end to end (16..128-channel convolutions, ViT GEMMs, fused attention) so that every
# mirror kind (plain cast, transposed-conv tap-major, conv forward + dgrad layouts) is exercised
CFG = dict(in_channels=1, out_channels=5, img_size=(48, 48, 48), feature_size=16, hidden_size=128, mlp_dim=256, num_heads=2,
           pos_embed="perceptron", norm_name="instance", res_block=True)


def make(pkg, seed=0, mode="bf16"):
    torch.manual_seed(seed)
    return pkg.UNETR(**CFG).to(DEV).set_mode(mode)


def data(seed=1, batch=2):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 1, 48, 48, 48, generator=g).to(DEV)
    y = torch.randint(0, 5, (batch, 1, 48, 48, 48), generator=g).float().to(DEV)
    return x, y


def test_adamw_mirrors_equal_a_fresh_pack(pkg):
    """FusedAdamW(mirror=model) writes the packed bf16 copies itself; after 3 steps the next forward (which skips the re-pack) must
    equal, bit for bit, the forward of a fresh module loaded with the same fp32 weights (which packs from scratch)."""
    model = make(pkg)
    loss_fn = pkg.DiceCELoss(to_onehot_y=True, softmax=True)
    opt = pkg.FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-2, mirror=model)
    x, y = data()
    lib = pkg._lib.load()
    for i in range(3):
        _, logits = model(x)
        loss_fn(logits, y).backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
    key = model._version_key(model._ordered_params())
    assert model._packed_key[torch.device(DEV)] == key, "the optimizer did not mark the mirrors current"
    l0 = lib.b200_launch_count()
    with torch.no_grad():
        _, got = model(x)
    n_mirrored = lib.b200_launch_count() - l0
    fresh = make(pkg, seed=5)
    fresh.load_state_dict(model.state_dict())
    l0 = lib.b200_launch_count()
    with torch.no_grad():
        _, want = fresh(x)
    n_fresh = lib.b200_launch_count() - l0
    assert n_fresh == n_mirrored + 2, (n_fresh, n_mirrored)      # the cast + re-layout launches are what the mirror saves
    assert torch.equal(got, want)


def test_mirror_is_not_trusted_after_an_outside_update(pkg):
    """A parameter changed behind the optimizer's back (load_state_dict, manual edit) bumps its version: the next forward re-packs and
    the optimizer stops writing mirrors until a forward has verified them again."""
    model = make(pkg)
    loss_fn = pkg.DiceCELoss(to_onehot_y=True, softmax=True)
    opt = pkg.FusedAdamW(model.parameters(), lr=1e-3, mirror=model)
    x, y = data()
    _, logits = model(x)
    loss_fn(logits, y).backward()
    with torch.no_grad():
        model.out.conv.conv.bias.add_(0.5)                         # outside update between forward and step
        model.vit.blocks[0].mlp.linear1.weight.mul_(1.5)
    opt.step()
    opt.zero_grad(set_to_none=True)
    assert model._packed_key.get(torch.device(DEV)) != model._version_key(model._ordered_params())
    with torch.no_grad():
        _, got = model(x)
    fresh = make(pkg, seed=7)
    fresh.load_state_dict(model.state_dict())
    with torch.no_grad():
        _, want = fresh(x)
    assert torch.equal(got, want)


def test_gradient_accumulation_with_persistent_flat_buffer(pkg):
    """zero_grad(set_to_none=False) + two backward passes: the second backward must not overwrite the gradients the parameters hold
    (the flat buffer is reused only when no parameter holds a gradient).  fp32 mode: accumulation of two identical passes = 2x."""
    model = make(pkg, mode="fp32")
    loss_fn = pkg.DiceCELoss(to_onehot_y=True, softmax=True)
    x, y = data(batch=1)
    _, logits = model(x)
    loss_fn(logits, y).backward()
    single = [p.grad.clone() for p in model.parameters() if p.grad is not None]
    _, logits = model(x)
    loss_fn(logits, y).backward()                                   # accumulates into the held gradients
    double = [p.grad for p in model.parameters() if p.grad is not None]
    for a, b in zip(single, double):
        assert torch.allclose(b, 2 * a, rtol=1e-3, atol=1e-6 * a.abs().max().item() + 1e-12)
    model.zero_grad(set_to_none=False)
    _, logits = model(x)
    loss_fn(logits, y).backward()
    for a, p in zip(single, [p for p in model.parameters() if p.grad is not None]):
        assert torch.allclose(p.grad, a, rtol=1e-3, atol=1e-6 * a.abs().max().item() + 1e-12)


def test_deepcopy_and_pickle_do_not_share_handles(pkg):
    """copy.deepcopy / torch.save of a module that has already run (it owns C handles, a packed buffer, cached workspaces): the copy
    must rebuild its own runtime state, produce the same output, and both must be destructible."""
    model = make(pkg)
    x, _ = data()
    with torch.no_grad():
        _, want = model(x)
    clone = copy.deepcopy(model)
    assert clone._handles == {} and clone._packed == {}
    with torch.no_grad():
        _, got = clone(x)
    assert torch.equal(got, want)
    buf = io.BytesIO()
    torch.save(model, buf)
    buf.seek(0)
    loaded = torch.load(buf, weights_only=False)
    with torch.no_grad():
        _, got2 = loaded(x)
    assert torch.equal(got2, want)
    del clone, loaded
    with torch.no_grad():
        _, again = model(x)
    assert torch.equal(again, want)


def test_graphed_train_step_matches_eager(pkg):
    """GraphedTrainStep: 2 warm-up steps + 3 replays against 5 eager steps on the same batches (bf16 mode, FusedAdamW with mirrors and
    the device-side update count).  The losses of every step must agree (the kernels are the same; only fp32-atomic ordering in the
    backward differs run to run), the update counts must be 5 on both sides, and the replayed model's packed weights must equal a
    fresh pack of its fp32 parameters."""
    loss_fn = pkg.DiceCELoss(to_onehot_y=True, softmax=True)
    batches = [data(seed=10 + i) for i in range(5)]

    eager = make(pkg)
    opt_e = pkg.FusedAdamW(eager.parameters(), lr=1e-3, weight_decay=1e-2, mirror=eager)
    want = []
    for x, y in batches:
        _, logits = eager(x)
        loss = loss_fn(logits, y)
        loss.backward()
        opt_e.step()
        opt_e.zero_grad(set_to_none=True)
        want.append(loss.item())

    model = make(pkg)
    opt = pkg.FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-2, mirror=model, capturable=True)
    # steps 1-2 eager, step 3 = the constructor's warm-up step (a real optimisation step on the construction batch), steps 4-5 replays
    got = []
    step = None
    for i, (x, y) in enumerate(batches):
        if i < 2:
            _, logits = model(x)
            loss = loss_fn(logits, y)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            got.append(loss.item())
        elif i == 2:
            del loss, logits, _   # (enc4 in `_` too) a live autograd graph from an earlier stream keeps AccumulateGrad nodes that break the capture
            step = pkg.GraphedTrainStep(model, loss_fn, opt, x, y, warmup=1)
            got.append(want[2])
        else:
            got.append(step(x, y).item())
    assert step.launches_per_step > 100
    for a, b in zip(got, want):
        assert abs(a - b) <= 2e-3 * max(1.0, abs(b)), (got, want)
    p0 = next(iter(model.parameters()))
    assert opt.state[p0]["step"] == 5 and all(int(ds[0].item()) == 5 for ds in opt._dev_steps.values())
    with torch.no_grad():
        _, out = model(batches[0][0])
    fresh = make(pkg, seed=3)
    fresh.load_state_dict(model.state_dict())
    with torch.no_grad():
        _, ref = fresh(batches[0][0])
    assert torch.equal(out, ref)


def test_mode_switch_repacks(pkg):
    """fp32-mode calls must not mark the (not yet existing) packed bf16 weights as current: a module that ran in fp32 mode and is then
    switched to bf16 has to pack on its first bf16 forward."""
    model = make(pkg, mode="fp32")
    x, _ = data()
    with torch.no_grad():
        model(x)
        model.set_mode("bf16")
        _, got = model(x)
    fresh = make(pkg, seed=11)
    fresh.load_state_dict(model.state_dict())
    with torch.no_grad():
        _, want = fresh(x)
    assert torch.isfinite(got).all() and torch.equal(got, want)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_deferred_conv_weight_gradients_match_the_in_place_order(pkg, mode):
    """With gradient-ready events set (data-parallel overlap) the conv-stack weight gradients are launched after the ViT backward from
    per-block copies of their output gradients (exec.cuh: defer_wg) and the conv range is handed to the reducer last: same gradients
    as the in-place order (fp32 atomics: order-of-summation noise only), ranges still a partition of the flat buffer."""
    model = make(pkg, mode=mode)
    loss_fn = pkg.DiceCELoss(to_onehot_y=True, softmax=True)
    x, y = data()

    def grads():
        model.zero_grad(set_to_none=True)
        _, logits = model(x)
        loss_fn(logits, y).backward()
        torch.cuda.synchronize()
        return {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}

    want = grads()
    model.overlap_grad_reduce = True
    model.defer_conv_wgrads = True
    got = grads()
    ready = model._grad_ready
    assert ready is not None
    flat, ranges, views = ready
    assert ranges[-1][2] == flat.numel() and ranges[-1][1] > 0, "the conv range is reduced last"
    covered = sorted((lo, hi) for _, lo, hi in ranges)
    assert covered[0][0] == 0 and all(a[1] == b[0] for a, b in zip(covered, covered[1:])) and covered[-1][1] == flat.numel()
    assert all(ev.query() for ev, _, _ in ranges)
    assert got.keys() == want.keys()
    for n in want:
        torch.testing.assert_close(got[n], want[n], rtol=1e-3, atol=1e-5 * float(want[n].abs().max()) + 1e-12, msg=lambda m: f"{n}: {m}")
    # events set, in-place order (the default): conv range first
    model.defer_conv_wgrads = False
    inplace = grads()
    assert model._grad_ready[1][0][2] == flat.numel(), "in-place order: the conv range is reduced first"
    for n in want:
        torch.testing.assert_close(inplace[n], want[n], rtol=1e-3, atol=1e-5 * float(want[n].abs().max()) + 1e-12, msg=lambda m: f"{n}: {m}")
    model.overlap_grad_reduce = False
    again = grads()
    for n in want:
        torch.testing.assert_close(again[n], want[n], rtol=1e-3, atol=1e-5 * float(want[n].abs().max()) + 1e-12, msg=lambda m: f"{n}: {m}")
