"""tcgen05 implicit-GEMM conv3d (forward and dgrad) against torch conv3d on the same bf16-rounded operands.
Floating-point kernel: torch fp32 is the reference; tolerance 2^-8 relative (bf16 output rounding) of max|ref|."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run_conv(pkg, x_cl, in_coff, Ci, w, Co, ks, out_cl, out_coff, accumulate, dgrad, stats):
    lib = pkg._lib.load()
    N, D, H, W, in_pitch = x_cl.shape
    scratch = torch.empty(2 * w.numel(), dtype=torch.bfloat16, device=DEV)
    pkg._lib.check(lib.b200_test_tc_conv(pkg._lib.ptr(x_cl), in_pitch, in_coff, Ci, N, D, H, W, pkg._lib.ptr(w), Co, ks,
                                          pkg._lib.ptr(out_cl), out_cl.shape[-1], out_coff, accumulate, dgrad,
                                          pkg._lib.ptr(stats), pkg._lib.ptr(scratch), pkg._lib.stream_ptr()), "tc_conv")
    torch.cuda.synchronize()


@pytest.mark.parametrize("Ci,Co", [(16, 16), (32, 16), (64, 32), (128, 64), (256, 128), (16, 32)])
@pytest.mark.parametrize("ks", [3, 1])
@pytest.mark.parametrize("dims", [(12, 12, 12), (8, 20, 24)])
def test_conv_forward_with_stats(pkg, Ci, Co, ks, dims):
    g = torch.Generator().manual_seed(Ci * 7 + Co + ks)
    N = 2
    x = torch.randn(N, Ci, *dims, generator=g).to(torch.bfloat16)
    w = (torch.randn(Co, Ci, ks, ks, ks, generator=g) / (Ci * ks ** 3) ** 0.5)
    want = F.conv3d(x.float(), w.to(torch.bfloat16).float(), padding=ks // 2)
    pitch_in, coff_in = Ci + 16, 8          # exercise channel windows (concat buffers)
    x_cl = torch.zeros(N, *dims, pitch_in, dtype=torch.bfloat16)
    x_cl[..., coff_in:coff_in + Ci] = x.permute(0, 2, 3, 4, 1)
    out = torch.full((N, *dims, Co + 8), 7.0, dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(N, Co, 2, dtype=torch.float64, device=DEV)
    run_conv(pkg, x_cl.to(DEV), coff_in, Ci, w.to(DEV), Co, ks, out, 8, 0, 0, stats)
    got = out[..., 8:].float().permute(0, 4, 1, 2, 3).cpu()
    err = ((got - want).abs().max() / want.abs().max()).item()
    assert err <= 2 ** -7, err
    assert (out[..., :8] == 7.0).all()      # untouched channels of the destination
    s = stats.cpu()
    assert torch.allclose(s[..., 0], want.double().sum((2, 3, 4)), rtol=1e-3, atol=1e-2 * want.abs().max().item())
    assert torch.allclose(s[..., 1], want.double().square().sum((2, 3, 4)), rtol=1e-3)


@pytest.mark.parametrize("Ci,Co", [(32, 16), (16, 16), (256, 128)])
@pytest.mark.parametrize("ks", [3, 1])
def test_conv_dgrad_and_accumulate(pkg, Ci, Co, ks):
    g = torch.Generator().manual_seed(Ci + Co * 3 + ks)
    dims = (8, 12, 16)
    dy = torch.randn(1, Co, *dims, generator=g).to(torch.bfloat16)
    w = torch.randn(Co, Ci, ks, ks, ks, generator=g) / (Co * ks ** 3) ** 0.5
    want = F.conv_transpose3d(dy.float(), w.to(torch.bfloat16).float(), padding=ks // 2)
    dy_cl = dy.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    base = torch.randn(1, *dims, Ci, generator=g).to(torch.bfloat16)
    out = base.clone().to(DEV)
    run_conv(pkg, dy_cl, 0, Ci, w.to(DEV), Co, ks, out, 0, 1, 1, None)
    got = out.float().permute(0, 4, 1, 2, 3).cpu()
    want_acc = want + base.float().permute(0, 4, 1, 2, 3)
    err = ((got - want_acc).abs().max() / want_acc.abs().max()).item()
    assert err <= 2 ** -7, err


@pytest.mark.parametrize("Ci,Co", [(16, 16), (32, 16), (32, 32), (64, 32), (64, 64), (128, 64), (128, 128), (256, 128)])
@pytest.mark.parametrize("ks", [3, 1])
@pytest.mark.parametrize("dims", [(12, 12, 12), (8, 20, 24)])
def test_conv_wgrad(pkg, Ci, Co, ks, dims):
    """tcgen05 weight gradient (voxels as the reduction dim, MN-major operands) vs torch.nn.grad.conv3d_weight."""
    lib = pkg._lib.load()
    g = torch.Generator().manual_seed(Ci * 5 + Co + ks)
    N = 2
    x = torch.randn(N, Ci, *dims, generator=g).to(torch.bfloat16)
    dy = torch.randn(N, Co, *dims, generator=g).to(torch.bfloat16)
    want = torch.nn.grad.conv3d_weight(x.float(), (Co, Ci, ks, ks, ks), dy.float(), padding=ks // 2)
    xp, xo, dp, do = Ci + 8, 8, Co + 16, 16
    x_cl = torch.zeros(N, *dims, xp, dtype=torch.bfloat16); x_cl[..., xo:] = x.permute(0, 2, 3, 4, 1)
    dy_cl = torch.zeros(N, *dims, dp, dtype=torch.bfloat16); dy_cl[..., do:] = dy.permute(0, 2, 3, 4, 1)
    x_cl, dy_cl = x_cl.to(DEV), dy_cl.to(DEV)
    dW = torch.full((Co, Ci, ks, ks, ks), float("nan"), device=DEV)
    pkg._lib.check(lib.b200_test_tc_wgrad(pkg._lib.ptr(x_cl), xp, xo, Ci, pkg._lib.ptr(dy_cl), dp, do, Co, N, *dims, ks,
                                           pkg._lib.ptr(dW), pkg._lib.stream_ptr()), "tc_wgrad")
    torch.cuda.synchronize()
    err = ((dW.cpu() - want).abs().max() / want.abs().max()).item()
    assert err <= 1e-3, err


@pytest.mark.parametrize("op", [1, 2, 4])
@pytest.mark.parametrize("Ci,Co", [(16, 16), (32, 16), (16, 32)])
def test_conv_halo_multi_plane_tiles(pkg, monkeypatch, op, Ci, Co):
    """The halo kernel computes `op` output d-planes per tile from op+2 halo planes (full-resolution layers use op = 4);
    forward with statistics and dgrad+accumulate must not depend on op."""
    monkeypatch.setenv("B200_HALO_MIN_TILES", "1")
    monkeypatch.setenv("B200_HALO_OP", str(op))
    g = torch.Generator().manual_seed(op * 100 + Ci + Co)
    N, dims = 2, (8, 20, 24)
    x = torch.randn(N, Ci, *dims, generator=g).to(torch.bfloat16)
    w = torch.randn(Co, Ci, 3, 3, 3, generator=g) / (Ci * 27) ** 0.5
    want = F.conv3d(x.float(), w.to(torch.bfloat16).float(), padding=1)
    x_cl = x.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    out = torch.zeros(N, *dims, Co, dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(N, Co, 2, dtype=torch.float64, device=DEV)
    run_conv(pkg, x_cl, 0, Ci, w.to(DEV), Co, 3, out, 0, 0, 0, stats)
    got = out.float().permute(0, 4, 1, 2, 3).cpu()
    assert ((got - want).abs().max() / want.abs().max()).item() <= 2 ** -7
    assert torch.allclose(stats.cpu()[..., 1], want.double().square().sum((2, 3, 4)), rtol=1e-3)
    # dgrad (input has Co channels, output Ci) accumulating onto a base tensor
    dy = torch.randn(N, Co, *dims, generator=g).to(torch.bfloat16)
    want_d = F.conv_transpose3d(dy.float(), w.to(torch.bfloat16).float(), padding=1)
    base = torch.randn(N, *dims, Ci, generator=g).to(torch.bfloat16)
    dx = base.clone().to(DEV)
    run_conv(pkg, dy.permute(0, 2, 3, 4, 1).contiguous().to(DEV), 0, Ci, w.to(DEV), Co, 3, dx, 0, 1, 1, None)
    want_d = want_d + base.float().permute(0, 4, 1, 2, 3)
    got_d = dx.float().permute(0, 4, 1, 2, 3).cpu()
    assert ((got_d - want_d).abs().max() / want_d.abs().max()).item() <= 2 ** -7


@pytest.mark.parametrize("op", [1, 4])
@pytest.mark.parametrize("Ci,Co", [(32, 16), (64, 32), (16, 16)])
def test_conv_halo_fused_1x1(pkg, monkeypatch, op, Ci, Co):
    """Residual block fusion: forward conv3(x) and conv1(x) in one launch (two accumulators, two statistic sets); dgrad
    dgrad3(dc1) + dgrad1(dc3) in one launch (second input tile, one accumulator)."""
    monkeypatch.setenv("B200_HALO_MIN_TILES", "1")
    monkeypatch.setenv("B200_HALO_OP", str(op))
    lib = pkg._lib.load(); L = pkg._lib
    g = torch.Generator().manual_seed(op + Ci * 3 + Co)
    N, dims = 2, (8, 20, 24)
    x = torch.randn(N, Ci, *dims, generator=g).to(torch.bfloat16)
    w3 = torch.randn(Co, Ci, 3, 3, 3, generator=g) / (Ci * 27) ** 0.5
    w1 = torch.randn(Co, Ci, 1, 1, 1, generator=g) / Ci ** 0.5
    scratch = torch.empty(2 * Co * Ci * 28, dtype=torch.bfloat16, device=DEV)
    w3_d, w1_d = w3.to(DEV), w1.to(DEV)      # keep the device copies alive across the asynchronous launches
    cl = lambda t: t.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    ncdhw = lambda t: t.float().permute(0, 4, 1, 2, 3).cpu()
    # forward
    out = torch.zeros(N, *dims, Co, dtype=torch.bfloat16, device=DEV); out2 = torch.zeros_like(out)
    st1 = torch.zeros(N, Co, 2, dtype=torch.float64, device=DEV); st2 = torch.zeros_like(st1)
    xc = cl(x)
    L.check(lib.b200_test_tc_conv_fused(L.ptr(xc), None, Ci, Co, N, *dims, L.ptr(w3_d), L.ptr(w1_d), 1, L.ptr(out), L.ptr(out2),
                                        L.ptr(st1), L.ptr(st2), L.ptr(scratch), L.stream_ptr()), "fused fwd")
    torch.cuda.synchronize()
    want3 = F.conv3d(x.float(), w3.to(torch.bfloat16).float(), padding=1)
    want1 = F.conv3d(x.float(), w1.to(torch.bfloat16).float())
    assert ((ncdhw(out) - want3).abs().max() / want3.abs().max()).item() <= 2 ** -7
    assert ((ncdhw(out2) - want1).abs().max() / want1.abs().max()).item() <= 2 ** -7
    assert torch.allclose(st1.cpu()[..., 1], want3.double().square().sum((2, 3, 4)), rtol=1e-3)
    assert torch.allclose(st2.cpu()[..., 0], want1.double().sum((2, 3, 4)), rtol=1e-3, atol=1e-2 * want1.abs().max().item())
    # dgrad
    d1 = torch.randn(N, Co, *dims, generator=g).to(torch.bfloat16); d3 = torch.randn(N, Co, *dims, generator=g).to(torch.bfloat16)
    dx = torch.zeros(N, *dims, Ci, dtype=torch.bfloat16, device=DEV)
    d1c, d3c = cl(d1), cl(d3)
    L.check(lib.b200_test_tc_conv_fused(L.ptr(d1c), L.ptr(d3c), Ci, Co, N, *dims, L.ptr(w3_d), L.ptr(w1_d), 2, L.ptr(dx), None,
                                        None, None, L.ptr(scratch), L.stream_ptr()), "fused dgrad")
    torch.cuda.synchronize()
    want = F.conv_transpose3d(d1.float(), w3.to(torch.bfloat16).float(), padding=1) + F.conv_transpose3d(d3.float(), w1.to(torch.bfloat16).float())
    assert ((ncdhw(dx) - want).abs().max() / want.abs().max()).item() <= 2 ** -7


@pytest.mark.parametrize("Ci,Co,dims", [(16, 16, (12, 16, 12)), (16, 32, (8, 16, 18)), (32, 16, (8, 16, 18)), (16, 16, (8, 20, 24))])
def test_conv_dgrad_with_folded_norm_backward_sums(pkg, Ci, Co, dims):
    """The dgrad of a 3^3 convolution whose input was a = lrelu(norm(c)) accumulates the first pass of that norm's backward in its
    epilogue (tc_conv_halo48.cuh): sum g and sum g*n per (sample, channel), g = dx * lrelu'(a), n recovered from a.  Against torch fp64 on
    the same bf16 operands: the sums are formed from the fp32 accumulators (before the bf16 rounding of dx): 2e-3 relative to the sum of |terms|."""
    import ctypes
    lib = pkg._lib.load()
    g = torch.Generator().manual_seed(Ci + Co + dims[0])
    N = 2
    dy = torch.randn(N, Co, *dims, generator=g).to(torch.bfloat16)
    w = torch.randn(Co, Ci, 3, 3, 3, generator=g) / (Co * 27) ** 0.5
    act = torch.nn.functional.leaky_relu(torch.randn(N, Ci, *dims, generator=g), 0.01).to(torch.bfloat16)
    dx = F.conv_transpose3d(dy.double(), w.to(torch.bfloat16).double(), padding=1)
    a = act.double()
    gg = dx * torch.where(a > 0, 1.0, 0.01)
    nn_ = torch.where(a > 0, a, a * 100.0)
    want = torch.stack([gg.sum((2, 3, 4)), (gg * nn_).sum((2, 3, 4))], -1)
    scale = torch.stack([gg.abs().sum((2, 3, 4)), (gg * nn_).abs().sum((2, 3, 4))], -1)
    cl = lambda t: t.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    dy_cl, act_cl = cl(dy), cl(act)
    out = torch.empty(N, *dims, Ci, dtype=torch.bfloat16, device=DEV)
    acc = torch.full((N, Ci, 3), 7.0, dtype=torch.float64, device=DEV)
    scratch = torch.empty(2 * w.numel(), dtype=torch.bfloat16, device=DEV)
    folded = ctypes.c_int(-1)
    pkg._lib.check(lib.b200_test_tc_conv_dgrad_normbwd(pkg._lib.ptr(dy_cl), Co, Ci, N, *dims, pkg._lib.ptr(w.to(DEV)), pkg._lib.ptr(act_cl), pkg._lib.ptr(out),
                                                        pkg._lib.ptr(acc), ctypes.byref(folded), pkg._lib.ptr(scratch), pkg._lib.stream_ptr()), "dgrad_normbwd")
    torch.cuda.synchronize()
    got_dx = out.float().permute(0, 4, 1, 2, 3).cpu()
    assert ((got_dx - dx.float()).abs().max() / dx.abs().max()).item() <= 2 ** -7
    if Ci == 16:                      # the output-channel count the stacked kernel takes
        assert folded.value == 1
        err = ((acc[..., :2].cpu() - want).abs() / scale).max().item()
        assert err <= 2e-3, err
        assert (acc[..., 2] == 0).all()
    else:
        assert folded.value == 0 and (acc == 0).all()
