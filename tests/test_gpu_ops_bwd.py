"""Each element-wise kernel of the backward on its own, in bf16 mode and fp32 mode, against fp32 torch autograd on the SAME (rounded)
operands with an injected upstream gradient -- so that every kernel upstream of the loss is pinned independently of the
whole-network comparison (whose bf16 gradients are chaotic for a randomly initialised network).
Bar: cosine >= 0.999 and relative L2 <= 2e-2 per output in bf16 mode (outputs are stored as bf16: 2^-9 relative rounding);
1e-4 relative L2 in fp32 mode."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cos_rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item(), ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def check(tag, got, want, bf16):
    c, r = cos_rel(got, want)
    assert c >= (0.999 if bf16 else 0.999999) and r <= (2e-2 if bf16 else 1e-4), (tag, c, r)


@pytest.mark.parametrize("bf16", [1, 0])
@pytest.mark.parametrize("M,H", [(432, 768), (2048, 768), (54, 128)])
def test_layernorm_backward(pkg, bf16, M, H):
    L = pkg._lib; lib = L.load()
    g_ = torch.Generator().manual_seed(M + H + bf16)
    T = torch.bfloat16 if bf16 else torch.float32
    x = (torch.randn(M, H, generator=g_) * 1.5 + 0.3)
    gamma = torch.rand(H, generator=g_) + 0.5
    beta = torch.randn(H, generator=g_)
    up = torch.randn(M, H, generator=g_).to(T)                      # upstream gradient as the kernel receives it
    res = torch.randn(M, H, generator=g_)
    xr = x.clone().requires_grad_(True); gr = gamma.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (H,), gr, br, 1e-5).backward(up.float())
    mean = x.mean(1); rstd = (x.var(1, unbiased=False) + 1e-5).rsqrt()
    stats = torch.stack([mean, rstd], 1).contiguous()
    d = lambda t: t.to(DEV).contiguous()
    dx = torch.empty(M, H, device=DEV); dxc = torch.empty(M, H, device=DEV, dtype=T)
    dg = torch.empty(H, device=DEV); db = torch.empty(H, device=DEV)
    a = [d(up), d(x), d(stats), d(gamma), d(res)]
    L.check(lib.b200_test_layernorm_bwd(L.ptr(a[0]), L.ptr(a[1]), L.ptr(a[2]), L.ptr(a[3]), L.ptr(a[4]), L.ptr(dx), L.ptr(dxc), L.ptr(dg), L.ptr(db),
                                        M, H, bf16, L.stream_ptr()), "ln_bwd")
    torch.cuda.synchronize()
    check("dx", dx.cpu() - res, xr.grad, bf16=False)               # the fp32 residual-stream output is not rounded
    check("dx_cast", dxc.float().cpu(), xr.grad + res, bf16)
    check("dgamma", dg.cpu(), gr.grad, bf16=False)
    check("dbeta", db.cpu(), br.grad, bf16=False)


def _instnorm(x):      # affine-free InstanceNorm3d, eps 1e-5, biased variance over the voxels of each (n, c)
    m = x.mean(1, keepdim=True); v = x.var(1, unbiased=False, keepdim=True)
    return (x - m) * (v + 1e-5).rsqrt(), m.squeeze(1), (v + 1e-5).rsqrt().squeeze(1)


@pytest.mark.parametrize("bf16", [1, 0])
@pytest.mark.parametrize("N,V,C", [(2, 4096, 16), (1, 13824, 32), (2, 1728, 128)])
def test_instnorm_lrelu_backward(pkg, bf16, N, V, C):
    """both forms of res_bwd's norm stage (exec.cuh): out = lrelu(norm(c2) + norm(c3)) and a1 = lrelu(norm(c1)), channels-last"""
    L = pkg._lib; lib = L.load()
    g_ = torch.Generator().manual_seed(N * V + C + bf16)
    T = torch.bfloat16 if bf16 else torch.float32
    TR = torch.float16 if bf16 else torch.float32                   # raw conv outputs are kept as fp16 in bf16 mode
    c2 = (torch.randn(N, V, C, generator=g_) * 3 + 1).to(TR); c3 = (torch.randn(N, V, C, generator=g_) * 0.5 - 2).to(TR)
    up = torch.randn(N, V, C, generator=g_).to(T)
    d = lambda t: t.to(DEV).contiguous()
    acc = torch.empty(N, C, 3, dtype=torch.float64, device=DEV)
    # ---- two inputs
    a2 = c2.float().clone().requires_grad_(True); a3 = c3.float().clone().requires_grad_(True)
    n2, m2, r2 = _instnorm(a2); n3, m3, r3 = _instnorm(a3)
    out = torch.nn.functional.leaky_relu(n2 + n3, 0.01)
    out.backward(up.float())
    act = out.detach().to(T)                                        # the saved activation the kernel reads its sign from
    mr2 = torch.stack([m2.detach(), r2.detach()], -1).contiguous(); mr3 = torch.stack([m3.detach(), r3.detach()], -1).contiguous()
    da = torch.empty(N, V, C, device=DEV, dtype=T); db = torch.empty(N, V, C, device=DEV, dtype=T)
    args = [d(up), d(act), d(c2), d(mr2), d(c3), d(mr3)]
    L.check(lib.b200_test_instnorm_bwd(1, *[L.ptr(t) for t in args], N, C, V, L.ptr(acc), L.ptr(da), L.ptr(db), bf16, L.stream_ptr()), "in_bwd two")
    torch.cuda.synchronize()
    check("dc2", da.float().cpu(), a2.grad, bf16); check("dc3", db.float().cpu(), a3.grad, bf16)
    # the engine's default: no saved activation -- the sign of n2 + n3 is recomputed from c2, c3 and their (mean, rstd)
    da.zero_(); db.zero_()
    args[1] = None
    L.check(lib.b200_test_instnorm_bwd(1, *[L.ptr(t) for t in args], N, C, V, L.ptr(acc), L.ptr(da), L.ptr(db), bf16, L.stream_ptr()), "in_bwd two, sign recomputed")
    torch.cuda.synchronize()
    check("dc2 (sign recomputed)", da.float().cpu(), a2.grad, bf16); check("dc3 (sign recomputed)", db.float().cpu(), a3.grad, bf16)
    # ---- one input, normalised value recovered from the saved activation
    a1 = c2.float().clone().requires_grad_(True)
    n1, m1, r1 = _instnorm(a1)
    o1 = torch.nn.functional.leaky_relu(n1, 0.01)
    o1.backward(up.float())
    act1 = o1.detach().to(T)
    mr1 = torch.stack([m1.detach(), r1.detach()], -1).contiguous()
    args = [d(up), d(act1), None, d(mr1), None, None]
    L.check(lib.b200_test_instnorm_bwd(0, *[L.ptr(t) for t in args], N, C, V, L.ptr(acc), L.ptr(da), None, bf16, L.stream_ptr()), "in_bwd one")
    torch.cuda.synchronize()
    check("dc1", da.float().cpu(), a1.grad, bf16)


@pytest.mark.parametrize("bf16", [1, 0])
@pytest.mark.parametrize("N,V,fs,ncls", [(2, 32768, 16, 14), (1, 5000, 8, 5), (2, 4096, 32, 4)])
def test_head_backward(pkg, bf16, N, V, fs, ncls):
    L = pkg._lib; lib = L.load()
    g_ = torch.Generator().manual_seed(V + fs + ncls + bf16)
    T = torch.bfloat16 if bf16 else torch.float32
    d0 = torch.randn(N, V, fs, generator=g_).to(T)
    Wh = torch.randn(ncls, fs, generator=g_) * 0.3
    dl = torch.randn(N, ncls, V, generator=g_) * 1e-3
    x = d0.float().clone().requires_grad_(True); w = Wh.clone().requires_grad_(True); b = torch.zeros(ncls, requires_grad=True)
    logits = torch.einsum("nvc,kc->nkv", x, w) + b[None, :, None]
    logits.backward(dl)
    d = lambda t: t.to(DEV).contiguous()
    g = torch.empty(N, V, fs, device=DEV, dtype=T); dW = torch.empty(ncls, fs, device=DEV); dB = torch.empty(ncls, device=DEV)
    a = [d(dl), d(d0), d(Wh)]
    L.check(lib.b200_test_head_bwd(L.ptr(a[0]), L.ptr(a[1]), L.ptr(a[2]), ncls, fs, N, V, L.ptr(g), L.ptr(dW), L.ptr(dB), bf16, L.stream_ptr()), "head_bwd")
    torch.cuda.synchronize()
    check("d(d0)", g.float().cpu(), x.grad, bf16)
    check("dW", dW.cpu(), w.grad, bf16=False) if not bf16 else check("dW", dW.cpu(), w.grad, True)
    check("db", dB.cpu(), b.grad, bf16=False)
