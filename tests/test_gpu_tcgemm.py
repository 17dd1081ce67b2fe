"""tcgen05/TMEM/TMA GEMM engine against a torch fp32 matmul of the same bf16 operands (floating-point kernel:
torch fp32 is the reference here).  Tolerance 1e-3 relative to max|ref| (fp32 accumulation of exact bf16 products;
only the summation order differs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (432, 768, 768), (216, 216, 64), (216, 64, 216), (1000, 2304, 136),
                                   (432, 3072, 768), (2048, 768, 3072)])
def test_tc_gemm_matches_fp32(pkg, a_mn, b_mn, M, N, K):
    if (a_mn and M % 8) or (b_mn and N % 8) or K % 8:
        pytest.skip("TMA needs 16-byte row pitch")
    lib = pkg._lib.load()
    g = torch.Generator(device="cpu").manual_seed(M * 31 + N * 7 + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    b = torch.randn(N, K, generator=g).to(torch.bfloat16)
    want = a.float() @ b.float().t()
    a_dev = (a.t().contiguous() if a_mn else a).to(DEV)
    b_dev = (b.t().contiguous() if b_mn else b).to(DEV)
    out = torch.full((M, N), float("nan"), device=DEV)
    pkg._lib.check(lib.b200_test_tc_gemm(pkg._lib.ptr(a_dev), pkg._lib.ptr(b_dev), pkg._lib.ptr(out), M, N, K, a_mn, b_mn,
                                          pkg._lib.stream_ptr()), "tc_gemm")
    torch.cuda.synchronize()
    err = ((out.cpu() - want).abs().max() / want.abs().max()).item()
    assert err <= 1e-3, err


@pytest.mark.parametrize("B,heads,L", [(2, 12, 216), (4, 12, 216), (1, 2, 64), (3, 4, 200), (1, 3, 256), (2, 2, 27)])
def test_fused_attention_matches_fp32(pkg, B, heads, L):
    """tc_attention.cuh (SABlock) against torch fp32 softmax(scale q k^T) v on the same bf16 q, k, v.
    Tolerances: probabilities are stored as bf16 (relative rounding 2^-9 -> 4e-3 of max|P|); the output accumulates bf16
    probabilities in fp32 and is stored as bf16 -> 1e-2 of max|O| (the bf16-mode budget of the north_star)."""
    lib = pkg._lib.load()
    H, Lp = heads * 64, (L + 7) & ~7
    g = torch.Generator().manual_seed(B * 1000 + L)
    qkv = torch.randn(B * L, 3 * H, generator=g).to(torch.bfloat16)
    scale = 64 ** -0.5
    f = qkv.float().view(B, L, 3, heads, 64)
    q, k, v = (f[:, :, i].permute(0, 2, 1, 3) for i in range(3))          # [B, heads, L, 64]; columns are [Q|K|V] x head x d (a7)
    p_want = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1)
    o_want = (p_want @ v).permute(0, 2, 1, 3).reshape(B * L, H)
    qkv_d = qkv.to(DEV)
    probs = torch.full((B, heads, L, Lp), float("nan"), dtype=torch.bfloat16, device=DEV)
    att = torch.full((B * L, H), float("nan"), dtype=torch.bfloat16, device=DEV)
    if L < 16:
        pytest.skip("fused kernel takes 16 <= L <= 256")
    pkg._lib.check(lib.b200_test_tc_attention(pkg._lib.ptr(qkv_d), pkg._lib.ptr(probs), pkg._lib.ptr(att), B, heads, L, Lp, H, scale,
                                               pkg._lib.stream_ptr()), "tc_attention")
    torch.cuda.synchronize()
    p_got = probs.float().cpu()
    assert torch.isfinite(p_got).all() and torch.isfinite(att.float()).all()
    assert (p_got[..., :L] - p_want).abs().max().item() <= 4e-3 * p_want.max().item()
    assert (p_got[..., L:] == 0).all()
    err = ((att.float().cpu() - o_want).abs().max() / o_want.abs().max()).item()
    assert err <= 1e-2, err
    # probabilities are optional (inference)
    att2 = torch.zeros_like(att)
    pkg._lib.check(lib.b200_test_tc_attention(pkg._lib.ptr(qkv_d), None, pkg._lib.ptr(att2), B, heads, L, Lp, H, scale,
                                               pkg._lib.stream_ptr()), "tc_attention")
    assert torch.equal(att2, att)


@pytest.mark.parametrize("B,heads,L", [(2, 12, 216), (1, 2, 64), (3, 4, 200), (1, 3, 256), (2, 2, 27)])
def test_fused_attention_backward_dq_matches_fp32(pkg, B, heads, L):
    """Query-row half of the attention backward (tc_attention.cuh, BWD = 1) against torch fp32 on the same bf16 inputs:
    dS = P * (dP - rowsum(dP * P)) * scale with dP = dO V^T, dQ = dS K.  dS and dQ are stored as bf16 and dQ accumulates the
    bf16 dS: 1e-2 of max|ref| (the bf16-mode budget); the K and V thirds of dqkv must stay untouched."""
    lib = pkg._lib.load()
    H, Lp = heads * 64, (L + 7) & ~7
    g = torch.Generator().manual_seed(B * 77 + L)
    qkv = torch.randn(B * L, 3 * H, generator=g).to(torch.bfloat16)
    datt = torch.randn(B * L, H, generator=g).to(torch.bfloat16)
    scale = 64 ** -0.5
    f = qkv.float().view(B, L, 3, heads, 64)
    q, k, v = (f[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    probs = torch.zeros(B, heads, L, Lp)
    probs[..., :L] = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1)
    probs = probs.to(torch.bfloat16)
    pf = probs.float()[..., :L]
    do = datt.float().view(B, L, heads, 64).permute(0, 2, 1, 3)
    dp = do @ v.transpose(-1, -2)
    ds_want = pf * (dp - (dp * pf).sum(-1, keepdim=True)) * scale
    dq_want = (ds_want @ k).permute(0, 2, 1, 3).reshape(B * L, H)
    dS = torch.full((B, heads, L, Lp), float("nan"), dtype=torch.bfloat16, device=DEV)
    dqkv = torch.full((B * L, 3 * H), 7.0, dtype=torch.bfloat16, device=DEV)
    qkv_d, probs_d, datt_d = qkv.to(DEV), probs.to(DEV), datt.to(DEV)      # named: temporaries would be freed (and aliased) before the launch
    pkg._lib.check(lib.b200_test_tc_attention_bwd(pkg._lib.ptr(qkv_d), pkg._lib.ptr(probs_d), pkg._lib.ptr(datt_d),
                                                   pkg._lib.ptr(dS), pkg._lib.ptr(dqkv), B, heads, L, Lp, H, scale, pkg._lib.stream_ptr()),
                   "tc_attention_bwd")
    torch.cuda.synchronize()
    ds_got = dS.float().cpu()
    assert torch.isfinite(ds_got).all()
    assert ((ds_got[..., :L] - ds_want).abs().max() / ds_want.abs().max()).item() <= 1e-2
    assert (ds_got[..., L:] == 0).all()
    got = dqkv.float().cpu()
    assert ((got[:, :H] - dq_want).abs().max() / dq_want.abs().max()).item() <= 1e-2
    assert (got[:, H:] == 7.0).all()


@pytest.mark.parametrize("B,heads,L", [(2, 12, 216), (1, 2, 64), (3, 4, 200), (1, 3, 256), (2, 2, 27)])
def test_fused_attention_backward_kv_matches_fp32(pkg, B, heads, L):
    """Key-row half of the attention backward (attn_bwd_kv_kernel): dV = P^T dO and dK = dS^T Q against torch fp32 on the same
    bf16 operands; 1e-2 of max|ref| (bf16 outputs); the Q third of dqkv must stay untouched."""
    lib = pkg._lib.load()
    H, Lp = heads * 64, (L + 7) & ~7
    g = torch.Generator().manual_seed(B * 131 + L)
    qkv = torch.randn(B * L, 3 * H, generator=g).to(torch.bfloat16)
    datt = torch.randn(B * L, H, generator=g).to(torch.bfloat16)
    probs = torch.zeros(B, heads, L, Lp)
    probs[..., :L] = torch.softmax(torch.randn(B, heads, L, L, generator=g), dim=-1)
    dS = torch.zeros(B, heads, L, Lp)
    dS[..., :L] = torch.randn(B, heads, L, L, generator=g) * 0.05
    probs, dS = probs.to(torch.bfloat16), dS.to(torch.bfloat16)
    q = qkv.float().view(B, L, 3, heads, 64)[:, :, 0].permute(0, 2, 1, 3)
    do = datt.float().view(B, L, heads, 64).permute(0, 2, 1, 3)
    dv_want = (probs.float()[..., :L].transpose(-1, -2) @ do).permute(0, 2, 1, 3).reshape(B * L, H)
    dk_want = (dS.float()[..., :L].transpose(-1, -2) @ q).permute(0, 2, 1, 3).reshape(B * L, H)
    qkv_d, probs_d, ds_d, datt_d = qkv.to(DEV), probs.to(DEV), dS.to(DEV), datt.to(DEV)
    dqkv = torch.full((B * L, 3 * H), 7.0, dtype=torch.bfloat16, device=DEV)
    pkg._lib.check(lib.b200_test_tc_attention_bwd_kv(pkg._lib.ptr(qkv_d), pkg._lib.ptr(probs_d), pkg._lib.ptr(ds_d), pkg._lib.ptr(datt_d),
                                                      pkg._lib.ptr(dqkv), B, heads, L, Lp, H, pkg._lib.stream_ptr()), "tc_attention_bwd_kv")
    torch.cuda.synchronize()
    got = dqkv.float().cpu()
    assert (got[:, :H] == 7.0).all()
    assert ((got[:, H:2 * H] - dk_want).abs().max() / dk_want.abs().max()).item() <= 1e-2
    assert ((got[:, 2 * H:] - dv_want).abs().max() / dv_want.abs().max()).item() <= 1e-2


@pytest.mark.parametrize("mn", [1, 0])
@pytest.mark.parametrize("tokens,layers", [(432, 4), (2048, 1), (216, 2), (104, 1)])
def test_grouped_gemm_matches_fp32(pkg, mn, tokens, layers):
    """tc_gemm_grouped.cuh: the deferred ViT weight gradients dW = dY^T X of `layers` transformer blocks (4 problems each: fc2, fc1,
    out_proj, qkv at ViT-B sizes, reduction over `tokens` rows with a ragged last k-block) in ONE launch, against torch fp32 on the
    same bf16 operands.  Tolerance 1e-3 of max|ref| (fp32 accumulation of exact bf16 products; only the order differs)."""
    import ctypes
    lib = pkg._lib.load()
    H, F = 768, 3072
    shapes = [(H, F), (F, H), (H, H), (3 * H, H)] * layers          # (out features = GEMM M, in features = GEMM N)
    if tokens == 104:
        shapes = [(136, 200), (64, 72), (520, 264)]                   # ragged M / N edges
    g = torch.Generator().manual_seed(tokens + layers)
    A, Bm, O, want = [], [], [], []
    for (m, n) in shapes:
        dy = torch.randn(tokens, m, generator=g).to(torch.bfloat16)     # [K, M]
        x = torch.randn(tokens, n, generator=g).to(torch.bfloat16)      # [K, N]
        want.append(dy.float().t() @ x.float())
        A.append((dy if mn else dy.t().contiguous()).to(DEV))
        Bm.append((x if mn else x.t().contiguous()).to(DEV))
        O.append(torch.full((m, n), float("nan"), device=DEV))
    n = len(shapes)
    arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
    ints = lambda vs: (ctypes.c_int * n)(*vs)
    pkg._lib.check(lib.b200_test_tc_gemm_grouped(arr(A), arr(Bm), arr(O), ints([s[0] for s in shapes]), ints([s[1] for s in shapes]),
                                                  ints([tokens] * n), n, mn, pkg._lib.stream_ptr()), "tc_gemm_grouped")
    torch.cuda.synchronize()
    for o, w in zip(O, want):
        err = ((o.cpu() - w).abs().max() / w.abs().max()).item()
        assert err <= 1e-3, err
