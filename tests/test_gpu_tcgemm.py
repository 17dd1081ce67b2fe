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
