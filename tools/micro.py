"""Micro-benchmark of the tcgen05 kernels on the shapes that dominate the step (dev tool; also the ncu target)."""
import importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dmedicalimagesegmentation_b200")
L = pkg._lib; lib = L.load()
dev = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20

def timeit(fn, n=iters):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

if which in ("all", "gemm"):
    for (M, N, K) in [(432, 3072, 768), (432, 768, 3072), (432, 2304, 768), (432, 768, 768), (4096, 4096, 4096)]:
        a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
        out = torch.empty(M, N, device=dev)
        us = timeit(lambda: L.check(lib.b200_test_tc_gemm(L.ptr(a), L.ptr(b), L.ptr(out), M, N, K, 0, 0, L.stream_ptr()), "g"))
        print(f"gemm {M}x{N}x{K}: {us:8.1f} us  {2*M*N*K/us/1e6:8.1f} TFLOP/s")
if which in ("all", "conv"):
    for (Ci, Co, S, ks) in [(16, 16, 96, 3), (32, 16, 96, 3), (64, 32, 48, 3), (128, 64, 24, 3)]:
        N = 2
        x = torch.randn(N, S, S, S, Ci, device=dev).bfloat16(); w = torch.randn(Co, Ci, ks, ks, ks, device=dev)
        out = torch.empty(N, S, S, S, Co, device=dev, dtype=torch.bfloat16)
        scratch = torch.empty(2 * w.numel(), dtype=torch.bfloat16, device=dev)
        stats = torch.zeros(N, Co, 2, dtype=torch.float64, device=dev)
        f = lambda: L.check(lib.b200_test_tc_conv(L.ptr(x), Ci, 0, Ci, N, S, S, S, L.ptr(w), Co, ks, L.ptr(out), Co, 0, 0, 0, L.ptr(stats), L.ptr(scratch), L.stream_ptr()), "c")
        us = timeit(f)
        fl = 2 * N * S**3 * Ci * Co * ks**3
        print(f"conv {Ci}->{Co} @{S} k{ks}: {us:8.1f} us  {fl/us/1e6:8.1f} TFLOP/s   (incl. ~10us weight pack)")
if which in ("all", "wgrad"):
    for (Ci, Co, S, ks) in [(16, 16, 96, 3), (32, 16, 96, 3), (64, 32, 48, 3)]:
        N = 2
        x = torch.randn(N, S, S, S, Ci, device=dev).bfloat16(); dy = torch.randn(N, S, S, S, Co, device=dev).bfloat16()
        dW = torch.empty(Co, Ci, ks, ks, ks, device=dev)
        f = lambda: L.check(lib.b200_test_tc_wgrad(L.ptr(x), Ci, 0, Ci, L.ptr(dy), Co, 0, Co, N, S, S, S, ks, L.ptr(dW), L.stream_ptr()), "w")
        us = timeit(f)
        fl = 2 * N * S**3 * Ci * Co * ks**3
        print(f"wgrad {Ci}x{Co} @{S} k{ks}: {us:8.1f} us  {fl/us/1e6:8.1f} TFLOP/s")
