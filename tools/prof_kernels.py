"""Three launches each of the hot tcgen05 kernels at their configs[1] shapes (ncu --set full target; also prints CUDA-event times)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dmedicalimagesegmentation_b200")
L = pkg._lib; lib = L.load(); dev = "cuda:0"
N, S = 2, 96
def timed(f, tag, flops, reps=3):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{tag}: {us:.1f} us/launch  {flops / us * 1e-6:.1f} TFLOP/s", flush=True)
# conv 16->16 @96 (forward with statistics)
x = torch.randn(N, S, S, S, 16, device=dev).bfloat16(); w = torch.randn(16, 16, 3, 3, 3, device=dev)
out = torch.empty(N, S, S, S, 16, device=dev, dtype=torch.bfloat16); scr = torch.empty(2 * w.numel(), dtype=torch.bfloat16, device=dev)
st = torch.zeros(N, 16, 2, dtype=torch.float64, device=dev)
timed(lambda: L.check(lib.b200_test_tc_conv(L.ptr(x), 16, 0, 16, N, S, S, S, L.ptr(w), 16, 3, L.ptr(out), 16, 0, 0, 0, L.ptr(st), L.ptr(scr), L.stream_ptr()), "c"),
      "conv_halo 16->16 @96 (+5 us weight pack)", 2.0 * N * S ** 3 * 16 * 16 * 27)
# wgrad 16x16 @96
dy = torch.randn(N, S, S, S, 16, device=dev).bfloat16(); dW = torch.zeros(16, 16, 3, 3, 3, device=dev)
timed(lambda: L.check(lib.b200_test_tc_wgrad(L.ptr(x), 16, 0, 16, L.ptr(dy), 16, 0, 16, N, S, S, S, 3, L.ptr(dW), L.stream_ptr()), "w"),
      "wgrad_halo 16x16 @96 (+memset)", 2.0 * N * S ** 3 * 16 * 16 * 27)
# ViT GEMMs
for (M, Nn, K) in ((432, 3072, 768), (432, 768, 3072)):
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(Nn, K, device=dev).bfloat16(); o = torch.empty(M, Nn, device=dev)
    timed(lambda: L.check(lib.b200_test_tc_gemm(L.ptr(a), L.ptr(b), L.ptr(o), M, Nn, K, 0, 0, L.stream_ptr()), "g"), f"gemm {M}x{Nn}x{K}", 2.0 * M * Nn * K)
