"""The hot tcgen05 kernels at their configs[1] shapes, one after the other (ncu --set full target; also prints CUDA-event times).
`--ncu`: two launches each (one warm, one to capture) so that a capture of every launch stays short."""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dmedicalimagesegmentation_b200")
L = pkg._lib; lib = L.load(); dev = "cuda:0"
N, S = 2, 96
ONCE = "--ncu" in sys.argv
def timed(f, tag, flops, reps=3):
    for _ in range(1 if ONCE else 3): f()
    torch.cuda.synchronize()
    reps = 1 if ONCE else reps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{tag}: {us:.1f} us/launch  {flops / us * 1e-6:.1f} TFLOP/s", flush=True)
# conv 16->16 @96 (forward with statistics): tc_conv_halo48.cuh
x = torch.randn(N, S, S, S, 16, device=dev).bfloat16(); w = torch.randn(16, 16, 3, 3, 3, device=dev)
out = torch.empty(N, S, S, S, 16, device=dev, dtype=torch.bfloat16); scr = torch.empty(2 * w.numel(), dtype=torch.bfloat16, device=dev)
st = torch.zeros(N, 16, 2, dtype=torch.float64, device=dev)
timed(lambda: L.check(lib.b200_test_tc_conv(L.ptr(x), 16, 0, 16, N, S, S, S, L.ptr(w), 16, 3, L.ptr(out), 16, 0, 0, 0, L.ptr(st), L.ptr(scr), L.stream_ptr()), "c"),
      "conv_halo48 16->16 @96 (+5 us weight pack)", 2.0 * N * S ** 3 * 16 * 16 * 27)
# wgrad 16x16 @96
dy = torch.randn(N, S, S, S, 16, device=dev).bfloat16(); dW = torch.zeros(16, 16, 3, 3, 3, device=dev)
timed(lambda: L.check(lib.b200_test_tc_wgrad(L.ptr(x), 16, 0, 16, L.ptr(dy), 16, 0, 16, N, S, S, S, 3, L.ptr(dW), L.stream_ptr()), "w"),
      "wgrad_halo 16x16 @96 (+memset)", 2.0 * N * S ** 3 * 16 * 16 * 27)
# ViT GEMMs
for (M, Nn, K) in ((432, 3072, 768), (432, 768, 3072)):
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(Nn, K, device=dev).bfloat16(); o = torch.empty(M, Nn, device=dev)
    timed(lambda: L.check(lib.b200_test_tc_gemm(L.ptr(a), L.ptr(b), L.ptr(o), M, Nn, K, 0, 0, L.stream_ptr()), "g"), f"gemm {M}x{Nn}x{K}", 2.0 * M * Nn * K)
# grouped weight gradients of 4 transformer blocks (tc_gemm_grouped.cuh)
H, F, tok = 768, 3072, 432
shapes = [(H, F), (F, H), (H, H), (3 * H, H)] * 4
A = [torch.randn(tok, m, device=dev).bfloat16() for m, n in shapes]; Bm = [torch.randn(tok, n, device=dev).bfloat16() for m, n in shapes]
O = [torch.empty(m, n, device=dev) for m, n in shapes]; n = len(shapes)
arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts]); ints = lambda vs: (ctypes.c_int * n)(*vs)
ga, gb, go, Ms, Ns, Ks = arr(A), arr(Bm), arr(O), ints([s[0] for s in shapes]), ints([s[1] for s in shapes]), ints([tok] * n)
timed(lambda: L.check(lib.b200_test_tc_gemm_grouped(ga, gb, go, Ms, Ns, Ks, n, 1, L.stream_ptr()), "gg"), "gemm_grouped x16 (4 blocks of ViT-B weight gradients)",
      sum(2.0 * m * nn * tok for m, nn in shapes))
# fused attention forward, 2 x 12 heads x 216 tokens
qkv = torch.randn(2 * 216, 3 * H, device=dev).bfloat16(); probs = torch.empty(2, 12, 216, 216, device=dev, dtype=torch.bfloat16)
att = torch.empty(2 * 216, H, device=dev, dtype=torch.bfloat16)
timed(lambda: L.check(lib.b200_test_tc_attention(L.ptr(qkv), L.ptr(probs), L.ptr(att), 2, 12, 216, 216, H, 0.125, L.stream_ptr()), "att"), "attn_fwd L216 b24",
      2 * 2.0 * 2 * 12 * 216 * 216 * 64)
# DiceCE staged kernels
lg = torch.randn(N, 14, S, S, S, device=dev); lab = torch.randint(0, 14, (N, 1, S, S, S), device=dev).float()
sc = torch.empty(lib.b200_dicece_scratch_bytes(N, 14), dtype=torch.uint8, device=dev); o3 = torch.empty(3, device=dev); dl = torch.empty_like(lg); up = torch.ones(1, device=dev)
timed(lambda: L.check(lib.b200_dicece_forward(L.ptr(lg), L.ptr(lab), N, 14, S ** 3, L.ptr(sc), L.ptr(o3), L.stream_ptr()), "df"), "dicece_fwd (bytes as flops)", N * S ** 3 * 60.0)
timed(lambda: L.check(lib.b200_dicece_backward(L.ptr(lg), L.ptr(lab), N, 14, S ** 3, L.ptr(sc), L.ptr(up), L.ptr(dl), L.stream_ptr()), "db"), "dicece_bwd (bytes as flops)", N * S ** 3 * 116.0)
