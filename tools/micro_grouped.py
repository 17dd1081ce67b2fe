"""Grouped tcgen05 GEMM (deferred ViT weight gradients of 4 transformer blocks) timed alone with CUDA events:
operand majors, tile width and a no-store variant (B200_GROUP_DBG=1) separate main loop from epilogue."""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dmedicalimagesegmentation_b200")
L = pkg._lib; lib = L.load(); dev = "cuda:0"
H, F = 768, 3072
def run(tokens, layers, mn, env, iters=10):
    shapes = [(H, F), (F, H), (H, H), (3 * H, H)] * layers
    A = [torch.randn((tokens, m) if mn else (m, tokens), device=dev).bfloat16() for m, n in shapes]
    B = [torch.randn((tokens, n) if mn else (n, tokens), device=dev).bfloat16() for m, n in shapes]
    O = [torch.empty(m, n, device=dev) for m, n in shapes]
    n = len(shapes)
    arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts]); ints = lambda vs: (ctypes.c_int * n)(*vs)
    a, b, o, Ms, Ns, Ks = arr(A), arr(B), arr(O), ints([s[0] for s in shapes]), ints([s[1] for s in shapes]), ints([tokens] * n)
    for k in ("B200_GROUP_BN", "B200_GROUP_DBG"): os.environ.pop(k, None)
    os.environ.update(env)
    f = lambda: L.check(lib.b200_test_tc_gemm_grouped(a, b, o, Ms, Ns, Ks, n, mn, L.stream_ptr()), "gg")
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); us = ts[len(ts) // 2]
    fl = sum(2.0 * m * nn * tokens for m, nn in shapes); by = sum(4.0 * m * nn for m, nn in shapes)
    print(f"grouped x{n} tokens {tokens} mn={mn} {env}: {us:7.1f} us  {fl / us * 1e-6:7.1f} TFLOP/s  output {by / us * 1e-3:7.1f} GB/s", flush=True)
if __name__ == "__main__" and len(sys.argv) > 1:      # ncu target: the production configuration only
    run(432, 4, 1, {}, iters=1)
elif __name__ == "__main__":
    for mn in (1, 0):
        for env in ({}, {"B200_GROUP_DBG": "1"}, {"B200_GROUP_BN": "128"}, {"B200_GROUP_BN": "128", "B200_GROUP_DBG": "1"}):
            run(432, 4, mn, env)
    run(2048, 4, 1, {}); run(2048, 4, 1, {"B200_GROUP_DBG": "1"})
