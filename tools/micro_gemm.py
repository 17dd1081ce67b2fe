import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dmedicalimagesegmentation_b200")
L = pkg._lib; lib = L.load(); dev = "cuda:0"
dbg = torch.zeros(512, dtype=torch.int64, device=dev)
def run(M, N, K, a_mn=0, b_mn=0, iters=20, cold=False):
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    if a_mn: a = a.t().contiguous()
    if b_mn: b = b.t().contiguous()
    out = torch.empty(M, N, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    f = lambda: L.check(lib.b200_test_tc_gemm(L.ptr(a), L.ptr(b), L.ptr(out), M, N, K, a_mn, b_mn, L.stream_ptr()), "g")
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if cold: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); us = ts[len(ts) // 2]
    lib.b200_test_set_debug_buffer(L.ptr(dbg)); dbg.zero_()
    if cold: flush.zero_()
    f(); torch.cuda.synchronize(); lib.b200_test_set_debug_buffer(None)
    d = dbg.cpu().tolist(); t0 = d[0]
    rel = lambda i: (d[i] - t0) if d[i] else -1
    import numpy as np
    g = np.array(d[64:64 + 2 * 148]).reshape(148, 2); g = g[g[:, 0] > 0]
    if len(g):
        t00 = g[:, 0].min(); dur = g[:, 1] - g[:, 0]
        print(f"   globaltimer: {len(g)} CTAs, start spread {g[:,0].max()-t00} ns, kernel span {g[:,1].max()-t00} ns, per-CTA duration min/med/max {dur.min()}/{int(np.median(dur))}/{dur.max()} ns")
    print(f"gemm {M}x{N}x{K} mn={a_mn}{b_mn} {'cold' if cold else 'hot '}: {us:7.1f} us {2*M*N*K/us/1e6:7.1f} TF | cyc: setup {rel(1)} prod_done {rel(2)} kb0..3 {[rel(8+i) for i in range(4)]} mma_done {rel(3)} acc_ready {rel(4)} end {rel(5)} | mma_issued {[rel(12+i) for i in range(4)]} prod_issued {[rel(16+i) for i in range(4)]}")
run(432, 3072, 768); run(432, 3072, 768, cold=True); run(432, 768, 3072); run(432, 768, 3072, cold=True); run(432, 2304, 768); run(432, 768, 768); run(432, 768, 768, cold=True)
run(3072, 768, 432, a_mn=1, b_mn=1); run(768, 3072, 432, a_mn=1, b_mn=1)
