import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dmedicalimagesegmentation_b200")
L = pkg._lib; lib = L.load(); dev = "cuda:0"
dbg = torch.zeros(64, dtype=torch.int64, device=dev)
def run(Ci, Co, S, ks=3, N=2, stats=True):
    x = torch.randn(N, S, S, S, Ci, device=dev).bfloat16(); w = torch.randn(Co, Ci, ks, ks, ks, device=dev)
    out = torch.empty(N, S, S, S, Co, device=dev, dtype=torch.bfloat16)
    scratch = torch.empty(2 * w.numel(), dtype=torch.bfloat16, device=dev)
    st = torch.zeros(N, Co, 2, dtype=torch.float64, device=dev) if stats else None
    f = lambda: L.check(lib.b200_test_tc_conv(L.ptr(x), Ci, 0, Ci, N, S, S, S, L.ptr(w), Co, ks, L.ptr(out), Co, 0, 0, 0, L.ptr(st), L.ptr(scratch), L.stream_ptr()), "c")
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    lib.b200_test_set_debug_buffer(L.ptr(dbg)); dbg.zero_(); f(); torch.cuda.synchronize(); lib.b200_test_set_debug_buffer(None)
    d = dbg.cpu().tolist(); t0 = d[0]; r = lambda i: d[i] - t0 if d[i] else -1
    print(f"conv {Ci}->{Co} @{S} k{ks} stats={stats}: {us:.1f} us/launch")
    print("   producer stage-issue  tile0:", [r(1 + i) for i in range(8)], " tile1:", [r(9 + i) for i in range(8)])
    print("   mma stage-ready       tile0:", [r(17 + i) for i in range(8)], " tile1:", [r(25 + i) for i in range(8)])
    print("   epilogue (ready,done) tiles0-3:", [(r(33 + 2 * i), r(34 + 2 * i)) for i in range(4)], " kernel end:", r(48))
run(16, 16, 96); run(16, 16, 96, stats=False); run(32, 16, 96); run(64, 32, 48)
