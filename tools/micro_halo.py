"""Halo conv kernel tuning: time per launch under the B200_HALO_DBG skip modes + CTA-0 clock stamps."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dmedicalimagesegmentation_b200")
L = pkg._lib; lib = L.load(); dev = "cuda:0"
dbg = torch.zeros(128, dtype=torch.int64, device=dev)
def run(Ci, Co, S, N=2, stats=True, modes=(0, 1, 2, 4, 3, 6, 7), stamps=True):
    x = torch.randn(N, S, S, S, Ci, device=dev).bfloat16(); w = torch.randn(Co, Ci, 3, 3, 3, device=dev)
    out = torch.empty(N, S, S, S, Co, device=dev, dtype=torch.bfloat16)
    scratch = torch.empty(2 * w.numel(), dtype=torch.bfloat16, device=dev)
    st = torch.zeros(N, Co, 2, dtype=torch.float64, device=dev) if stats else None
    f = lambda: L.check(lib.b200_test_tc_conv(L.ptr(x), Ci, 0, Ci, N, S, S, S, L.ptr(w), Co, 3, L.ptr(out), Co, 0, 0, 0, L.ptr(st), L.ptr(scratch), L.stream_ptr()), "c")
    for mode in modes:
        os.environ["B200_HALO_DBG"] = str(mode)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        flops = 2.0 * N * S ** 3 * Ci * Co * 27
        print(f"halo conv {Ci}->{Co} @{S} stats={stats} mode={mode} (1=noTMA 2=noMMA 4=noEpi): {us:.1f} us/launch (incl ~5us pack)  {flops / us * 1e-6:.1f} TFLOP/s", flush=True)
        if stamps and mode == 0:
            lib.b200_test_set_debug_buffer(L.ptr(dbg)); dbg.zero_(); f(); torch.cuda.synchronize(); lib.b200_test_set_debug_buffer(None)
            d = dbg.cpu().tolist(); t0 = d[0]; r = lambda i: d[i] - t0 if d[i] else -1
            print("   producer issue (tile,kd):", [r(1 + i) for i in range(12)])
            print("   mma stage ready (tile,kd):", [r(16 + i) for i in range(12)])
            print("   mma issued (tile,kd):", [r(64 + i) for i in range(12)])
            print("   committed (tile,kd):", [r(80 + i) for i in range(12)])
            print("   epilogue (ready,done) tiles0-3:", [(r(32 + 2 * i), r(33 + 2 * i)) for i in range(4)], " kernel end:", r(48))
    os.environ["B200_HALO_DBG"] = "0"
if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "prof":
        run(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), modes=(0,), stamps=False)
    else:
        os.environ["B200_HALO_OP"] = "4"
        run(16, 16, 96, modes=(0, 64, 80), stamps=False)
        run(32, 16, 96, modes=(0, 64), stamps=False)
        run(64, 32, 48, modes=(0, 64), stamps=False)
