"""Phase breakdown of one sliding-window predictor call (4 windows of 96^3, no_grad) and of the gather/accumulate kernels."""
import importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dmedicalimagesegmentation_b200")
import bench
L = pkg._lib; lib = L.load(); dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = pkg.MonaiUNETR(**bench.MODEL_KW).to(dev).set_mode("bf16").eval()
x = torch.rand(4, 1, 96, 96, 96, device=dev)
with torch.no_grad():
    for _ in range(5): model(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): model(x)
    e1.record(); torch.cuda.synchronize()
    print(f"predictor call (4 windows): {e0.elapsed_time(e1) / 20:.3f} ms")
    for lvl in (2, 1):
        lib.b200_prof_enable(lvl); model(x); rep = L.prof_report(); lib.b200_prof_enable(0)
        tot = sum(v[0] for v in rep.values())
        print(f"--- prof level {lvl}: total {tot:.3f} ms")
        for k, v in sorted(rep.items(), key=lambda kv: -kv[1][0])[:22]:
            print(f"  {k:36s} {v[0]:8.3f} ms {v[1]:4d} calls")
    vol = torch.rand(1, 1, 512, 512, 256, device=dev)
    pkg.sliding_window_inference(vol, (96,) * 3, 4, model, overlap=0.5); torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record()
    pkg.sliding_window_inference(vol, (96,) * 3, 4, model, overlap=0.5)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"sliding window: GPU {e0.elapsed_time(e1):.1f} ms, host enqueue {1e3 * (t1 - t0):.1f} ms")
    # the loop's own kernels without the predictor: a predictor that returns a fixed tensor leaves gather + accumulate + finalize
    fixed = torch.randn(4, 14, 96, 96, 96, device=dev)
    fake = lambda w: fixed[:w.shape[0]]
    for _ in range(2):
        pkg.sliding_window_inference(vol, (96,) * 3, 4, fake, overlap=0.5)
    torch.cuda.synchronize()
    e0.record()
    pkg.sliding_window_inference(vol, (96,) * 3, 4, fake, overlap=0.5)
    e1.record(); torch.cuda.synchronize()
    print(f"sliding window with a constant predictor (gather + accumulate + normalise only): {e0.elapsed_time(e1):.1f} ms")
    e0.record()
    pkg.sliding_window_inference(vol, (96,) * 3, 4, fake, overlap=0.5, return_argmax=True)
    e1.record(); torch.cuda.synchronize()
    print(f"  ... with the argmax mask: {e0.elapsed_time(e1):.1f} ms")
