"""How long does the host take to ENQUEUE one training step vs how long the GPU takes to run it?"""
import importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dmedicalimagesegmentation_b200")
import bench
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = pkg.MonaiUNETR(**bench.MODEL_KW).to(dev).set_mode("bf16")
loss_fn = pkg.DiceCELoss(to_onehot_y=True, softmax=True)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, fused=True)
x = torch.rand(2, 1, 96, 96, 96, device=dev); y = torch.randint(0, 14, (2, 1, 96, 96, 96), device=dev).float()
def step(parts):
    t = [time.perf_counter()]
    logits = model(x); t.append(time.perf_counter())
    loss = loss_fn(logits, y); t.append(time.perf_counter())
    loss.backward(); t.append(time.perf_counter())
    opt.step(); t.append(time.perf_counter())
    opt.zero_grad(set_to_none=True); t.append(time.perf_counter())
    parts.append([b - a for a, b in zip(t, t[1:])])
for _ in range(5): step([])
torch.cuda.synchronize()
parts = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(20): step(parts)
t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize(); t2 = time.perf_counter()
import numpy as np
p = np.array(parts).mean(0) * 1e3
print(f"host enqueue per step {1e3 * (t1 - t0) / 20:.2f} ms (fwd {p[0]:.2f} loss {p[1]:.2f} bwd {p[2]:.2f} opt {p[3]:.2f} zero {p[4]:.2f}); GPU per step {e0.elapsed_time(e1) / 20:.2f} ms; drain after last enqueue {1e3 * (t2 - t1):.2f} ms")
