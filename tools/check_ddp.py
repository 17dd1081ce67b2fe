"""2+ ranks: the overlapped (event-gated, 4-group) gradient all-reduce must equal the plain flat all-reduce bit for bit."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dmedicalimagesegmentation_b200")
par = importlib.import_module("3dmedicalimagesegmentation_b200.parallel")
import bench
rank, world, local = par.init_from_env()
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
torch.manual_seed(0)
model = pkg.MonaiUNETR(**bench.MODEL_KW).to(dev).set_mode("bf16")
loss_fn = pkg.DiceCELoss(to_onehot_y=True, softmax=True)
g = torch.Generator().manual_seed(100 + rank)
x = torch.rand(2, 1, 96, 96, 96, generator=g).to(dev); y = torch.randint(0, 14, (2, 1, 96, 96, 96), generator=g).float().to(dev)
def grads(overlap):
    model.overlap_grad_reduce = False
    red = par.GradientAllReduce(model, world, overlap=overlap)
    model.zero_grad(set_to_none=True)
    loss_fn(model(x), y).backward()
    red.reduce()
    torch.cuda.synchronize()
    return torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None]).clone()
a = grads(False); a2 = grads(False); b = grads(True); c = grads(True)
rel = lambda u, v: ((u - v).norm() / u.norm()).item()
# the backward itself is not bitwise reproducible (fp32 atomics in the weight-gradient kernels): flat-vs-flat is the noise floor;
# a bucket reduced too early would differ by O(1)
print(f"rank {rank}: flat vs flat {rel(a, a2):.2e}   overlapped vs flat {rel(a, b):.2e} {rel(a, c):.2e}   max|d| {(a - b).abs().max().item():.2e} (noise {(a - a2).abs().max().item():.2e})  |g| {a.norm().item():.4e}", flush=True)
assert rel(a, b) <= 10 * max(rel(a, a2), 1e-7) and rel(a, c) <= 10 * max(rel(a, a2), 1e-7)
par.shutdown(world)
