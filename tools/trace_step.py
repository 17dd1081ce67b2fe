"""In-situ timeline of the tcgen05 launches of one training step (globaltimer stamps written by the kernels themselves)."""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dmedicalimagesegmentation_b200")
import bench
L = pkg._lib; lib = L.load(); dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = pkg.MonaiUNETR(**bench.MODEL_KW).to(dev).set_mode("bf16")
loss_fn = pkg.DiceCELoss(to_onehot_y=True, softmax=True)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
x = torch.rand(2, 1, 96, 96, 96, device=dev); y = torch.randint(0, 14, (2, 1, 96, 96, 96), device=dev).float()
def step():
    loss = loss_fn(model(x), y); loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(5): step()
torch.cuda.synchronize()
cap = 1024
buf = torch.empty(cap, 2, dtype=torch.int64, device=dev); buf[:, 0] = torch.iinfo(torch.int64).max; buf[:, 1] = 0
lib.b200_trace_begin(L.ptr(buf), cap)
step(); torch.cuda.synchronize()
n = lib.b200_trace_count()
tb = ctypes.create_string_buffer(1 << 17); lib.b200_trace_tags(tb, len(tb)); tags = tb.value.decode().splitlines()
lib.b200_trace_begin(None, 0)
t = buf[:n].cpu().tolist()
t0 = t[0][0]
import collections
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
prev_end = None
for (a, b), tag in zip(t, tags):
    gap = (a - prev_end) / 1e3 if prev_end else 0.0
    agg[tag][0] += 1; agg[tag][1] += (b - a) / 1e3; agg[tag][2] += gap
    prev_end = b
print(f"{n} traced launches; span {(t[-1][1] - t0) / 1e3:.1f} us; sum of durations {sum(b - a for a, b in t) / 1e3:.1f} us")
print(f"{'kind':44s} {'n':>4s} {'avg us':>8s} {'total us':>9s} {'avg gap before (us, incl. untraced kernels)':>12s}")
for tag, (c, d, g) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{tag:44s} {c:4d} {d / c:8.1f} {d:9.1f} {g / c:8.1f}")
if len(sys.argv) > 1:
    for (a, b), tag in list(zip(t, tags))[:int(sys.argv[1])]:
        print(f"{(a - t0) / 1e3:9.1f} {(b - a) / 1e3:7.1f}  {tag}")
