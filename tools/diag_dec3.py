"""Bisect the decoder3 block backward: compare GPU scratch buffers with a CPU emulation (fp64 autograd hooks)."""
import copy, importlib, os, sys, ctypes
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unetr_oracle as O
pkg = importlib.import_module("3dmedicalimagesegmentation_b200")
U = importlib.import_module("3dmedicalimagesegmentation_b200.unetr")
lib = pkg._lib.load()
cfg = dict(in_channels=1, out_channels=5, img_size=(32, 32, 32), feature_size=8, hidden_size=64, mlp_dim=128, num_heads=4, pos_embed="conv", norm_name="instance", res_block=True)
torch.manual_seed(0)
ref = O.UNETR(**cfg); ref64 = copy.deepcopy(ref).double()
x = torch.rand(1, 1, 32, 32, 32, generator=torch.Generator().manual_seed(5))
cap = {}
blk = ref64.decoder3.conv_block
blk.register_forward_hook(lambda m, i, o: cap.update(x=i[0].detach(), out=o.detach()))
blk.register_full_backward_hook(lambda m, gi, go: cap.update(dout=go[0].detach(), dx=gi[0].detach()))
blk.conv1.register_full_backward_hook(lambda m, gi, go: cap.update(dc1=go[0].detach()))
blk.conv2.register_full_backward_hook(lambda m, gi, go: cap.update(dc2=go[0].detach(), da1=gi[0].detach()))
blk.conv2.register_forward_hook(lambda m, i, o: cap.update(a1=i[0].detach()))
e, l = ref64(x.double()); l.square().mean().backward()

mine = pkg.UNETR(**cfg); mine.load_state_dict(ref.state_dict()); mine = mine.cuda().set_mode("fp32")
# patch the flags to stop after decoder3
orig = lib.b200_unetr_backward
class Wrap:
    def __call__(self, h, p, g, x, ws, de, dl, flags, st):
        Wrap.h = h
        return orig(h, p, g, x, ws, de, dl, flags | 16, st)
lib.b200_unetr_backward = Wrap()
e2, l2 = mine(x.cuda()); l2.square().mean().backward()
torch.cuda.synchronize()
def peek(name, shape):
    t = torch.empty(shape, dtype=torch.float32, device="cuda")
    pkg._lib.check(orig.__self__.b200_unetr_peek(Wrap.h, name.encode(), pkg._lib.ptr(t), t.numel() * 4) if False else lib.b200_unetr_peek(Wrap.h, name.encode(), pkg._lib.ptr(t), t.numel() * 4), "peek")
    return t.cpu()
def cl2nc(t):  # [N,D,H,W,C] -> [N,C,D,H,W]
    return t.permute(0, 4, 1, 2, 3)
def err(a, b): return ((a.double() - b).abs().max() / b.abs().max()).item()
S, C = 16, 16
for name, key in (("dc2", "dc2"), ("da1", "da1"), ("dc1", "dc1"), ("a1_dec3", "a1")):
    t = cl2nc(peek(name, (1, S, S, S, C)))
    print(f"{name:8s} err {err(t, cap[key]):.3e}   max|ref| {cap[key].abs().max().item():.3e}")
t = cl2nc(peek("dcat", (1, S, S, S, 2 * C)))
print(f"dcat     err {err(t, cap['dx']):.3e}")
t = cl2nc(peek("gA", (1, S, S, S, C)))
print(f"gA(dout) err {err(t, cap['dout']):.3e}")
mr = peek("mr1_dec3", (C, 2))
c1 = F.conv3d(cap['x'], blk.conv1.conv.weight.detach(), padding=1)
print("mean err", (mr[:, 0].double() - c1.mean((0, 2, 3, 4))).abs().max().item(), " rstd err", (mr[:, 1].double() - 1 / torch.sqrt(c1.var((0, 2, 3, 4), unbiased=False) + 1e-5)).abs().max().item())
# where is dc1 wrong?
d = (cl2nc(peek("dc1", (1, S, S, S, C))).double() - cap['dc1']).abs()
print("dc1 abs err: max", d.max().item(), "at", torch.nonzero(d == d.max())[0].tolist(), " #(>1e-4*max)", (d > 1e-4 * cap['dc1'].abs().max()).sum().item(), "of", d.numel())
a1g = cl2nc(peek("a1_dec3", (1, S, S, S, C)))
print("sign mismatches a1:", ((a1g > 0) != (cap['a1'] > 0)).sum().item())
