"""Per-parameter gradient error table (mine vs fp64 oracle, oracle32 vs fp64) for small configs. Dev tool."""
import copy, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unetr_oracle as O
pkg = importlib.import_module("3dmedicalimagesegmentation_b200")

def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()

def run(mode, kw, loss_kind):
    cfg = dict(in_channels=1, out_channels=5, img_size=(32, 32, 32), feature_size=8, hidden_size=64, mlp_dim=128,
               num_heads=4, pos_embed="perceptron", norm_name="instance", res_block=True); cfg.update(kw)
    torch.manual_seed(0)
    ref = O.UNETR(**cfg); ref64 = copy.deepcopy(ref).double()
    mine = pkg.UNETR(**cfg); mine.load_state_dict(ref.state_dict()); mine = mine.cuda().set_mode(mode)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(1, cfg["in_channels"], *cfg["img_size"], generator=g)
    def loss(e, l):
        return {"both": l.square().mean() + e.square().mean(), "logits": l.square().mean(), "enc4": e.square().mean()}[loss_kind]
    loss(*ref(x)).backward(); loss(*ref64(x.double())).backward(); loss(*mine(x.cuda())).backward()
    print(f"--- mode={mode} kw={kw} loss={loss_kind}")
    for (k, p), (_, q), (_, q64) in zip(mine.named_parameters(), ref.named_parameters(), ref64.named_parameters()):
        if q.grad is None: continue
        em, er = relerr(p.grad, q64.grad), relerr(q.grad, q64.grad)
        if not k.startswith("vit.blocks") or k.startswith("vit.blocks.11.mlp.linear1.w"):
            print(f"  {k:55s} mine {em:.2e}  oracle32 {er:.2e}  |g|max {q64.grad.abs().max().item():.2e}")

run("fp32", dict(in_channels=2), "logits")
run("fp32", dict(pos_embed="conv"), "logits")
