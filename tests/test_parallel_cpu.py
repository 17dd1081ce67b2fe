"""world_size-2 gloo tests of the multi-GPU host logic (no GPU): gradient averaging equals the single-process
gradient of the concatenated batch, the flat-buffer fast path is taken, and window shards partition the list."""
import importlib
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    par = importlib.import_module("3dmedicalimagesegmentation_b200.parallel")
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    par.init_from_env("gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 6, generator=g)
    # per-rank half of the batch; mean loss per rank -> averaged grads == grads of the global mean loss
    xr = x[rank * 2:(rank + 1) * 2]
    net(xr).square().mean().backward()
    # (a) generic bucket path
    red = par.GradientAllReduce(net, world, bucket_mb=1)
    red.reduce()
    bucket = [p.grad.clone() for p in net.parameters()]
    # (b) flat-buffer fast path: re-home grads as consecutive views of one buffer (what the UNETR autograd node does)
    net.zero_grad(set_to_none=True)
    net(xr).square().mean().backward()
    flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    o = 0
    for p in net.parameters():
        p.grad = flat[o:o + p.numel()].view_as(p)
        o += p.numel()
    assert red._flat_view([p.grad for p in net.parameters()]) is not None
    red.reduce()
    fast = [p.grad.clone() for p in net.parameters()]
    t = par.max_over_ranks(float(rank + 1), world, "cpu")
    if rank == 0:
        torch.save({"bucket": bucket, "fast": fast, "max": t}, out)
    par.barrier(world)
    par.shutdown(world)


def test_gradient_allreduce_two_ranks_gloo(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29000 + os.getpid() % 2000
    mp.start_processes(_worker, args=(2, port, out), nprocs=2, join=True, start_method="spawn")
    got = torch.load(out)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 6, generator=g)
    net(x).square().mean().backward()
    for p, a, b in zip(net.parameters(), got["bucket"], got["fast"]):
        assert torch.allclose(p.grad, a, atol=1e-6) and torch.allclose(p.grad, b, atol=1e-6)
    assert got["max"] == 2.0


def _slab_worker(rank, world, port, out):
    """Halo exchange of the slab-owned sliding window over gloo: every rank fills its send buffers with (window, row) codes laid out as
    the plan says and checks that what arrives is exactly the list of pieces the plan promises, in window order."""
    sys.path.insert(0, ROOT)
    par = importlib.import_module("3dmedicalimagesegmentation_b200.parallel")
    inf = importlib.import_module("3dmedicalimagesegmentation_b200.inferers")
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    par.init_from_env("gloo")
    size, roi = (64, 32, 16), (16, 16, 16)
    _, flat = inf.window_starts(size, roi, 0.5)
    chunks, bounds, pieces = inf.slab_plan(flat, roi[0], size[0], world)
    code = lambda w, x: float(w * 1000 + x)                  # one value per (window, padded row); a piece row carries it `per` times
    per = 3
    sends = {}
    for d in range(world):
        if d == rank:
            continue
        mine = [p for p in pieces[d] if p[1] == rank]
        if mine:
            sends[d] = torch.tensor([code(w, x) for (w, _, lo, hi) in mine for x in range(lo, hi) for _ in range(per)])
    sizes = {}
    for (w, src, lo, hi) in pieces[rank]:
        if src != rank:
            sizes[src] = sizes.get(src, 0) + (hi - lo) * per
    recv = inf._exchange_nccl(sends, sizes, "cpu", None)
    ok = True
    for src in sizes:
        want = [code(w, x) for (w, s, lo, hi) in pieces[rank] if s == src for x in range(lo, hi) for _ in range(per)]
        ok = ok and recv[src].tolist() == want
    flags = [None] * world
    dist.all_gather_object(flags, (ok, len(sizes)))
    if rank == 0:
        torch.save(flags, out)
    par.barrier(world)
    par.shutdown(world)


def test_slab_halo_exchange_three_ranks_gloo(tmp_path):
    out = str(tmp_path / "slab.pt")
    port = 31000 + os.getpid() % 2000
    mp.start_processes(_slab_worker, args=(3, port, out), nprocs=3, join=True, start_method="spawn")
    flags = torch.load(out)
    assert all(f[0] for f in flags)
    assert all(f[1] >= 1 for f in flags)                                       # every rank receives halo pieces
