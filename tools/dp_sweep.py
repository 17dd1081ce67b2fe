"""N > 1: the configs[1] training step (eager launches, event-gated gradient all-reduce) under different NCCL communicator options,
all in ONE process per rank: every configuration gets its own process group (`dist.new_group(pg_options=...)`), so a sweep costs one
start-up.  Run under torchrun; rank 0 prints one line per configuration.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/dp_sweep.py
"""
import importlib
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dmedicalimagesegmentation_b200")
par = importlib.import_module("3dmedicalimagesegmentation_b200.parallel")
import bench  # noqa: E402  (MODEL_KW)

STEPS = int(os.environ.get("SWEEP_STEPS", "20"))


def main():
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)                      # NCCL banners go to stderr; results to the saved stdout
    rank, world, local = par.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    model = pkg.MonaiUNETR(**bench.MODEL_KW).to(dev).set_mode("bf16")
    loss_fn = pkg.DiceCELoss(to_onehot_y=True, softmax=True)
    opt = pkg.FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, mirror=model, capturable=True)
    g = torch.Generator().manual_seed(100 + rank)
    xs = [torch.rand(2, 1, 96, 96, 96, generator=g).to(dev) for _ in range(4)]
    ys = [torch.randint(0, 14, (2, 1, 96, 96, 96), generator=g).float().to(dev) for _ in range(4)]

    def step(ddp, i):
        loss = loss_fn(model(xs[i % 4]), ys[i % 4])
        loss.backward()
        ddp.reduce_and_step(opt)
        opt.zero_grad(set_to_none=True)

    def timed(ddp):
        for i in range(4):
            step(ddp, i)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(STEPS):
            step(ddp, i)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        return par.max_over_ranks(e0.elapsed_time(e1), world, dev) / STEPS

    def group(**cfg):
        if not cfg:
            return None
        o = dist.ProcessGroupNCCL.Options()
        for k, v in cfg.items():
            if k == "high_priority":
                o.is_high_priority_stream = bool(v)
            else:
                setattr(o.config, k, v)
        return dist.new_group(ranks=list(range(world)), pg_options=o)

    # the GPU drifts by ~0.1 ms over the sweep (clocks under sustained load): the default is measured first AND last
    configs = [
        ("default", {}),
        ("cta_policy=efficiency", {"cta_policy": 1}),
        ("max_ctas=24", {"max_ctas": 24}),
        ("max_ctas=16", {"max_ctas": 16}),
        ("nvls_ctas=8", {"nvls_ctas": 8}),
        ("high_priority", {"high_priority": 1}),
        ("default again", {}),
    ]
    only = os.environ.get("SWEEP_ONLY")
    out = []
    for name, cfg in configs:
        if only and name not in only.split(","):
            continue
        try:
            ddp = par.GradientAllReduce(model, world, group=group(**cfg))
            for order in ("deferred", "in_place") if name.startswith("default") else ("deferred",):
                model.defer_conv_wgrads = order == "deferred"
                ms = timed(ddp)
                out.append({"nccl": name, "wgrad_order": order, "ms_per_step": round(ms, 4), "samples_per_s": round(2 * world / ms * 1e3, 1)})
                if rank == 0:
                    print(json.dumps(out[-1]), file=sys.stderr, flush=True)
            model.defer_conv_wgrads = True
        except Exception as exc:               # an option this NCCL build rejects
            out.append({"nccl": name, "error": f"{type(exc).__name__}: {exc}"[:200]})
    if rank == 0:
        os.write(saved, (json.dumps({"n_gpus": world, "steps": STEPS, "results": out}) + "\n").encode())
    par.shutdown(world)


if __name__ == "__main__":
    main()
