#!/usr/bin/env python
"""Benchmark of the UNETR hot path (BASELINE.json metric: UNETR 96^3 fwd+bwd samples/s).

    python bench.py --gpus N --steps K --warmup W            # this framework, on N B200s (torchrun for N>1)
    python bench.py --impl reference --steps K --warmup W     # the reference arithmetic on the host CPU cores

Workload at N=1 = BASELINE.json configs[1]: UNETR(1->14, 96^3, feature 16, ViT-B) segmentation training step,
batch 2 per GPU, bf16 mode: forward + DiceCE + backward (+ AdamW step).  N>1 = data parallel, fixed per-GPU batch
(weak scaling), one gradient all-reduce per step.  One JSON line on stdout (rank 0).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE_96 = 379.72e9     # fwd+bwd, SURVEY 8(d) / Appendix A (3 x 126.57 GFLOP)
MODEL_KW = dict(in_channels=1, out_channels=14, img_size=(96, 96, 96), feature_size=16, hidden_size=768, mlp_dim=3072,
                num_heads=12, pos_embed="perceptron", norm_name="instance", res_block=True)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  The sampler is started before the
    warm-up (nvidia-smi needs ~1 s to come up, longer with 8 ranks starting one each) and every row is stamped on arrival; `stop`
    reports the rows that fall inside [mark_begin, mark_end] -- the timed region -- and, if the region was shorter than the
    sampling period, the rows of the whole loaded run (warm-up included), saying which."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc:
            self.proc.terminate()
        inside = [r for t, r in self.rows if self.t0 is not None and self.t1 is not None and self.t0 <= t <= self.t1 + 0.05]
        window = "timed region"
        if not inside:
            inside = [r for t, r in self.rows if self.t0 is None or t >= self.t0 - 2.0]
            window = "loaded run (warm-up + timed region): the timed region was shorter than one sampling period"
        sm = sorted(int(r[0]) for r in inside if r and r[0].isdigit())
        mx = max([int(r[1]) for r in inside if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in inside)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm), "window": window}


def cpu_sliding_window_ms_per_window():
    """The reference's whole-volume inference (oracle restatement of MONAI 0.6.0 `sliding_window_inference` around the oracle UNETR) on
    all host cores, on a BOUNDED sample: a 144x144x96 volume = 4 windows of 96^3 at overlap 0.5 = one predictor call of sw_batch_size 4.
    Returns milliseconds per window (the full 512x512x256 volume is 500 windows of the same cost)."""
    from oracle import unetr_oracle as O
    torch.set_num_threads(os.cpu_count())
    model = O.make_model(tuple_output=False).eval()
    vol = torch.rand(1, 1, 144, 144, 96, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        O.sliding_window_inference(vol, (96,) * 3, 4, model, overlap=0.5)
        t0 = time.perf_counter()
        O.sliding_window_inference(vol, (96,) * 3, 4, model, overlap=0.5)
        t = time.perf_counter() - t0
    return 1e3 * t / 4


def cpu_ranking_step_time(batch=4):
    """configs[2] on the host cores: oracle UNETR forward on `batch` crops of 96^3 -> enc4 -> the reference's 576-triplet Bradley-Terry loss
    in its own per-triplet form (oracle.bt_ranking_loss, rank:202-212) -> backward.  One untimed and one timed step."""
    import numpy as np
    from oracle import unetr_oracle as O
    torch.set_num_threads(os.cpu_count())
    model = O.make_model(tuple_output=True)
    x = torch.rand(batch, 1, 96, 96, 96, generator=torch.Generator().manual_seed(3))
    np.random.seed(0)
    t = None
    for i in range(2):
        t0 = time.perf_counter()
        model.zero_grad(set_to_none=True)
        enc4, _ = model(x)
        f1, f2 = torch.split(enc4, [batch // 2, batch // 2], dim=0)
        sd = 2 + i % 3
        loss = O.bt_ranking_loss(f1, f2, sd, O.slice_indices(f1.shape[sd]), 0.1)
        loss.backward()
        float(loss.detach())
        t = time.perf_counter() - t0
    return t


def cpu_reference_step_time(batch, steps, warmup):
    """The reference's arithmetic (oracle restatement of MONAI 0.6.0 UNETR + DiceCELoss; MONAI itself cannot be
    installed here) on all host cores: forward + DiceCE + backward on `batch` 96^3 crops."""
    from oracle import unetr_oracle as O
    torch.set_num_threads(os.cpu_count())
    model = O.make_model(tuple_output=False)
    x, y = O.make_inputs(batch=batch)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        model.zero_grad(set_to_none=True)
        loss = O.dice_ce_loss(model(x), y)
        loss.backward()
        float(loss.detach())
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def run_dp_check(pkg, par, model_kw, dev, rank, world):
    """Untimed correctness of the two sharded paths on the real hardware (N > 1).
    (a) data parallel, fp32 mode, one crop per rank: rank 0's all-reduced (AVG) gradients against the single-process gradient of
        the concatenated batch -- relative L2 per tensor; the two differ by fp32 summation order only;
    (b) slab-owned sliding window on a reduced volume (240x240x144: 32 windows, whole sw_batch chunks on 2/4/8 ranks): logits + mask
        against the single-GPU call, with a
        batch-composition-independent predictor (bit-equality required) and with the UNETR itself (reported: its forward rounds
        differently when a window sits in a different sw_batch chunk, so equality is not a property of the sharding)."""
    import torch.distributed as dist
    out = {}
    torch.manual_seed(1234)
    net = pkg.MonaiUNETR(**model_kw).to(dev).set_mode("fp32")
    loss_fn = pkg.DiceCELoss(to_onehot_y=True, softmax=True)
    S = model_kw["img_size"][0]
    gens = [torch.Generator().manual_seed(900 + r) for r in range(world)]
    xs = [torch.rand(1, 1, S, S, S, generator=g) for g in gens]
    ys = [torch.randint(0, 14, (1, 1, S, S, S), generator=g).float() for g in gens]
    red = par.GradientAllReduce(net, world)
    loss_fn(net(xs[rank].to(dev)), ys[rank].to(dev)).backward()
    red.reduce()
    torch.cuda.synchronize()
    mine = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    net.zero_grad(set_to_none=True)
    if rank == 0:
        loss_fn(net(torch.cat(xs).to(dev)), torch.cat(ys).to(dev)).backward()
        torch.cuda.synchronize()
        worst, worst_name = 0.0, ""
        for name, p in net.named_parameters():
            if p.grad is None:
                continue
            e = ((mine[name] - p.grad).norm() / p.grad.norm().clamp_min(1e-30)).item()
            if e > worst:
                worst, worst_name = e, name
        out["dp_grad_rel_l2_worst"] = worst
        out["dp_grad_worst_tensor"] = worst_name
        out["dp_grad_ok"] = bool(worst <= 2e-3)
    net.zero_grad(set_to_none=True)
    del red
    par.barrier(world)

    vol = torch.randn(1, 1, 240, 240, 144, generator=torch.Generator().manual_seed(77)).to(dev)
    wts = torch.linspace(-1.5, 1.5, 14, device=dev).view(1, 14, 1, 1, 1)

    def synth(t):      # elementwise: the value of a voxel does not depend on which windows share its predictor call
        return t * wts + 0.25 * t * t

    net.set_mode("bf16").eval()
    for tag, f in (("synthetic", synth), ("unetr_bf16", net)):
        with torch.no_grad():
            got, gmask = pkg.sliding_window_inference(vol, (96,) * 3, 4, f, overlap=0.5, rank=rank, world_size=world, return_argmax=True)
            if rank == 0:
                want, wmask = pkg.sliding_window_inference(vol, (96,) * 3, 4, f, overlap=0.5, return_argmax=True)
                out[f"sw_{tag}_bit_equal"] = bool(torch.equal(got, want) and torch.equal(gmask, wmask))
                out[f"sw_{tag}_max_abs_diff"] = (got - want).abs().max().item()
                out[f"sw_{tag}_mask_mismatch"] = int((gmask != wmask).sum().item())
                del want, wmask
            del got, gmask
        par.barrier(world)
    if rank == 0:
        out["sw_ok"] = out["sw_synthetic_bit_equal"]
    del net
    torch.cuda.empty_cache()
    return out if rank == 0 else None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.batch          # the GPU arm's per-step batch (configs[1]: 2 crops)
    t = cpu_reference_step_time(batch, args.steps, max(1, min(args.warmup, 1)))
    val = batch / t
    line = {"impl": "reference", "metric": "UNETR 96^3 fwd+bwd samples/s", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[1]: UNETR(1->14,96^3,fs16,ViT-B) segmentation training step (fwd+DiceCE+bwd), batch {batch}, CPU fp32 "
                                   "(oracle restatement of the MONAI 0.6.0 path on the host cores; the optimizer step is left out of this arm)",
                       "global_batch": batch},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{args.steps} steps x {batch} crops of 96^3, oracle restatement of the MONAI 0.6.0 path, torch {torch.__version__} CPU"},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=2, help="crops per GPU")
    ap.add_argument("--img", type=int, default=96, choices=[96, 128],
                    help="crop edge; 128 with --batch 4 is configs[3] (a parity-test configuration, measured on request; the metric line is configs[1])")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-optimizer", action="store_true")
    ap.add_argument("--torch-adamw", action="store_true", help="use torch.optim.AdamW(fused=True) instead of the package's FusedAdamW")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every step's launches from the host instead of replaying the captured CUDA graph")
    ap.add_argument("--breakdown", action="store_true", help="print the per-op CUDA-event breakdown to stderr")
    ap.add_argument("--no-sliding-window", action="store_true", help="skip the configs[4] whole-CT sliding-window measurement")
    ap.add_argument("--bf16-allreduce", action="store_true", help="N > 1: gradients cross NVLink as bf16 (cast, all-reduce, cast back on side streams)")
    ap.add_argument("--graph-dp", action="store_true", help="N > 1: capture the step including the NCCL all-reduces as a CUDA graph")
    ap.add_argument("--no-augment", action="store_true", help="skip the leg that feeds the step from the GPU-side crop sampler (SURVEY 8f N4)")
    ap.add_argument("--no-dp128", action="store_true", help="skip the configs[3] leg (128^3 crops, 4 per GPU)")
    ap.add_argument("--no-dp-check", action="store_true", help="skip the untimed multi-GPU correctness checks (N > 1)")
    ap.add_argument("--no-ranking", action="store_true", help="skip the configs[2] ranking pre-training step measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    pkg = importlib.import_module("3dmedicalimagesegmentation_b200")
    par = importlib.import_module("3dmedicalimagesegmentation_b200.parallel")
    # NCCL logs on stdout (the "NCCL version" banner at communicator creation, and everything NCCL_DEBUG=INFO prints, which the
    # driver reads for its rank check); the driver expects ONE JSON line there.  For the whole run file descriptor 1 points at
    # stderr, NCCL_DEBUG is left as the caller set it, and the JSON line is written to the saved descriptor at the end.
    sys.stdout.flush()
    saved_out = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = par.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    par.barrier(world)
    torch.cuda.synchronize()
    lib = pkg._lib.load()

    torch.manual_seed(0)
    S = args.img
    model_kw = dict(MODEL_KW, img_size=(S, S, S))
    flop_per_sample = FLOP_PER_SAMPLE_96 if S == 96 else 916.84e9      # 3 x 305.61 GFLOP at 128^3
    n_tok = (S // 16) ** 3
    model = pkg.MonaiUNETR(**model_kw).to(dev).set_mode(args.mode)
    loss_fn = pkg.DiceCELoss(to_onehot_y=True, softmax=True)
    if args.torch_adamw:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, fused=True)
    else:
        # one launch; keeps the packed bf16 weights current; update count on the device (graph replay)
        opt = pkg.FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, mirror=model, capturable=True)
    # gradients cross NVLink as fp32; --bf16-allreduce compresses them (parallel.GradientAllReduce(compress="bf16")) -- measured at N = 2:
    # 5.92 vs 5.67 ms/step (the two cast passes cost more than the halved all-reduce saves when the wire is not the limiter)
    compress = "bf16" if args.bf16_allreduce else None
    ddp = par.GradientAllReduce(model, world, compress=compress) if world > 1 else None
    if os.environ.get("B200_GRAD_GROUPS"):
        model.grad_groups = int(os.environ["B200_GRAD_GROUPS"])       # tuning aid: 4 / 7 / 13 gradient-ready events per backward
    B = args.batch
    g = torch.Generator().manual_seed(100 + rank)
    # 8 distinct host batches (pinned) cycled through: per-step inputs exceed nothing cached on the device side
    host_x = [torch.rand(B, 1, S, S, S, generator=g).pin_memory() for _ in range(4)]
    host_y = [torch.randint(0, 14, (B, 1, S, S, S), generator=g).float().pin_memory() for _ in range(4)]
    dev_x = [t.to(dev) for t in host_x]
    dev_y = [t.to(dev) for t in host_y]

    def eager_step(x, y):
        logits = model(x)
        loss = loss_fn(logits, y)
        loss.backward()
        if ddp and not args.no_optimizer and not args.torch_adamw:
            ddp.reduce_and_step(opt)      # AdamW of each gradient group behind that group's all-reduce
        else:
            if ddp:
                ddp.reduce()
            if not args.no_optimizer:
                opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    gstep, graph_note = None, "off (--no-graph)"
    if args.torch_adamw or args.no_optimizer:
        graph_note = "off (needs the capturable FusedAdamW)"
    elif world > 1 and not args.graph_dp:
        # NCCL collectives inside a replayed graph: works (measured at N=2) but one run in two hung on this pool's driver/NCCL pair;
        # with more than one rank the host enqueue (3.3 ms) hides behind the 6+ ms step anyway, so data parallel keeps eager launches
        graph_note = "off (data parallel: eager launches; --graph-dp to capture the all-reduce too)"
    elif not args.no_graph:
        try:
            # the whole step (forward, DiceCE, backward, gradient all-reduce, AdamW) captured once and replayed: graph.py
            gstep = pkg.GraphedTrainStep(model, loss_fn, opt, dev_x[0], dev_y[0], reducer=ddp, warmup=2)
            graph_note = "whole step replayed as one CUDA graph"
        except Exception as exc:      # capture refused (driver / NCCL): same kernels, launched from the host
            print(f"[bench] CUDA-graph capture of the step failed, running eager launches: {exc}", file=sys.stderr)
            torch.cuda.synchronize()
            opt.zero_grad(set_to_none=True)
            graph_note = f"off (capture failed: {type(exc).__name__})"

    def step(x, y):
        return gstep(x, y) if gstep is not None else eager_step(x, y)

    def timed(fn, steps):
        par.barrier(world)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        par.barrier(world)
        return par.max_over_ranks(e0.elapsed_time(e1), world, dev)

    sampler = ClockSampler(local)
    sampler.start()
    for i in range(args.warmup):
        step(dev_x[i % 4], dev_y[i % 4])
    torch.cuda.synchronize()
    sampler.mark_begin()
    l0 = lib.b200_launch_count()
    ms = timed(lambda i: step(dev_x[i % 4], dev_y[i % 4]), args.steps)
    launches = lib.b200_launch_count() - l0
    if gstep is not None:
        launches = gstep.launches_per_step * args.steps      # kernel nodes of the replayed graph (counted while it was captured)
    # end-to-end: every step's inputs come from pinned host memory and its loss is read back (a blocking .item()).  The copy of
    # batch i+1 is enqueued on a copy stream while step i computes (what a prefetching loader does; the reference's DataLoader
    # has pin_memory + 4 workers, seg:587) -- every copy and every read-back is inside the timed region.
    copy_stream = torch.cuda.Stream()
    slots = [(torch.empty_like(dev_x[0]), torch.empty_like(dev_y[0]), torch.cuda.Event()) for _ in range(2)]

    def prefetch(i):
        bx, by, ev = slots[i % 2]
        with torch.cuda.stream(copy_stream):      # slot i%2 was last read by step i-2, which finished before step i-1's .item()
            bx.copy_(host_x[i % 4], non_blocking=True)
            by.copy_(host_y[i % 4], non_blocking=True)
            ev.record(copy_stream)

    def e2e_step(i):
        bx, by, ev = slots[i % 2]
        torch.cuda.current_stream().wait_event(ev)
        loss = step(bx, by)
        prefetch(i + 1)
        return loss.item()
    prefetch(0)
    e2e_step(0)
    # steps are numbered 1..K here so that the slot primed by the warm-up call is the one step 1 reads
    ms_e2e = timed(lambda i: e2e_step(i + 1), args.steps)
    sampler.mark_end()
    clocks = sampler.stop()

    # per-op CUDA-event breakdown of one step -> roofline of the dominant kernel class
    lib.b200_prof_enable(1)
    eager_step(dev_x[0], dev_y[0])
    prof = pkg._lib.prof_report()
    lib.b200_prof_enable(0)
    pk, pk_kind = peaks()
    import re
    # kernel classes of the tcgen05 engine, algorithmic FLOPs from the op tags (2*M*N*K; 2*B*S^3*Cin*Cout*taps)
    classes = {"tc::gemm_kernel (ViT linear layers: fwd, dgrad, wgrad)": [0.0, 0.0, 0],
               "tc::conv_halo_kernel (conv3d 3^3 [+fused 1^3] implicit GEMM: fwd, dgrad)": [0.0, 0.0, 0],
               "tc::wgrad_halo_kernel / tc::wgrad_kernel (conv3d weight gradients)": [0.0, 0.0, 0],
               "tc::attn_kernel (fused attention: fwd; bwd = dS/dQ kernel + dV, dK GEMMs)": [0.0, 0.0, 0]}
    names = list(classes)
    for k, v in prof.items():
        m = re.match(r"linear_wgrad grouped x(\d+) mflop=(\d+)", k)
        if m:     # deferred weight gradients of a group of transformer blocks: one grouped launch, FLOPs carried in the tag
            c = classes[names[0]]; c[0] += v[0]; c[1] += v[1] * float(m.group(2)) * 1e6; c[2] += v[1]
            continue
        m = re.match(r"linear_(fwd|dgrad|wgrad) (\d+)x(\d+)x(\d+)", k)
        if m:
            c = classes[names[0]]; c[0] += v[0]; c[1] += v[1] * 2.0 * int(m.group(2)) * int(m.group(3)) * int(m.group(4)); c[2] += v[1]
            continue
        m = re.match(r"attention_(fwd|bwd)", k)
        if m:     # L = 216 tokens, 12 heads x 64: 2 (fwd) / 4 (bwd) products of 2*L*L*64 FLOPs per (sample, head)
            c = classes[names[3]]; c[0] += v[0]; c[2] += v[1]
            c[1] += v[1] * (2 if m.group(1) == "fwd" else 4) * 2.0 * B * 12 * n_tok * n_tok * 64
            continue
        m = re.match(r"conv_(fwd|dgrad) k3(\+k1)? (\d+)->(\d+) @(\d+)", k)
        if m:
            taps = 28 if m.group(2) else 27
            c = classes[names[1]]; c[0] += v[0]; c[2] += v[1]
            c[1] += v[1] * 2.0 * B * int(m.group(5)) ** 3 * int(m.group(3)) * int(m.group(4)) * taps
            continue
        m = re.match(r"conv_wgrad k(\d)(\+k1)? (\d+)x(\d+) @(\d+)", k)
        if m:     # "k3+k1": the residual block's 1^3 weight gradient formed by the same halo launch (one more tap)
            c = classes[names[2]]; c[0] += v[0]; c[2] += v[1]
            c[1] += v[1] * 2.0 * B * int(m.group(5)) ** 3 * int(m.group(3)) * int(m.group(4)) * (int(m.group(1)) ** 3 + (1 if m.group(2) else 0))
    # DRAM bytes per launch of the hot kernels: `ncu --set full` captures of tools/prof_kernels.py, summarised by tools/ncu_traffic.py
    # into profiles/r02_traffic.json together with the git revision of the build that was profiled
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        traffic = {}
    traffic_rev = traffic.get("_git_sha")
    def roof_of(name):
        ms_, fl, n = classes[name]
        r = {"bound": "tensor", "kernel": name, "launches": n, "avg_launch_us": 1e3 * ms_ / n if n else None,
             "achieved": fl / (ms_ * 1e-3) / 1e12 if ms_ else None, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
             "peak_kind": pk_kind + " sustained cuBLAS bf16 (kernel timed inside a long step)",
             "algorithmic": "sum over launches of 2*M*N*K (linear) / 2*B*S^3*Cin*Cout*taps (conv) / 2*L*L*64 per product, sample and head (attention)", "share_of_step_ms": ms_,
             "timing": "CUDA events around every launch of one step on the launching stream (b200_prof_*)"}
        r["frac"] = r["achieved"] / r["peak"] if r["achieved"] else None
        t = traffic.get(name.split(" ")[0])
        r["traffic"] = t["dram_bytes_per_launch"] if t else None
        if t:
            r["traffic_of"] = t["launch"]
            r["tensor_pipe_busy"] = t.get("tensor_pipe_busy")
            r["traffic_build"] = traffic_rev
        return r
    top_class = max(names, key=lambda nme: classes[nme][0])
    roof = roof_of(top_class)
    roof["other_kernels"] = [roof_of(nme) for nme in names if nme != top_class]
    # HBM-bound kernel classes: algorithmic bytes / CUDA-event time of the same profiled step
    V0 = float(S) ** 3
    act = 2 if args.mode == "bf16" else 4
    n_par = sum(p.numel() for p in model.parameters() if p.requires_grad)
    n_mir = sum(p.numel() for p in model.parameters() if p.dim() >= 2 and p.numel() >= 768 * 768) if args.mode == "bf16" else 0
    # InstanceNorm backward of the 5 residual blocks: final norm pair (4 reads + 4 reads + 2 writes) + first norm (2 + 2 + 1) = 15 passes
    in_bytes = B * 15 * act * sum(v * c for v, c in ((V0, 16), (V0, 16), (V0 / 8, 32), (V0 / 64, 64), (V0 / 512, 128)))
    hbm_classes = {
        "adamw_kernel (FusedAdamW: g,p,m,v read; p,m,v written; bf16 mirror of the GEMM weights written)": ("adamw", 28.0 * n_par + 2.0 * n_mir),
        "dicece_staged_kernel<fwd> (logits + labels read once)": ("dicece_fwd", B * V0 * (4 * 14 + 4)),
        "dicece_staged_kernel<bwd> (logits + labels read, dlogits written)": ("dicece_bwd", B * V0 * (8 * 14 + 4)),
        "in_bwd_reduce / in_bwd_apply (InstanceNorm + LeakyReLU backward, 5 residual blocks)": ("instnorm_bwd", in_bytes)}
    hbm = []
    for name, (tag, nbytes) in hbm_classes.items():
        if tag in prof and prof[tag][0] > 0:
            ms_ = prof[tag][0]
            hbm.append({"bound": "hbm", "kernel": name, "launches": prof[tag][1], "share_of_step_ms": ms_, "algorithmic_bytes": nbytes,
                        "achieved": nbytes / (ms_ * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": nbytes / (ms_ * 1e-3) / 1e9 / pk["hbm_gbs"],
                        "peak_kind": pk_kind + " STREAM-style copy", "timing": "CUDA events around the launches of one step (includes ~4 us of event overhead per launch)"})
    roof["hbm_kernels"] = hbm
    top = max(prof.items(), key=lambda kv: kv[1][0]) if prof else ("none", (0.0, 0))
    roof["top_op"] = top[0]; roof["top_op_ms"] = top[1][0]
    if args.breakdown and rank == 0:
        lib.b200_prof_enable(2)            # phase-level regions only (per-op events have a ~8 us floor each)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eager_step(dev_x[1], dev_y[1]); e1.record()
        phases = pkg._lib.prof_report()
        lib.b200_prof_enable(0)
        print(f"  phases of one step ({e0.elapsed_time(e1):.3f} ms incl. loss + AdamW): " +
              "  ".join(f"{k} {v[0]:.3f}" for k, v in sorted(phases.items())), file=sys.stderr)
        tot = sum(v[0] for v in prof.values())
        print(f"  profiled total {tot:.3f} ms", file=sys.stderr)
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            print(f"  {k:34s} {v[0]:9.3f} ms  {v[1]:4d} calls  {100 * v[0] / tot:5.1f}%", file=sys.stderr)

    # the same step fed from a device-resident CT volume by the GPU-side crop sampler + augmentation (seg:341-375):
    # RandCropByPosNegLabeld (one crop pair per step) + 3 flips + rot90 + intensity shift as one fused gather, then the training step
    aug = None
    if not args.no_augment and S == 96:
        Tm = pkg.transforms
        gv = torch.Generator().manual_seed(11 + rank)
        vol_i = torch.rand(1, 320, 320, 192, generator=gv).to(dev)
        vol_l = (torch.rand(1, 320, 320, 192, generator=gv) < 0.02).float().to(dev) * torch.randint(1, 14, (1, 320, 320, 192), generator=gv).float().to(dev)
        pipe = Tm.Compose([
            Tm.RandCropByPosNegLabeld(keys=["image", "label"], label_key="label", spatial_size=(96, 96, 96), pos=1, neg=1, num_samples=B,
                                      image_key="image", image_threshold=0),
            Tm.RandFlipd(keys=["image", "label"], spatial_axis=[0], prob=0.10), Tm.RandFlipd(keys=["image", "label"], spatial_axis=[1], prob=0.10),
            Tm.RandFlipd(keys=["image", "label"], spatial_axis=[2], prob=0.10), Tm.RandRotate90d(keys=["image", "label"], prob=0.10, max_k=3),
            Tm.RandShiftIntensityd(keys=["image"], offsets=0.10, prob=0.50)]).set_random_state(seed=rank)
        sample = {"image": vol_i, "label": vol_l}

        def aug_step(i):
            crops = pipe(sample)
            return step(torch.stack([c["image"] for c in crops]), torch.stack([c["label"] for c in crops]))
        for i in range(3):
            aug_step(i)
        l0 = lib.b200_launch_count()
        ms_aug = timed(aug_step, args.steps) / args.steps
        aug = {"value": B * world / (ms_aug * 1e-3), "unit": "samples/s", "ms_per_step": ms_aug, "steps": args.steps,
               "sampler_launches_per_step": (lib.b200_launch_count() - l0) // args.steps if gstep is not None else None,
               "workload": "configs[1] step fed by transforms.Compose([RandCropByPosNegLabeld(num_samples=batch), RandFlipd x3, RandRotate90d, "
                           "RandShiftIntensityd]) from a device-resident 320x320x192 volume (foreground/background index built once)"}
        del vol_i, vol_l, pipe, sample

    # untimed multi-GPU correctness checks; printed as "dp_check"
    dp_check = None
    if world > 1 and S == 96 and not args.no_dp_check:
        dp_check = run_dp_check(pkg, par, model_kw, dev, rank, world)

    # second half of the metric: sliding-window inference on a synthetic 512x512x256 CT (configs[4]), windows sharded over ranks:
    # every rank owns an x-slab of the accumulator, halo rows travel by NCCL send/recv, the uint8 mask is all-reduced, logits stay sharded
    sw = None
    if not args.no_sliding_window and S == 96:
        model.eval()
        vol = torch.rand(1, 1, 512, 512, 256, generator=torch.Generator().manual_seed(5)).to(dev)
        swi = lambda: pkg.sliding_window_inference(vol, (96,) * 3, 4, model, overlap=0.5, rank=rank, world_size=world, return_argmax=True,
                                                   gather_logits=False)
        with torch.no_grad():
            # warm-up on the FULL volume: the first call pays cudaMalloc of the accumulator / output buffers and NCCL's lazy set-up
            # of the point-to-point channels; steady state is what a validation loop over many volumes sees
            swi()
            torch.cuda.synchronize()
            t_sw = min(timed(lambda i: swi(), 1) for _ in range(2))
        sw = {"value": 1e3 / t_sw, "unit": "volumes/s", "ms_per_volume": t_sw, "windows": 500, "volume": "512x512x256", "roi": 96,
              "overlap": 0.5, "sw_batch_size": 4, "tflops_algorithmic": 63.29e12 / (t_sw * 1e-3) / 1e12,
              "outputs": "normalised logits (sharded by x-slab across ranks) + uint8 argmax mask (complete on every rank)",
              "sharding": None if world == 1 else "x-slab accumulators, halo pieces by NCCL send/recv, mask all-reduce (inferers.py)"}
        del vol
        model.train()

    # configs[3]: data-parallel training at 128^3 crops, 4 per GPU (same step: fwd + DiceCE + bwd + gradient all-reduce + AdamW)
    dp128 = None
    if not args.no_dp128 and S == 96:
        gstep = None
        m128 = pkg.MonaiUNETR(**dict(MODEL_KW, img_size=(128, 128, 128))).to(dev).set_mode(args.mode)
        o128 = pkg.FusedAdamW(m128.parameters(), lr=1e-4, weight_decay=1e-5, mirror=m128, capturable=True)
        d128 = par.GradientAllReduce(m128, world, compress=compress) if world > 1 else None
        x128 = [torch.rand(4, 1, 128, 128, 128, generator=g).to(dev) for _ in range(2)]
        y128 = [torch.randint(0, 14, (4, 1, 128, 128, 128), generator=g).float().to(dev) for _ in range(2)]

        def step128(i):
            loss = loss_fn(m128(x128[i % 2]), y128[i % 2])
            loss.backward()
            if d128:
                d128.reduce_and_step(o128)
            else:
                o128.step()
            o128.zero_grad(set_to_none=True)
        for i in range(3):
            step128(i)
        n128 = max(6, min(args.steps, 10))
        ms128 = timed(step128, n128) / n128
        dp128 = {"value": 4 * world / (ms128 * 1e-3), "unit": "samples/s", "ms_per_step": ms128, "batch_per_gpu": 4, "steps": n128,
                 "tflops_algorithmic": 4 * world * 916.84e9 / (ms128 * 1e-3) / 1e12,
                 "workload": "configs[3]: UNETR(1->14,128^3,fs16,ViT-B) training step, 4 crops per GPU, gradient all-reduce over NVLink (weak scaling)"}
        del m128, o128, d128, x128, y128
        torch.cuda.empty_cache()

    # configs[2]: ranking pre-training step (rank:238-274) -- batch 8 x 96^3, "feat" stage: full forward, 576 triplets of enc4
    # slices along one axis, Bradley-Terry loss, backward through encoder4 + ViT blocks 0-9 + patch embedding, AdamW step
    rk = None
    if not args.no_ranking and S == 96:
        gstep = None
        del model, opt, ddp
        torch.cuda.empty_cache()
        rmodel = pkg.UNETR(**MODEL_KW).to(dev).set_mode(args.mode)
        ropt = pkg.FusedAdamW(rmodel.parameters(), lr=1e-4, weight_decay=1e-5, mirror=rmodel)
        rddp = par.GradientAllReduce(rmodel, world, compress=compress) if world > 1 else None

        class _Opt:                       # BTLoss(reference, similar, dissimilar, optimizer) steps the optimizer itself (rank:213-215)
            def step(self):
                if rddp:
                    rddp.reduce_and_step(ropt)
                else:
                    ropt.step()

            def zero_grad(self, set_to_none=True):
                ropt.zero_grad(set_to_none=True)
        xs = [torch.rand(8, 1, 96, 96, 96, generator=g).to(dev) for _ in range(2)]
        import numpy as np
        np.random.seed(rank)

        def rank_step(i):
            enc4, _ = rmodel(xs[i % 2])
            f1, f2 = torch.split(enc4, [4, 4], dim=0)
            ref, sim, dis = pkg.extract_triplets_more_partitions(f1, f2, 2 + i % 3)
            return pkg.BTLoss(ref, sim, dis, _Opt())
        for i in range(3):
            rank_step(i)
        n_rk = max(6, min(args.steps, 12))
        ms_rk = timed(rank_step, n_rk) / n_rk
        rk = {"value": 8 * world / (ms_rk * 1e-3), "unit": "samples/s", "ms_per_step": ms_rk, "batch_per_gpu": 8, "steps": n_rk,
              "workload": "configs[2]: UNETR.forward -> enc4 -> extract_triplets_more_partitions (576 triplets, slice dims 2,3,4 in turn) "
                          "-> BTLoss (backward + AdamW inside, loss.item() read back every step)"}
        del rmodel, ropt, xs
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and sw is not None:
        # the sliding-window half of the metric on the host cores (bounded sample, extrapolated by window count: every window costs the same)
        try:
            mspw = cpu_sliding_window_ms_per_window()
            sw["cpu_baseline"] = {"value": 1e3 / (mspw * 500), "unit": "volumes/s", "ms_per_window": mspw, "cores": os.cpu_count(), "kind": "port",
                                  "sample": "oracle sliding_window_inference on a 144x144x96 volume (4 windows of 96^3 = one predictor call), per-window time x 500 windows"}
        except Exception as exc:      # never lose the bench line to its side measurement
            sw["cpu_baseline"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline and rk is not None:
        try:
            t_rk = cpu_ranking_step_time(4)
            rk["cpu_baseline"] = {"value": 4 / t_rk, "unit": "samples/s", "s_per_step": t_rk, "cores": os.cpu_count(), "kind": "port",
                                  "sample": "1 timed step (after 1 untimed) of batch 4 (the reference's own batch, rank:251): oracle UNETR forward -> enc4 -> per-triplet BTLoss form of the reference -> backward, no optimizer step"}
        except Exception as exc:
            rk["cpu_baseline"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t = cpu_reference_step_time(1, 5, 1)
        cpu = {"value": 1 / t, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "5 timed steps x 1 crop of 96^3 (fwd+DiceCE+bwd), oracle restatement of the MONAI 0.6.0 path on the host CPU"}
    if rank == 0:
        samples = B * world * args.steps
        line = {"metric": f"UNETR {S}^3 fwd+bwd samples/s", "value": samples / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.mode, "data": "synthetic",
                "config": {"workload": ("configs[1]" if S == 96 else "configs[3]") + f": UNETR(1->14,{S}^3,fs16,ViT-B) segmentation training step (fwd+DiceCE+bwd" +
                           ("" if args.no_optimizer else "+AdamW") + f"), batch {B}/GPU", "global_batch": B * world,
                           "parallelism": f"dp{world}", "gradient_allreduce": (None if world == 1 else (compress or "fp32")), "launch": graph_note, "l2": "4 rotating input batches; activations per step (~1 GB) exceed the 126 MB L2"},
                "tflops_algorithmic": samples * flop_per_sample / (ms * 1e-3) / 1e12,
                "e2e": {"value": samples / (ms_e2e * 1e-3), "unit": "samples/s",
                        "h2d_bytes_per_step": host_x[0].numel() * 4 + host_y[0].numel() * 4, "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "sliding_window": sw, "ranking_step": rk, "dp_128": dp128, "augmented_step": aug, "dp_check": dp_check}
        sys.stdout.flush()
        os.write(saved_out, (json.dumps(line) + "\n").encode())
    par.shutdown(world)


if __name__ == "__main__":
    main()
