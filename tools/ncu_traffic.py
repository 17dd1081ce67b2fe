"""profiles/r0N_traffic.json from `ncu --set full` captures (run here, no GPU needed):
    python tools/ncu_traffic.py gpurun_out/r02_hot.ncu-rep > profiles/r02_traffic.json
One entry per kernel class bench.py reports (last capture of each kernel wins), stamped with the git revision of the tree."""
import csv, io, json, subprocess, sys
CLASSES = [("tc::gemm_kernel", "gemm_kernel", "gemm 432x3072x768 / 432x768x3072 (ViT linear layers)"),
           ("tc::conv_halo_kernel", "conv_halo48_kernel", "conv_halo48 16->16 @96^3 x2 (3x3x3 forward with InstanceNorm statistics)"),
           ("tc::wgrad_halo_kernel", "wgrad_halo_kernel", "wgrad_halo 16x16 @96^3 x2"),
           ("tc::attn_kernel", "attn_kernel", "fused attention forward, 2 x 12 heads x 216 tokens (48 CTAs)"),
           ("tc::gemm_grouped_kernel", "gemm_grouped_kernel", "grouped weight gradients of 4 ViT-B blocks (16 problems, 864 tiles)"),
           ("dicece_staged_kernel<fwd>", "dicece_staged_kernel<16, 0>", "DiceCE forward 2 x 14 x 96^3"),
           ("dicece_staged_kernel<bwd>", "dicece_staged_kernel<16, 1>", "DiceCE backward 2 x 14 x 96^3")]
prev = {}
out = {"_git_sha": subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip(), "_source": sys.argv[1:]}
for path in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    col = lambda key: next(i for i, h in enumerate(hdr) if h.endswith(key))
    unit = rows[1]
    def val(r, key):
        i = col(key); u = unit[i].lower()
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            return float("nan")
        return v * {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1, "usecond": 1e-6, "us": 1e-6, "msecond": 1e-3, "ms": 1e-3, "nsecond": 1e-9, "ns": 1e-9}.get(u, 1)
    for r in rows[2:]:
        name = r[col("Kernel Name")]
        for key, pat, desc in CLASSES:
            if pat in name:
                out[key] = {"launch": desc, "kernel": name[:100], "dram_bytes_per_launch": int(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")),
                            "duration_us_under_ncu": round(val(r, "gpu__time_duration.sum") * 1e6, 1),
                            # the metric /opt/skills/guides/B200_PROFILING.md names: tensor pipe active cycles as a fraction of the launch
                            # (rounds before r02's final summary divided sm__pipe_tensor_subpipe_hmma_cycles_active_realtime by
                            # sm__cycles_elapsed, which is not a fraction -- it exceeds 1 -- and overstated the pipe occupancy)
                            "tensor_pipe_busy": round(val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") / 100, 3)}
                if out[key]["tensor_pipe_busy"] != out[key]["tensor_pipe_busy"]:      # metric not collected in this pass
                    out[key]["tensor_pipe_busy"] = prev.get(key, {}).get("tensor_pipe_busy")
                prev[key] = dict(out[key])
print(json.dumps(out, indent=1))
