"""Key roofline metrics of every launch in an .ncu-rep (run here, no GPU needed): python tools/ncu_summary.py a.ncu-rep ..."""
import csv, subprocess, sys, io
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__inst_executed_pipe_uniform.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for path in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print(f"## {path}: {d.get('Kernel Name', '?')[:110]}")
        for k in KEYS:
            hit = [h for h in hdr if h.endswith(k)]
            for h in hit[:1]:
                print(f"  {k:72s} {d[h]:>16s} {u[h]}")
