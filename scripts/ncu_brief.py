"""Brief text summary of every profiled launch in an .ncu-rep (raw page): duration, DRAM traffic, pipes, occupancy, stalls.
    python scripts/ncu_brief.py file.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "launch__cluster_dim_x", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    print("=" * 100)
    for k in want:
        if k in d:
            print(f"{k:66s} {d[k]} {u.get(k, '')}")
    st = {k: float(v.replace(",", "")) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")}
    tot = sum(st.values()) or 1.0
    print("stall reasons (share of stalled-warp samples): " + ", ".join(
        f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:7]))
