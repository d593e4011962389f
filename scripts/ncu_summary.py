"""Summarise an .ncu-rep (raw page + SASS source page joined with nvdisasm line info) - runs on CPU.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-instantiation, default 10 | corridor | map | reftraj | sense] [lib.so]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep = sys.argv[1]
ninst = sys.argv[2] if len(sys.argv) > 2 else "10"
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "multi_agent_pkgs_b200", "libhdsm.so")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
LAUNCH = int(os.environ.get("LAUNCH", "0"))  # which profiled launch of the report (0 = first)
hdr, vals = rows[0], rows[2 + LAUNCH]
print("=" * 100)
for name in ("Kernel Name", "Grid Size", "Block Size", "launch__cluster_dim_x"):
    if name in hdr:
        print(f"{name:70s} {vals[hdr.index(name)]}")
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "launch__grid_size"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:70s} {vals[i]} {rows[1][i]}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]  # one block per profiled launch
s0 = starts[LAUNCH]
end0 = starts[LAUNCH + 1] if len(starts) > LAUNCH + 1 else len(rows)
hdr, data = rows[s0 + 1], [r for r in rows[s0 + 2:end0] if len(r) == len(rows[s0 + 1])]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("static instructions", len(data), "samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
print("stalls:", ", ".join(f"{s[6:]} {100 * v / tot:.1f}%" for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
sym = {"corridor": ".text._ZN8hdsm_cor15corridor_kernel", "map": ".text._ZN7hdsm_mp10map_kernel",
       "reftraj": ".text._ZN7hdsm_rt14reftraj_kernel", "sense": ".text._ZN7hdsm_sn12sense_kernel"}.get(ninst, f".text._ZN4hdsm17hdsm_solve_kernelILi{ninst}E")
srcname = {"corridor": "hdsm_corridor.cu", "map": "hdsm_map.cu", "reftraj": "hdsm_reftraj.cu", "sense": "hdsm_sense.cu"}.get(ninst, "hdsm_kernel.cuh")
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):  # one cubin per .cu file
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    if any(l.startswith(sym) for l in dis):
        break
start = next(i for i, l in enumerate(dis) if l.startswith(sym))
end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith("\t.section")), len(dis))
cur, per = None, []
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        per.append(cur)
if len(per) != len(data):
    print("WARNING: disassembly / profile length mismatch", len(per), len(data))
n = min(len(per), len(data))
static, samp, execd = collections.Counter(), collections.Counter(), collections.Counter()
for i in range(n):
    static[per[i]] += 1
    samp[per[i]] += int(data[i][ix["# Samples"]])
    execd[per[i]] += int(data[i][ix["Instructions Executed"]])
CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi_agent_pkgs_b200", "csrc")
src = open(os.path.join(CSRC, srcname)).read().split("\n")
print("--- top source lines by stall samples")
for k, v in samp.most_common(int(os.environ.get("TOP", "30"))):
    s = src[k[1] - 1].strip()[:100] if k and k[0] == srcname else ""
    if k and not s and os.path.exists(os.path.join(CSRC, k[0])):
        s = open(os.path.join(CSRC, k[0])).read().split("\n")[k[1] - 1].strip()[:100]
    print(f"{str(k):30s} static {static[k]:5d} samples {100 * v / tot:5.1f}% exec {execd[k] / 1e6:8.1f}M | {s}")
