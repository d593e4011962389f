"""GPU box: the other BASELINE.json configs (parity-test cases, not bench lines) - timing + parity summary.

    python scripts/extra_configs.py [config4|config5|config1|config3] [steps]

Closed loop on the GPU through hdsm_solve_batch (host buffers); every step is also solved by the C port
of the oracle for the parity columns.  Prints one JSON line per config.
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_agent_pkgs_b200 import scenarios as sc
from multi_agent_pkgs_b200.planner import TrajectoryPlanner
from oracle import c_oracle as co

which = sys.argv[1] if len(sys.argv) > 1 else "config4"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
sw = {"config1": sc.config1_single_agent, "config3": sc.config3_line, "config4": sc.config4_circle256,
      "config5": lambda: sc.config5_random(n_rob=4096)}[which]()
nn = int((sw.group_end - sw.group_begin).max())
pl = TrajectoryPlanner(sw.params, max_agents=sw.n, max_neighbours=nn, max_nodes=64)
gpu_ms, cpu_ms, gaps, kkts, mism, rows, nodes = [], [], [], [], 0, 0, []
for step in range(steps):
    b = sw.make_batch()
    pl.solve_batch(b)  # warm
    t0 = time.perf_counter(); out = pl.solve_batch(b); gpu_ms.append(1e3 * (time.perf_counter() - t0))
    t0 = time.perf_counter(); ref = co.solve_batch(b, max_nodes=64); cpu_ms.append(1e3 * (time.perf_counter() - t0))
    r0, r1 = ref["res"], out["res"]
    mism += int((r0["status"] != r1["status"]).sum())
    both = (r0["status"] == 0) & (r1["status"] == 0)
    if both.any():
        gaps.append(float((np.abs(r0["obj"][both] - r1["obj"][both]) / np.maximum(1, np.abs(r0["obj"][both]))).max()))
        kkts.append(float(r1["kkt_res"][both].max()))
    rows = max(rows, int(r1["rows"].max()))
    nodes.append(float(r1["nodes"].mean()))
    ok = (r1["status"] == 0) | ((r1["status"] == 4) & np.isfinite(r1["obj"]))
    sw.advance(out["traj"], out["ctrl"], ok)
print(json.dumps({"config": which, "agents": sw.n, "neighbour_candidates": nn, "steps": steps,
                  "gpu_e2e_ms_per_step_median": float(np.median(gpu_ms)), "gpu_solves_per_s": sw.n / (np.median(gpu_ms) * 1e-3),
                  "cpu_port_ms_per_step_median": float(np.median(cpu_ms)), "cpu_solves_per_s": sw.n / (np.median(cpu_ms) * 1e-3),
                  "cpu_threads": co.max_threads(), "status_mismatches_vs_c_port": mism, "max_rel_objective_gap": max(gaps) if gaps else None,
                  "max_kkt_residual": max(kkts) if kkts else None, "max_rows_after_pruning": rows,
                  "mean_relaxations_per_agent": float(np.mean(nodes)), "smem_bytes": pl.smem_bytes}))
