"""Debug helper: CUDA path vs C port on one config-5 slice with the search devices switched off one by one."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_agent_pkgs_b200 import scenarios as sc
from multi_agent_pkgs_b200.planner import TrajectoryPlanner
from oracle import c_oracle as co

sw = sc.config5_random(seed=11, n_rob=300, side=60.0)
for step in range(3):
    b = sw.make_batch()
    ref = co.solve_batch(b, max_nodes=64)
    for dbg in (0, 1, 2, 3):
        os.environ["HDSM_DEBUG"] = str(dbg)
        for warps in (4, 1):
            os.environ["HDSM_WARPS"] = str(warps)
            pl = TrajectoryPlanner(sw.params, b.n, b.n, max_nodes=64)
            out = pl.solve_batch(b)
            pl.close()
            mis = out["res"]["status"] != ref["res"]["status"]
            ok = (out["res"]["status"] == 0) & (ref["res"]["status"] == 0)
            gap = np.abs(out["res"]["obj"][ok] - ref["res"]["obj"][ok]) / np.maximum(1, np.abs(ref["res"]["obj"][ok]))
            print(f"step {step} dbg {dbg} warps {warps}: status mismatches {int(mis.sum())} gpu {np.bincount(out['res']['status'], minlength=6)} "
                  f"port {np.bincount(ref['res']['status'], minlength=6)} max gap {gap.max() if ok.any() else 0:.2e} iters gpu {out['res']['iters'].mean():.1f} port {ref['res']['iters'].mean():.1f}", flush=True)
            if mis.any() and dbg == 0 and warps == 4:
                i = int(np.nonzero(mis)[0][0])
                print("  first mismatch agent", i, out["res"][i], ref["res"][i], "rows", b.poly_rows[i])
    sw.advance(ref["traj"], ref["ctrl"], (ref["res"]["status"] == 0))
