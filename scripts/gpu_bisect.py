"""Debug helper: CUDA path vs C port on one config-5 slice for every search width and cluster size."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_agent_pkgs_b200 import scenarios as sc
from multi_agent_pkgs_b200.planner import TrajectoryPlanner
from oracle import c_oracle as co

sw = sc.config5_random(seed=11, n_rob=300, side=60.0)
for step in range(4):
    b = sw.make_batch()
    base = co.solve_batch(b, max_nodes=64)
    if step >= 2:
        for width in (1,):
            ref = co.solve_batch(b, max_nodes=64, width=width)
            for csize, dbg in ((1, 0), (1, 4)):
                os.environ["HDSM_DEBUG"] = str(dbg)
                if csize > width:
                    continue
                os.environ["HDSM_CLUSTER"] = str(csize)
                pl = TrajectoryPlanner(sw.params, b.n, b.n, max_nodes=64, width=width)
                out = pl.solve_batch(b)
                t0 = time.perf_counter(); out = pl.solve_batch(b); dt = time.perf_counter() - t0
                pl.close()
                mis = out["res"]["status"] != ref["res"]["status"]
                ok = (out["res"]["status"] == 0) & (ref["res"]["status"] == 0)
                gap = np.abs(out["res"]["obj"][ok] - ref["res"]["obj"][ok]) / np.maximum(1, np.abs(ref["res"]["obj"][ok]))
                for i in np.nonzero(mis)[0][:6]:
                    print("   agent", i, "gpu", out["res"][i], "port", ref["res"][i])
                print(f"step {step} dbg {dbg} width {width} csize {csize}: mismatches {int(mis.sum())} gap {gap.max():.1e} nodes equal {np.array_equal(out['res']['nodes'], ref['res']['nodes'])} "
                      f"iters gpu {out['res']['iters'].mean():.1f} port {ref['res']['iters'].mean():.1f} max {out['res']['iters'].max()} wall {1e3 * dt:.2f} ms", flush=True)
    sw.advance(base["traj"], base["ctrl"], (base["res"]["status"] == 0))
