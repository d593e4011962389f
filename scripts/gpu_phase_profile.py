"""Per-phase cycle counters of the solver kernel (library built with HDSM_ENABLE_PROFILE=1, run with HDSM_PROFILE=1)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["HDSM_PROFILE"] = "1"
from multi_agent_pkgs_b200 import scenarios as sc
from multi_agent_pkgs_b200.planner import TrajectoryPlanner
from oracle import c_oracle as co
sw = sc.config5_random(seed=11, n_rob=300, side=60.0)
for step in range(4):
    b = sw.make_batch()
    base = co.solve_batch(b, max_nodes=64)
    if step == 3:
        pl = TrajectoryPlanner(sw.params, b.n, b.n, max_nodes=64, width=1)
        out = pl.solve_batch(b)
        print("iters per agent", out["res"]["iters"].mean(), "max", out["res"]["iters"].max())
        pl.close()
    sw.advance(base["traj"], base["ctrl"], (base["res"]["status"] == 0))
