"""Profiling driver: one 10 240-agent launch of the config-2 replica workload (K2-dominated: 10 neighbour candidates per agent).
    ncu ... -k regex:hdsm_solve_kernel -s 3 -c 1 python scripts/profile_config2.py [width]
Solver launches: 3 per call (row-pool tiers); call 0 = launches 0-2 (warm-up), call 1 starts at launch 3."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multi_agent_pkgs_b200 import scenarios as sc
from oracle import c_oracle as co
width = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sw = sc.config2_circle(seed=2, n_swarms=24)
for step in range(12):
    b = sw.make_batch()
    out = co.solve_batch(b, max_nodes=64)
    st = out["res"]["status"]
    sw.advance(out["traj"], out["ctrl"], (st == 0) | ((st == 4) & np.isfinite(out["res"]["obj"])))
b = sw.make_batch()
t = bench.tile_batch(b, 43).take(np.arange(10240))
n_rob = int(t.nbr_end.max())
t.all_pos, t.all_valid = t.all_pos[:n_rob], t.all_valid[:n_rob]
import torch
from multi_agent_pkgs_b200.planner import TrajectoryPlanner
from multi_agent_pkgs_b200.swarm import DeviceBatch
dev = torch.device("cuda:0")
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
pl = TrajectoryPlanner(sw.params, t.n, 10, max_nodes=64, width=width)
db = DeviceBatch(t, dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for k in range(3):
    ev[k].record(st)
    pl.solve_batch_device(db.t, db.n_rob, st.cuda_stream)
ev[3].record(st)
torch.cuda.synchronize()
r = db.results()
print("ms per call", [round(ev[k].elapsed_time(ev[k + 1]), 3) for k in range(3)], "iters", r["iters"].mean(), "max", r["iters"].max(), np.bincount(r["status"], minlength=6))
