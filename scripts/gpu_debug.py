"""Debug aid (GPU box): closed-loop config-2 swarm, CUDA library vs the C oracle, per-agent report."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle as co
from multi_agent_pkgs_b200 import scenarios as sc
from multi_agent_pkgs_b200.planner import TrajectoryPlanner

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
nsw = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sw = sc.config2_circle(n_swarms=nsw)
pl = TrajectoryPlanner(sw.params, max_agents=sw.n, max_neighbours=10, max_nodes=400)
print("smem bytes", pl.smem_bytes)
worst = 0
for step in range(nsteps):
    b = sw.make_batch()
    ref = co.solve_batch(b, max_nodes=400)
    t = time.time(); out = pl.solve_batch(b); dt = time.time() - t
    r0, r1 = ref["res"], out["res"]
    both = (r0["status"] == 0) & (r1["status"] == 0)
    gap = np.abs(r0["obj"] - r1["obj"]) / np.maximum(1, np.abs(r0["obj"]))
    dtraj = np.abs(ref["traj"] - out["traj"]).reshape(b.n, -1).max(1)
    mism = np.nonzero(r0["status"] != r1["status"])[0]
    print(f"step {step}: gpu {dt*1e3:.2f} ms  status ref {np.bincount(r0['status'], minlength=6)} gpu {np.bincount(r1['status'], minlength=6)} "
          f"max gap {gap[both].max() if both.any() else -1:.2e} max dtraj {dtraj[both].max() if both.any() else -1:.2e} "
          f"iters ref {r0['iters'].sum()} gpu {r1['iters'].sum()} nodes ref {r0['nodes'].sum()} gpu {r1['nodes'].sum()} mism {mism[:8]}")
    if both.any():
        worst = max(worst, gap[both].max())
    for i in mism[:4]:
        print("   agent", i, "ref", r0[i], "gpu", r1[i])
    sw.advance(ref["traj"], ref["ctrl"], r0["status"] == 0)
print("worst gap", worst)
