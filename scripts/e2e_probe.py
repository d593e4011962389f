"""GPU box: end-to-end time of hdsm_solve_batch (host buffers) on the bench workload for the current
HDSM_CHUNKS / HDSM_ONE_STREAM environment.   python scripts/e2e_probe.py [swarms]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multi_agent_pkgs_b200 import scenarios as sc
from multi_agent_pkgs_b200.planner import TrajectoryPlanner
n_swarms = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
params = sc.agile_params(10)
gen = TrajectoryPlanner(params, max_agents=240, max_neighbours=10, max_nodes=64)
snaps = bench.make_snapshots(gen.solve_batch, 2, n_swarms)
gen.close()
pl = TrajectoryPlanner(params, max_agents=n_swarms * 10, max_neighbours=10, max_nodes=64)
for k in range(2):
    pl.solve_batch(snaps[k % 4])
ts = []
for k in range(8):
    t0 = time.perf_counter(); pl.solve_batch(snaps[k % 4]); ts.append(time.perf_counter() - t0)
print(f"chunks={os.environ.get('HDSM_CHUNKS','auto')} one_stream={os.environ.get('HDSM_ONE_STREAM','0')}: "
      f"{1e3*np.mean(ts):.1f} ms/step, {n_swarms*10/np.mean(ts):.0f} solves/s (min {1e3*min(ts):.1f} ms)")
