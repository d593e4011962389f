"""GPU box: per-agent latency of the solver (one block alone on an SM) from a hard golden instance."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden
from multi_agent_pkgs_b200.planner import TrajectoryPlanner
b, exp = load_golden("config2_step20")
hard = int(np.argmax(exp["nodes"]))
for reps in (1, 148, 148 * 4, 148 * 9):
    bb = b.take([hard] * reps)
    pl = TrajectoryPlanner(b.params, max_agents=reps, max_neighbours=10, max_nodes=5000)
    pl.solve_batch(bb)
    t = []
    for _ in range(5):
        t0 = time.perf_counter(); out = pl.solve_batch(bb); t.append(time.perf_counter() - t0)
    r = out["res"][0]
    print(f"reps {reps:5d}: best {min(t)*1e3:8.3f} ms  iters {r['iters']} nodes {r['nodes']}  -> {min(t)*1e6/r['iters']:.2f} us per IPM iteration (wall, incl. copies)")
    pl.close()
