"""GPU box: how much of a launch is the tail of a few expensive agents?  Times the bench snapshots in
their natural order and with the agents sorted by their own iteration count (descending = longest first)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multi_agent_pkgs_b200 import scenarios as sc
from multi_agent_pkgs_b200.planner import TrajectoryPlanner
from multi_agent_pkgs_b200.swarm import DeviceBatch
n_swarms = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
params = sc.agile_params(10)
gen = TrajectoryPlanner(params, max_agents=240, max_neighbours=10, max_nodes=64)
snaps = bench.make_snapshots(gen.solve_batch, 2, n_swarms)
gen.close()
dev = torch.device("cuda:0")
pl = TrajectoryPlanner(params, max_agents=n_swarms * 10, max_neighbours=10, max_nodes=64)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
def timed(b):
    db = DeviceBatch(b, dev)
    for _ in range(2):
        pl.solve_batch_device(db.t, db.n_rob, stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); pl.solve_batch_device(db.t, db.n_rob, stream.cuda_stream); e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), db.results()
for k, b in enumerate(snaps):
    t_nat, res = timed(b)
    order = np.argsort(-res["iters"], kind="stable")
    t_lpt, _ = timed(b.take(order))
    t_rev, _ = timed(b.take(order[::-1]))
    rng = np.random.default_rng(0)
    t_rnd, _ = timed(b.take(rng.permutation(b.n)))
    print(f"snapshot {k}: natural {t_nat:.1f} ms, longest-first {t_lpt:.1f} ms, shortest-first {t_rev:.1f} ms, random {t_rnd:.1f} ms; "
          f"iters mean {res['iters'].mean():.1f} max {res['iters'].max()}")
# history-based order: dispatch by the iteration counts of the PREVIOUS snapshot of the same agents
prev = None
allres = []
for b in snaps:
    _, r = timed(b)
    allres.append(r["iters"].copy())
for k, b in enumerate(snaps):
    hist = allres[k - 1]
    order = np.argsort(-hist, kind="stable")
    t_hist, _ = timed(b.take(order))
    t_nat, _ = timed(b)
    c = np.corrcoef(np.log1p(hist), np.log1p(allres[k]))[0, 1]
    print(f"snapshot {k}: natural {t_nat:.1f} ms, by previous snapshot's iterations {t_hist:.1f} ms (log-corr {c:.2f})")
