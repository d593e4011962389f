"""Debug helper: device-pointer path vs C port (config-2 tiles with and without dispatch order, ClosedLoop)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_agent_pkgs_b200 import scenarios as sc
from multi_agent_pkgs_b200.planner import TrajectoryPlanner
from multi_agent_pkgs_b200.swarm import DeviceBatch, ClosedLoop
from oracle import c_oracle as co
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

dev = torch.device("cuda:0")
sw = sc.config2_circle(seed=2, n_swarms=24)
for step in range(7):
    b = sw.make_batch()
    ref = co.solve_batch(b, max_nodes=64)
    sw.advance(ref["traj"], ref["ctrl"], ref["res"]["status"] == 0)
b = sw.make_batch()
ref = co.solve_batch(b, max_nodes=64)
print("port", np.bincount(ref["res"]["status"], minlength=6), ref["res"]["iters"].mean())
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
for reps in (1, 10):
    t = bench.tile_batch(b, reps) if reps > 1 else b
    if reps > 1:
        n_rob = int(t.nbr_end.max()); t.all_pos, t.all_valid = t.all_pos[:n_rob], t.all_valid[:n_rob]
    pl = TrajectoryPlanner(sw.params, t.n, 10, max_nodes=64)
    db = DeviceBatch(t, dev)
    for call in range(3):
        pl.solve_batch_device(db.t, db.n_rob, stream.cuda_stream)
        torch.cuda.synchronize()
        r = db.results()
        want = np.tile(ref["res"]["status"], reps)
        print(f"device path reps {reps} call {call}: gpu {np.bincount(r['status'], minlength=6)} mismatches {(r['status'] != want).sum()} iters {r['iters'].mean():.1f}", flush=True)
    h = pl.solve_batch(t)
    print(f"host path reps {reps}: gpu {np.bincount(h['res']['status'], minlength=6)} mismatches {(h['res']['status'] != np.tile(ref['res']['status'], reps)).sum()}", flush=True)
    pl.close()
sw5 = sc.config5_random(seed=11, n_rob=300, side=60.0)
loop = ClosedLoop(sw5, 1, 0, "cuda:0", 64, None)
loop.checker = lambda bb: co.solve_batch(bb, max_nodes=64)
print("closed loop parity", loop.preroll(6, parity_sample=300), [np.bincount(s["status"], minlength=6).tolist() for s in loop.stats])
