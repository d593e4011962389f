"""Profiling driver: a config-5 closed loop (pre-roll with the host producers, then device-only replay).
    ncu ... python scripts/profile_loop.py [agents] [preroll_steps] [width]
Each step launches the solver once per row-pool tier (3); the replay of step s starts at solver launch 3 * (preroll + s)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_agent_pkgs_b200 import scenarios as sc

agents = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
width = int(sys.argv[3]) if len(sys.argv) > 3 else 4
sw = sc.config5_random(seed=5, n_rob=4096, side=200.0)
if agents < 4096:   # a shard of the 4096-agent swarm, as one of 4096 / agents ranks would see it (others' plans frozen at their starts)
    pass
pool = sc.InputPool(sw)
import torch
from multi_agent_pkgs_b200.swarm import ClosedLoop

class Shard(ClosedLoop):
    pass

loop = ClosedLoop(sw, 1, 0, "cuda:0", 64, pool, width=width)
loop.preroll(steps)
pool.close()
if agents < 4096:
    # re-solve the first `agents` agents of the last step as a small batch (what a rank of a sharded run launches)
    from multi_agent_pkgs_b200.planner import TrajectoryPlanner
    t = dict(loop._bind(steps - 1))
    small = {k: (v[:agents].contiguous() if k not in ("all_pos", "all_valid") and hasattr(v, "shape") and v.shape[0] == loop.n else v) for k, v in t.items()}
    pl = TrajectoryPlanner(sw.params, agents, 4096, 0, max_nodes=64, width=width)
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(3):
        pl.solve_batch_device(small, loop.n_rob, st.cuda_stream)
    torch.cuda.synchronize()
    print("small batch done", agents)
else:
    loop.reset()
    with torch.cuda.stream(loop.stream):
        for s in range(steps):
            loop.device_step(s)
    torch.cuda.synchronize()
    print("replay done")
