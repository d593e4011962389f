"""Profiling driver: a config-5 closed loop (pre-roll with the host producers, then device-only replay).
    ncu ... -k regex:hdsm_solve_kernel -s SKIP -c 2 python scripts/profile_loop.py 4096 [steps] [width] [warm_start]
Solver launches per step: first pass + cluster pass per row-pool tier (2 tiers -> 4); the replay of step s starts at solver
launch 4 * (steps + s)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_agent_pkgs_b200 import scenarios as sc

agents = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
width = int(sys.argv[3]) if len(sys.argv) > 3 else 4
warm = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
sw = sc.config5_random(seed=5, n_rob=4096, side=200.0)
pool = sc.InputPool(sw)
import torch
from multi_agent_pkgs_b200.swarm import ClosedLoop

loop = ClosedLoop(sw, 1, 0, "cuda:0", 64, pool, width=width, warm_start=warm)
loop.preroll(steps)
pool.close()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
loop.reset()
with torch.cuda.stream(loop.stream):
    for s in range(steps):
        ev[s].record(loop.stream)
        loop.device_step(s)
    ev[steps].record(loop.stream)
torch.cuda.synchronize()
print("replay ms per step", [round(ev[s].elapsed_time(ev[s + 1]), 2) for s in range(steps)])
