"""Generates the golden fixtures in this directory (run once, here, on CPU; results are committed).

    python tests/golden/make_golden.py

The reference ships no golden vectors for this path and cannot be run (Gurobi/ROS2 absent), so the
expected values come from the NumPy oracle (oracle/hdsm_oracle.py): exact branch and bound with its
full-space interior-point solver, every optimum re-solved by HiGHS at the final assignment.  The
closed loops that lead to the snapshots are advanced with the C port for speed; the snapshots
themselves are solved by the NumPy oracle.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import c_oracle as co  # noqa: E402
from oracle import hdsm_oracle as o  # noqa: E402
from multi_agent_pkgs_b200 import scenarios as sc  # noqa: E402


def expected(batch, max_nodes=3000):
    p = o.Params(**batch.params)
    N = p.n_hor
    n = batch.n
    exp = dict(status=np.zeros(n, np.int32), obj=np.full(n, np.inf), traj=np.zeros((n, N + 1, 9)),
               ctrl=np.zeros((n, N, 3)), sigma=np.full((n, N), -1, np.int32), highs_obj=np.full(n, np.inf),
               kkt=np.full(n, np.inf), nodes=np.zeros(n, np.int32))
    for i in range(n):
        polys = batch.polys_of(i)
        lo, hi = batch.nbr_begin[i], batch.nbr_end[i]
        planes = o.time_aware_planes(p, batch.prev_self_pos[i], batch.all_pos[lo:hi], batch.all_valid[lo:hi],
                                     batch.global_id[i] - lo)
        r = o.solve_miqp_bnb(p, batch.x0[i], batch.ref[i], polys, planes, max_nodes=max_nodes)
        exp["status"][i], exp["nodes"][i] = r.status, r.nodes
        if r.status == o.OPTIMAL:
            exp["obj"][i], exp["traj"][i], exp["ctrl"][i], exp["sigma"][i] = r.obj, r.traj, r.ctrl, r.sigma
            exp["kkt"][i] = r.qp.kkt
            h = o.solve_qp_highs(o.build_qp_full(p, batch.x0[i], batch.ref[i], polys, planes, r.sigma))
            exp["highs_obj"][i] = h.obj
    return exp


def closed_loop(swarm, steps, snaps):
    out = {}
    for step in range(max(snaps) + 1):
        b = swarm.make_batch()
        if step in snaps:
            out[step] = b
        r = co.solve_batch(b, max_nodes=400)
        swarm.advance(r["traj"], r["ctrl"], r["res"]["status"] == 0)
    return out


def save(name, batch, exp):
    batch.save(os.path.join(HERE, name + "_in.npz"))
    np.savez_compressed(os.path.join(HERE, name + "_exp.npz"), **exp)
    ok = exp["status"] == 0
    gap = np.abs(exp["obj"][ok] - exp["highs_obj"][ok]) / np.maximum(1, np.abs(exp["obj"][ok]))
    print(f"{name}: n={batch.n} status={np.bincount(exp['status'], minlength=5)} nodes max {exp['nodes'].max()} "
          f"max |obj-highs| rel {gap.max() if ok.any() else 0:.2e} max kkt {exp['kkt'][ok].max() if ok.any() else 0:.2e}")


def main():
    # config 1: single agent, empty map, steps 0, 1, 5 (N = 10)
    for step, b in closed_loop(sc.config1_single_agent(), 6, {0, 1, 5}).items():
        save(f"config1_step{step}", b, expected(b))
    # config 2: 10-agent circle over the forest, steps 1, 8, 20
    for step, b in closed_loop(sc.config2_circle(), 21, {1, 8, 20}).items():
        save(f"config2_step{step}", b, expected(b))
    # shipped horizon N = 9
    sw = sc.config2_circle(seed=12, n_hor=9)
    for step, b in closed_loop(sw, 7, {6}).items():
        save(f"config2_n9_step{step}", b, expected(b))
    # RK4 + drag
    sw = sc.config2_circle(seed=13)
    sw.params.update(rk4=True, drag=(0.3, 0.1, 0.2))
    for step, b in closed_loop(sw, 6, {5}).items():
        save(f"config2_rk4_step{step}", b, expected(b))
    # config 3: line of agents through forest / wall / forest
    for step, b in closed_loop(sc.config3_line(), 13, {12}).items():
        save(f"config3_step{step}", b, expected(b))
    # dense 24-agent crossing at the centre (many active inter-agent planes)
    sw = sc.config2_circle(seed=14, n_rob=24, radius=6.0, centre=(-30.0, -30.0))
    for step, b in closed_loop(sw, 9, {8}).items():
        save(f"dense24_step{step}", b, expected(b))


if __name__ == "__main__":
    main()


def export_lp_files():
    """Gurobi-readable LP files of a few golden instances (tests/golden/lp/): `gurobi_cl Threads=1
    ResultFile=x.sol <file>.lp` on a licensed machine confirms `obj` / `traj` of <file>.expected.json."""
    from oracle import export_lp
    from multi_agent_pkgs_b200.scenarios import Batch
    os.makedirs(os.path.join(HERE, "lp"), exist_ok=True)
    for name, want in (("config1_step1", 0), ("config2_step8", 3), ("config3_step12", 2), ("dense24_step8", 5)):
        b = Batch.load(os.path.join(HERE, name + "_in.npz"))
        exp = dict(np.load(os.path.join(HERE, name + "_exp.npz")))
        ok = np.flatnonzero(exp["status"] == 0)
        i = int(ok[min(want, len(ok) - 1)])
        export_lp.export_agent(b, i, os.path.join(HERE, "lp", f"{name}_agent{i}"), exp)
