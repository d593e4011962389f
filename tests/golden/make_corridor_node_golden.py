"""Regenerates tests/golden/corridor_node_ref.npz from the REFERENCE ITSELF: the reference's own Agent::GenerateSafeCorridor
(multi_agent_planner/src/agent_class.cpp:1236-1447; agent_class.cpp and convex_decomp.cpp compiled unmodified into
oracle/_ref/libref_agent.so, `make -C oracle ref`) on the agents of the 10-agent forest scenario: a first update, and a follow-up
update after the agents moved, with the first update's polytopes, seeds, synthetic used flags and a previous plan supplied.

    python tests/golden/make_corridor_node_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from multi_agent_pkgs_b200 import corridor as cr, scenarios as sc  # noqa: E402
from oracle import hdsm_oracle as ho, ref_agent as ra  # noqa: E402


def first_batch():
    return cr.corridor_batch(sc.config2_circle(n_swarms=1))


def follow_up_batch(cb, out):
    """The same agents 0.6 m further along their paths, with `out` (the first update's result) as the kept polytopes."""
    rng = np.random.default_rng(5)
    n, P = cb.n, cb.poly_hor
    used = np.zeros((n, P), np.uint8)
    for i in range(n):
        k = int((out["poly_rows"][i] > 0).sum())
        used[i, :k] = rng.random(k) < 0.6
        used[i, 0] = 1
    traj = np.zeros((n, 11, 3))
    for i in range(n):
        d = cb.path[i, 0] - cb.pos[i]
        d = d / max(np.linalg.norm(d), 1e-9)
        traj[i] = cb.pos[i] + np.outer(np.linspace(0.0, 3.0 if i % 2 else 0.8, 11), d)   # odd agents leave their last polytope
    cb.with_previous(out, used, traj)
    cb.pos = traj[:, 2].copy()
    return cb


def run_reference(cb, i, with_prev):
    ag = ra.RefAgent(ho.Params(poly_hor=cb.poly_hor), cb.n, i, [0.0] * 9)
    gi = i if cb.grid_index is None else int(cb.grid_index[i])
    prev = None
    if with_prev:
        k = int(cb.prev_n[i])
        prev = dict(rows=cb.prev_rows[i, :k], A=cb.prev_A[i, :k], b=cb.prev_b[i, :k], seeds=cb.prev_seeds[i, :k], used=cb.prev_used[i, :k])
    return ag.safe_corridor(cb.grids[gi], cb.origins[i], cb.voxel, cb.pos[i], cb.path[i, :cb.n_path[i]], cb.n_it, cb.use_cvx_new, prev,
                            cb.prev_traj[i] if with_prev else np.zeros((0, 3)), cb.rmax)


def collect(cb, with_prev):
    n, P, R = cb.n, cb.poly_hor, cb.rmax
    out = dict(poly_rows=np.zeros((n, P), np.int32), poly_A=np.zeros((n, P, R, 3)), poly_b=np.zeros((n, P, R)), seeds=np.zeros((n, P, 3)))
    for i in range(n):
        _, out["poly_rows"][i], out["poly_A"][i], out["poly_b"][i], out["seeds"][i] = run_reference(cb, i, with_prev)
    return out


def main():
    cb = first_batch()
    a = collect(cb, False)
    cb2 = follow_up_batch(cb, a)
    b = collect(cb2, True)
    path = os.path.join(ROOT, "tests", "golden", "corridor_node_ref.npz")
    np.savez_compressed(path, **{f"first_{k}": v for k, v in a.items()}, **{f"next_{k}": v for k, v in b.items()})
    print(f"wrote 2 x {cb.n} corridor updates to {path} ({os.path.getsize(path) / 1024:.0f} KB); polytopes per agent "
          f"{(a['poly_rows'] > 0).sum(1).tolist()} -> {(b['poly_rows'] > 0).sum(1).tolist()}")


if __name__ == "__main__":
    main()
