"""Regenerates tests/golden/agent_model_ref.npz from the REFERENCE ITSELF.

Needs oracle/_ref/libref_agent.so: the reference's own multi_agent_planner/src/agent_class.cpp compiled unmodified on the stand-in
ROS / Eigen headers and the RECORDING stand-in for the Gurobi C++ API (`make -C oracle ref`, only where /root/reference is mounted).
Every case is one call of the reference's GenerateTimeAwareSafeCorridor + SolveOptimizationProblem on random inputs: the polytopes
with the inter-agent planes it appended (poly_const_final_vec_) and the model it handed to Gurobi (objective, bounds, dynamics rows,
one-hot rows, indicator rows).

    python tests/golden/make_agent_model_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hdsm_oracle as ho, ref_agent as ra  # noqa: E402

CASES = [  # (Params overrides, n_rob, id, number of polytopes, have previous plan)
    (dict(), 4, 1, 2, True),
    (dict(), 6, 0, 3, False),
    (dict(n_hor=9, poly_hor=3), 3, 2, 3, True),
    (dict(rk4=True, drag=(0.3, 0.3, 0.1)), 5, 4, 1, True),
    (dict(n_hor=6, poly_hor=2, dt=0.2, r_u=0.05, r_x=(50.0, 60.0, 70.0, 2.0, 3.0, 4.0), r_n=(500.0, 400.0, 300.0, 5.0, 6.0, 7.0), max_vel=9.5,
          min_acc_xy=-20.0, max_acc_xy=25.0, min_acc_z=-10.0, max_acc_z=12.0, max_jerk=40.0, drone_radius=0.3, drone_z_offset=0.45), 8, 3, 4, True),
]


def random_polytope(rng, centre):
    rows = rng.integers(6, 13)
    A = rng.normal(size=(rows, 3))
    A[:6] = np.repeat(np.eye(3), 2, 0) * np.tile([1.0, -1.0], 3)[:, None]
    b = A @ centre + rng.uniform(1.0, 3.0, rows)
    return A, b


def make_case(rng, over, n_rob, aid, n_poly, have_prev):
    p = ho.Params(**over)
    N = p.n_hor
    x0 = np.r_[rng.uniform(-2, 2, 3), rng.uniform(-3, 3, 3), rng.uniform(-2, 2, 3)]
    ref = np.zeros((N + 1, 6))
    ref[:, :3] = x0[:3] + np.outer(np.arange(1, N + 2), rng.uniform(-0.5, 0.5, 3))
    ref[:, 3:] = rng.uniform(-5, 5, 3)
    polys = [random_polytope(rng, x0[:3] + q * rng.uniform(-1, 1, 3)) for q in range(n_poly)]
    prev = None
    if have_prev:
        prev = np.zeros((N + 1, 9))
        prev[:, :3] = x0[:3] + np.outer(np.arange(N + 1), rng.uniform(-0.4, 0.4, 3))
    state_ini = np.r_[rng.uniform(-2, 2, 3), np.zeros(6)]
    all_pos = rng.uniform(-4, 4, (n_rob, N + 1, 3))
    all_valid = (rng.random(n_rob) < 0.8).astype(np.uint8)
    return p, x0, ref, polys, prev, state_ini, all_pos, all_valid


def main():
    rng = np.random.default_rng(20261018)
    out = {"n_cases": np.array(len(CASES))}
    for c, (over, n_rob, aid, n_poly, have_prev) in enumerate(CASES):
        p, x0, ref, polys, prev, state_ini, all_pos, all_valid = make_case(rng, over, n_rob, aid, n_poly, have_prev)
        ag = ra.RefAgent(p, n_rob, aid, state_ini)
        r = ag.step(x0, ref, polys, prev, all_pos, all_valid)
        out[f"c{c}_over"] = np.array(repr(over))
        out[f"c{c}_meta"] = np.array([n_rob, aid, n_poly, int(have_prev)])
        out[f"c{c}_x0"], out[f"c{c}_ref"], out[f"c{c}_state_ini"] = x0, ref, state_ini
        out[f"c{c}_prev"] = prev if prev is not None else np.zeros((0, 9))
        out[f"c{c}_all_pos"], out[f"c{c}_all_valid"] = all_pos, all_valid
        for q, (A, b) in enumerate(polys):
            out[f"c{c}_polyA{q}"], out[f"c{c}_polyb{q}"] = A, b
        for k in range(p.n_hor):
            for q in range(n_poly):
                out[f"c{c}_finA{k}_{q}"], out[f"c{c}_finb{k}_{q}"] = r["final"][k][q]
        for key in ("obj_diag", "obj_lin", "lb", "ub", "vtype"):
            out[f"c{c}_{key}"] = r[key]
        out[f"c{c}_obj_const"], out[f"c{c}_obj_offdiag"], out[f"c{c}_failed"] = np.array(r["obj_const"]), np.array(r["obj_offdiag"]), np.array(r["failed"])
        out[f"c{c}_lin"], out[f"c{c}_lin_c"], out[f"c{c}_lin_s"] = r["lin"]
        out[f"c{c}_ind"], out[f"c{c}_ind_c"], out[f"c{c}_ind_b"] = r["ind"]
    path = os.path.join(ROOT, "tests", "golden", "agent_model_ref.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {len(CASES)} cases to {path} ({os.path.getsize(path) / 1024:.0f} KB)")


if __name__ == "__main__":
    main()
