"""The optimisation's DATA pinned against the reference's own planner node.

tests/golden/agent_model_ref.npz holds what the reference's OWN Agent::GenerateTimeAwareSafeCorridor (agent_class.cpp:1086-1215)
and Agent::SolveOptimizationProblem (:858-1023, on the model of CreateGurobiModel :2071-2153) produce on random inputs -
agent_class.cpp compiled unmodified on stand-in ROS / Eigen headers, with a recording stand-in for the Gurobi C++ API in place of
the closed-source solver (tests/golden/make_agent_model_golden.py).  oracle/hdsm_oracle.py - the specification the C port and the
CUDA kernels are tested against - must reproduce: the inter-agent planes, the objective, the variable bounds, the dynamics rows,
the one-hot rows and every indicator row.  What Gurobi then does with that model stays unpinned (KKT certificate)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import hdsm_oracle as ho

GRB_INF = 1e100


def cases():
    z = np.load(os.path.join(GOLDEN, "agent_model_ref.npz"))
    for c in range(int(z["n_cases"])):
        over = eval(str(z[f"c{c}_over"]), {"__builtins__": {}})
        n_rob, aid, n_poly, have_prev = (int(v) for v in z[f"c{c}_meta"])
        p = ho.Params(**over)
        g = lambda k: z[f"c{c}_{k}"]  # noqa: E731
        yield dict(p=p, n_rob=n_rob, id=aid, x0=g("x0"), ref=g("ref"), state_ini=g("state_ini"), prev=g("prev") if have_prev else None,
                   all_pos=g("all_pos"), all_valid=g("all_valid"), polys=[(g(f"polyA{q}"), g(f"polyb{q}")) for q in range(n_poly)],
                   final=[[(g(f"finA{k}_{q}"), g(f"finb{k}_{q}")) for q in range(n_poly)] for k in range(p.n_hor)],
                   obj_diag=g("obj_diag"), obj_lin=g("obj_lin"), obj_const=float(g("obj_const")), obj_offdiag=float(g("obj_offdiag")),
                   lb=g("lb"), ub=g("ub"), vtype=g("vtype"), lin=(g("lin"), g("lin_c"), g("lin_s")), ind=(g("ind"), g("ind_c"), g("ind_b")),
                   failed=bool(g("failed")))


def check_against_oracle(c, r):
    """r: what the reference produced (fixture or live); c: the inputs."""
    p, N, P = c["p"], c["p"].n_hor, len(c["polys"])
    nx, nzc = 9 * (N + 1), 9 * (N + 1) + 3 * N
    P_eff = min(P, p.poly_hor)
    # --- inter-agent planes (GenerateTimeAwareSafeCorridor + AddHyperplane): appended to EVERY polytope, neighbour-id order
    prev_pos = c["prev"][:, :3] if c["prev"] is not None else np.tile(c["state_ini"][:3], (N + 1, 1))
    planes = ho.time_aware_planes(p, prev_pos, c["all_pos"], c["all_valid"], c["id"])
    for k in range(N):
        for q in range(P):
            A, b = r["final"][k][q]
            R = len(c["polys"][q][1])
            assert np.array_equal(A[:R], c["polys"][q][0]) and np.array_equal(b[:R], c["polys"][q][1])
            assert A.shape[0] == R + len(planes[k][1])
            assert np.allclose(A[R:], planes[k][0], rtol=0, atol=1e-14) and np.allclose(b[R:], planes[k][1], rtol=1e-14, atol=1e-14), (k, q)
    # --- objective (:870-883 on obj_grb_ of :2098): diagonal, no cross terms, linear part, constant
    qp = ho.build_qp_full(p, c["x0"], c["ref"][:N], c["polys"], planes, [0] * N, drop_constant_rows=False)
    assert r["obj_offdiag"] == 0.0
    assert np.array_equal(r["obj_diag"][:nzc], qp.Pdiag) and not r["obj_diag"][nzc:].any()
    assert np.allclose(r["obj_lin"][:nzc], qp.q, rtol=1e-15, atol=0) and not r["obj_lin"][nzc:].any()
    assert abs(r["obj_const"] - qp.c0) <= 1e-12 * max(1.0, abs(qp.c0))
    # --- variables: x0 fixed through its bounds (:886-889), terminal velocity / acceleration fixed at 0 (:2078-2081), boxes, binaries
    lb, ub = r["lb"], r["ub"]
    assert len(lb) == nzc + N * p.poly_hor
    assert np.array_equal(lb[:9], c["x0"]) and np.array_equal(ub[:9], c["x0"])
    xl, xu = np.where(np.isinf(p.x_lb()), -GRB_INF, p.x_lb()), np.where(np.isinf(p.x_ub()), GRB_INF, p.x_ub())
    for k in range(1, N):
        assert np.array_equal(lb[9 * k:9 * k + 9], xl) and np.array_equal(ub[9 * k:9 * k + 9], xu)
    assert np.array_equal(lb[9 * N:9 * N + 3], xl[:3]) and np.array_equal(ub[9 * N:9 * N + 3], xu[:3])
    assert not lb[9 * N + 3:nx].any() and not ub[9 * N + 3:nx].any()
    assert np.array_equal(lb[nx:nzc], np.tile(p.u_lb(), N)) and np.array_equal(ub[nx:nzc], np.tile(p.u_ub(), N))
    assert (r["vtype"][:nzc] == ord("C")).all() and (r["vtype"][nzc:] == ord("B")).all()
    # --- linear rows: 9 N dynamics equalities (:2146-2151) then N one-hot rows (:939-940)
    L, Lc, Ls = r["lin"]
    assert L.shape[0] == 9 * N + N and (Ls == ord("=")).all()
    dyn = qp.Aeq[9:9 + 9 * N]                      # the oracle's rows, between its x0 rows and its terminal rows
    assert np.allclose(L[:9 * N, :nzc], dyn, rtol=1e-14, atol=1e-16) and not L[:9 * N, nzc:].any() and not Lc[:9 * N].any()
    for k in range(N):
        row = L[9 * N + k]
        want = np.zeros_like(row)
        want[nzc + k * p.poly_hor:nzc + k * p.poly_hor + P_eff] = 1.0
        assert np.array_equal(row, want) and Lc[9 * N + k] == -1.0
    # --- indicator rows (:909-937 with GetGurobiPolyhedronConstraints :1071-1084): for step k, polytope q, every final row on x_k and on
    #     x_{k+1}, position components only, switched by binary b[k][q]
    I, Ic, Ib = r["ind"]
    n = 0
    for k in range(N):
        for q in range(P_eff):
            A, b = r["final"][k][q]
            for i in range(len(b)):
                for kk in (k, k + 1):
                    want = np.zeros(I.shape[1])
                    want[9 * kk:9 * kk + 3] = A[i]
                    assert np.array_equal(I[n], want) and Ic[n] == -b[i] and Ib[n] == nzc + k * p.poly_hor + q, (k, q, i, kk)
                    n += 1
    assert n == I.shape[0]
    # --- and the same rows are what the oracle's QP imposes for the assignment "polytope 0 everywhere"
    rows0 = sum(2 * len(r["final"][k][0][1]) for k in range(N))
    assert qp.C.shape[0] == 12 * (N - 1) + 6 * N + rows0
    # without a solver the reference reports a failed optimisation (GRBException -> :988-995)
    assert r["failed"]


def test_oracle_reproduces_the_reference_models_of_the_fixture():
    n = 0
    for c in cases():
        check_against_oracle(c, c)
        n += 1
    assert n == 5


def test_live_against_the_compiled_reference_node():
    from oracle import ref_agent as ra
    if not ra.have_ref():
        pytest.skip("oracle/_ref/libref_agent.so not built (needs /root/reference)")
    for c in cases():
        ag = ra.RefAgent(c["p"], c["n_rob"], c["id"], c["state_ini"])
        r = ag.step(c["x0"], c["ref"], c["polys"], c["prev"], c["all_pos"], c["all_valid"])
        check_against_oracle(c, r)
        for key in ("obj_diag", "obj_lin", "lb", "ub"):          # and the fixture is what the reference produces today
            assert np.array_equal(r[key], c[key]), key
        # a second step on the same node: the reference removes and re-adds its per-step rows (:892-902)
        r2 = ag.step(c["x0"] + 0.01, c["ref"], c["polys"], c["prev"], c["all_pos"], c["all_valid"])
        assert r2["lin"][0].shape == r["lin"][0].shape and r2["ind"][0].shape == r["ind"][0].shape
        assert np.array_equal(r2["lb"][:9], c["x0"] + 0.01)


# ---- the reference-trajectory generator against the reference's own node ---------------------------------------------
def _reftraj_checker(rb, i, prev_ref, increment):
    """oracle/reftraj_oracle.c on agent i of the batch with the given previous reference."""
    import copy
    from oracle import reftraj as ort
    r = copy.copy(rb)
    r.prev_ref, r.have_prev, r.increment = rb.prev_ref.copy(), rb.have_prev.copy(), rb.increment.copy()
    if prev_ref is not None:
        r.prev_ref[i], r.have_prev[i] = prev_ref, 1
    else:
        r.have_prev[i] = 0
    r.increment[i] = increment
    out = ort.c_generate(r)
    return out["ref"][i], out["path_vel"][i]


def test_reference_trajectory_checker_equals_the_reference_node():
    """GenerateReferenceTrajectory (agent_class.cpp:1449-1553 + SamplePath, KeepOnlyFreeReference, ComputePathVelocity) of the reference's
    own node (fixture tests/golden/reftraj_node_ref.npz, and live where the reference is present) against oracle/reftraj_oracle.c:
    first call, follow-up call and follow-up call with increment_traj_ref_ - equal to the last bit (both run on the host's libm)."""
    import sys
    sys.path.insert(0, os.path.join(GOLDEN))
    import make_reftraj_node_golden as mk
    from oracle import ref_agent as ra
    z = np.load(os.path.join(GOLDEN, "reftraj_node_ref.npz"))
    rb = mk.load_batch(z)   # the batch the fixture was generated on (stored with it: independent of the solver's tie-breaking)
    for i in (int(v) for v in z["agents"]):
        first, v0 = _reftraj_checker(rb, i, None, 0)
        assert np.array_equal(first, z[f"a{i}_first"]) and v0 == float(z[f"a{i}_first_vel"]), i
        for inc in (0, 1):
            nxt, v = _reftraj_checker(rb, i, first[:, :3], inc)
            assert np.array_equal(nxt, z[f"a{i}_next{inc}"]) and v == float(z[f"a{i}_next{inc}_vel"]), (i, inc)
    if ra.have_ref():                       # live: every agent of the scenario
        for i in range(rb.n):
            got, vel = mk.run_reference(rb, i, None, 0)
            want, wv = _reftraj_checker(rb, i, None, 0)
            assert np.array_equal(got, want) and vel == wv, i
            got2, vel2 = mk.run_reference(rb, i, got[:, :3], 1)
            want2, wv2 = _reftraj_checker(rb, i, got[:, :3], 1)
            assert np.array_equal(got2, want2) and vel2 == wv2, i


# ---- the safe-corridor walk against the reference's own node ---------------------------------------------------------
def test_corridor_walk_checker_equals_the_reference_node():
    """GenerateSafeCorridor (agent_class.cpp:1236-1447: kept polytopes, walk along the path, GetPolyOcta3D, A / b conversion) of the
    reference's own node (fixture tests/golden/corridor_node_ref.npz, and live where the reference is present) against
    oracle/corridor_oracle.c: a first update and a follow-up update with kept polytopes - every array bit for bit."""
    import sys
    sys.path.insert(0, os.path.join(GOLDEN))
    import make_corridor_node_golden as mk
    from oracle import corridor as oc, ref_agent as ra
    z = np.load(os.path.join(GOLDEN, "corridor_node_ref.npz"))
    cb = mk.first_batch()
    first = oc.c_safe_corridor(cb)
    for k in ("poly_rows", "poly_A", "poly_b", "seeds"):
        assert np.array_equal(first[k], z[f"first_{k}"]), k
    assert not first["flags"].any()
    cb2 = mk.follow_up_batch(cb, first)
    nxt = oc.c_safe_corridor(cb2)
    for k in ("poly_rows", "poly_A", "poly_b", "seeds"):
        assert np.array_equal(nxt[k], z[f"next_{k}"]), k
    kept = [(nxt["seeds"][i, 0] == first["seeds"][i]).all(1).any() for i in range(cb.n)]
    assert any(kept) and (nxt["poly_rows"][:, 0] >= 6).all()           # kept polytopes lead the new list (:1245-1270)
    if ra.have_ref():
        live = mk.collect(cb2, True)
        for k in ("poly_rows", "poly_A", "poly_b", "seeds"):
            assert np.array_equal(live[k], nxt[k]), k


def test_c_port_planes_equal_the_reference_planes():
    """oracle/hdsm_oracle.c builds the inter-agent planes itself (orc_plane, the same arithmetic as kernel K1): compare it directly with
    the planes the reference's own GenerateTimeAwareSafeCorridor appended (fixture), not only through the NumPy oracle."""
    import dataclasses
    from oracle import c_oracle as co
    n = 0
    for c in cases():
        p, N = c["p"], c["p"].n_hor
        prm = dataclasses.asdict(p)
        prev_pos = c["prev"][:, :3] if c["prev"] is not None else np.tile(c["state_ini"][:3], (N + 1, 1))
        for k in range(N):
            A, b = c["final"][k][0]
            R = len(c["polys"][0][1])
            row = R
            for j in range(c["n_rob"]):
                if j == c["id"] or not c["all_valid"][j]:
                    continue
                nrm, off = co.plane(prm, prev_pos[k + 1], c["all_pos"][j, k + 1])
                assert np.allclose(nrm, A[row], rtol=0, atol=1e-14) and abs(off - b[row]) <= 1e-14 * max(1.0, abs(b[row])), (k, j)
                row += 1
                n += 1
            assert row == len(b)
    assert n > 100


# ---- solutions checked inside the reference's own model ---------------------------------------------------------------
def _batch_of(c):
    """The fixture case as a scenarios.Batch (the arrays hdsm_solve_batch takes)."""
    import dataclasses
    from multi_agent_pkgs_b200 import scenarios as sc
    p, N = c["p"], c["p"].n_hor
    P, R = p.poly_hor, 18
    pA, pb, pr = np.zeros((1, P, R, 3)), np.zeros((1, P, R)), np.zeros((1, P), np.int32)
    for q, (A, b) in enumerate(c["polys"][:P]):
        pA[0, q, :len(b)], pb[0, q, :len(b)], pr[0, q] = A, b, len(b)
    prev = c["prev"][:, :3] if c["prev"] is not None else np.tile(c["state_ini"][:3], (N + 1, 1))
    return sc.Batch(dataclasses.asdict(p), np.array([c["id"]], np.int32), np.array([0], np.int32), np.array([c["n_rob"]], np.int32),
                    c["x0"][None].copy(), c["ref"][None, :N].copy(), pA, pb, pr, prev[None].copy(), c["all_pos"].copy(), c["all_valid"].copy(), R)


def solution_in_reference_model(c, traj, ctrl, assign, obj, tol=1e-6):
    """Put a solution (traj (N+1, 9), ctrl (N, 3), assign (N,)) into the model the REFERENCE built for the same inputs (fixture) and return
    (largest violation over bounds, dynamics rows, one-hot rows and the indicator rows of the chosen binaries; objective of the reference's
    model at that point minus the reported objective)."""
    p, N = c["p"], c["p"].n_hor
    nzc = 9 * (N + 1) + 3 * N
    z = np.zeros(len(c["lb"]))
    z[:9 * (N + 1)], z[9 * (N + 1):nzc] = traj.reshape(-1), ctrl.reshape(-1)
    for k in range(N):
        z[nzc + k * p.poly_hor + int(assign[k])] = 1.0
    viol = max(np.max(c["lb"] - z), np.max(z - c["ub"]), 0.0)
    L, Lc, _ = c["lin"]
    viol = max(viol, np.abs(L @ z + Lc).max())
    I, Ic, Ib = c["ind"]
    active = z[Ib] > 0.5
    viol = max(viol, np.max((I @ z + Ic)[active], initial=0.0))
    val = float(z @ (c["obj_diag"] * z) + c["obj_lin"] @ z + c["obj_const"])
    return viol, val - obj


def test_c_port_solutions_are_feasible_and_equally_valued_in_the_reference_model():
    """The checker's solutions, evaluated in the model the reference's own code built for the same inputs: every bound, dynamics row,
    one-hot row and active indicator row holds to 1e-6, and the reference's objective at that point is the objective the solver reports."""
    from oracle import c_oracle as co
    n_opt = 0
    for c in cases():
        out = co.solve_batch(_batch_of(c))
        if out["res"]["status"][0] != 0:
            continue
        viol, dobj = solution_in_reference_model(c, out["traj"][0], out["ctrl"][0], out["assign"][0], float(out["res"]["obj"][0]))
        assert viol <= 1e-6, viol
        assert abs(dobj) <= 1e-7 * max(1.0, abs(out["res"]["obj"][0])), dobj
        n_opt += 1
    assert n_opt >= 2


@pytest.mark.gpu
def test_cuda_solutions_are_feasible_and_equally_valued_in_the_reference_model():
    """Same check for the CUDA path through the C ABI: its solutions satisfy the model the reference itself built, at the reported objective."""
    from multi_agent_pkgs_b200.planner import TrajectoryPlanner
    from oracle import c_oracle as co
    n_opt = 0
    for c in cases():
        b = _batch_of(c)
        pl = TrajectoryPlanner(b.params, max_agents=1, max_neighbours=c["n_rob"], device=0)
        out = pl.solve_batch(b)
        pl.close()
        want = co.solve_batch(b)
        assert out["res"]["status"][0] == want["res"]["status"][0]
        if out["res"]["status"][0] != 0:
            continue
        viol, dobj = solution_in_reference_model(c, out["traj"][0], out["ctrl"][0], out["assign"][0], float(out["res"]["obj"][0]))
        assert viol <= 1e-6, viol
        assert abs(dobj) <= 1e-7 * max(1.0, abs(out["res"]["obj"][0])), dobj
        n_opt += 1
    assert n_opt >= 2
