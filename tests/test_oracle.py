"""CPU tests of the oracle itself: the NumPy restatement against HiGHS, brute force and its own
KKT certificate; the C port against the NumPy restatement; both against the committed golden vectors.
The reference ships no tests for this path, so these are the pins."""
import math

import numpy as np
import pytest

from conftest import load_golden
from oracle import c_oracle as co
from oracle import hdsm_oracle as o
from multi_agent_pkgs_b200 import scenarios as sc


def _planes(p, b, i):
    lo, hi = b.nbr_begin[i], b.nbr_end[i]
    return o.time_aware_planes(p, b.prev_self_pos[i], b.all_pos[lo:hi], b.all_valid[lo:hi], b.global_id[i] - lo)


def test_dynamics_euler_matches_closed_form():
    p = o.Params()
    A, B = o.discrete_dynamics(p)
    dt = p.dt
    Ac = np.zeros((9, 9))
    Ac[0:3, 3:6] = np.eye(3)
    Ac[3:6, 6:9] = np.eye(3)
    assert np.allclose(A, np.eye(9) + dt * Ac, atol=0, rtol=0)
    assert np.allclose(B[6:9], dt * np.eye(3)) and not B[:6].any()
    # u_k first moves p_{k+3}: p_1, p_2 are fixed by x0
    x = o.rollout(p, np.arange(9.0), np.random.default_rng(0).normal(size=(p.n_hor, 3)))
    x2 = o.rollout(p, np.arange(9.0), np.zeros((p.n_hor, 3)))
    assert np.allclose(x[:3, :3], x2[:3, :3]) and not np.allclose(x[3, :3], x2[3, :3])


def test_dynamics_rk4_matches_series():
    p = o.Params(rk4=True, drag=(0.3, 0.1, 0.2))
    A, B = o.discrete_dynamics(p)
    Ac = np.zeros((9, 9))
    Ac[0:3, 3:6] = np.eye(3)
    Ac[3:6, 6:9] = np.eye(3)
    Ac[3:6, 3:6] = -np.diag(p.drag)
    Bc = np.zeros((9, 3))
    Bc[6:9] = np.eye(3)
    M = p.dt * Ac
    Aser = np.eye(9) + M + M @ M / 2 + M @ M @ M / 6 + M @ M @ M @ M / 24
    Bser = p.dt * (np.eye(9) + M / 2 + M @ M / 6 + M @ M @ M / 24) @ Bc
    assert np.allclose(A, Aser, atol=1e-15) and np.allclose(B, Bser, atol=1e-15)


def test_plane_properties_and_c_port():
    """agent_class.cpp:1152-1205: mirrored agents get opposite normals and disjoint half-spaces
    separated by min(2s, |n|) along the line of centres; the C port agrees to rounding."""
    rng = np.random.default_rng(3)
    for params in (sc.agile_params(), sc.crazyflie_params()):
        p = o.Params(**params)
        for _ in range(200):
            a, b = rng.normal(size=3) * 3, rng.normal(size=3) * 3
            n1, b1 = o.interagent_plane(p, a, b)
            n2, b2 = o.interagent_plane(p, b, a)
            assert np.allclose(n1, -n2, atol=1e-12)
            d = b - a
            u = d / np.linalg.norm(d)
            assert abs(n1 @ u - 1) < 1e-12  # tilt terms are orthogonal to the line of centres
            # own point is feasible, neighbour point is not (unless they touch)
            assert n1 @ a <= b1 + 1e-12
            gap = -(b1 + b2)  # n1.x <= b1 and -n1.x <= b2  ->  empty strip of width -(b1+b2) along u
            ang = math.pi / 2 - abs(math.acos(u[2]))
            t = math.atan(p.drone_radius / p.drone_z_offset * math.tan(ang))
            s = math.hypot(p.drone_radius * math.cos(t), p.drone_z_offset * math.sin(t))
            assert abs(gap - min(2 * s, np.linalg.norm(d))) < 1e-9
            nc, bc = co.plane(params, a, b)
            assert np.allclose(nc, n1, rtol=1e-13, atol=1e-13) and abs(bc - b1) <= 1e-12 * max(1, abs(b1))


def test_degenerate_plane_is_flagged():
    n, b = o.interagent_plane(o.Params(), np.ones(3), np.ones(3))
    assert np.isnan(n).all()


@pytest.mark.parametrize("name", ["config1_step1", "config2_step8", "config2_rk4_step5"])
def test_pdip_agrees_with_highs_and_kkt(name):
    b, exp = load_golden(name)
    p = o.Params(**b.params)
    for i in range(min(b.n, 4)):
        if exp["status"][i] != o.OPTIMAL:
            continue
        qp = o.build_qp_full(p, b.x0[i], b.ref[i], b.polys_of(i), _planes(p, b, i), exp["sigma"][i])
        r = o.solve_qp_pdip(qp)
        h = o.solve_qp_highs(qp)
        assert r.status == o.OPTIMAL and h.status == o.OPTIMAL
        assert r.kkt <= 1e-8
        assert abs(r.obj - h.obj) <= 1e-7 * max(1, abs(h.obj))
        assert abs(r.obj - exp["obj"][i]) <= 1e-8 * max(1, abs(r.obj))
        xs, us = qp.split(r.z)
        assert np.allclose(xs, o.rollout(p, b.x0[i], us), atol=1e-9)  # dynamics residual
        assert np.allclose(xs[0], b.x0[i]) and np.abs(xs[-1, 3:]).max() <= 1e-9  # x0 and terminal v = a = 0
        assert np.abs(us).max() <= p.max_jerk + 1e-6 and np.abs(xs[1:-1, 3:6]).max() <= p.max_vel + 1e-6


def test_golden_self_consistency(golden_names):
    """Every committed optimum was re-solved by HiGHS at its final assignment when it was generated
    (HiGHS' active-set QP solver gives up with "solve error" on a few instances: those are recorded
    as inf and are pinned by the KKT certificate alone)."""
    n_ok = n_highs = 0
    for name in golden_names:
        _, exp = load_golden(name)
        ok = exp["status"] == o.OPTIMAL
        assert ok.any()
        hs = ok & np.isfinite(exp["highs_obj"])
        gap = np.abs(exp["obj"][hs] - exp["highs_obj"][hs]) / np.maximum(1, np.abs(exp["obj"][hs]))
        assert gap.max() <= 1e-7 and exp["kkt"][ok].max() <= 1e-8
        n_ok += ok.sum()
        n_highs += hs.sum()
    assert n_highs >= 0.95 * n_ok


def test_miqp_bnb_matches_brute_force():
    """Small instances (N = 5, P = 2 -> 32 assignments): branch and bound == enumeration."""
    sw = sc.config2_circle(seed=7, n_hor=5)
    sw.params["poly_hor"] = 2
    rng = np.random.default_rng(0)
    sw.state[:, 3:6] = rng.normal(size=(sw.n, 3)) * 2
    b = sw.make_batch()
    p = o.Params(**b.params)
    checked = 0
    for i in range(b.n):
        polys = b.polys_of(i)
        if len(polys) < 2:
            continue
        planes = _planes(p, b, i)
        r1 = o.solve_miqp_bnb(p, b.x0[i], b.ref[i], polys, planes)
        r2 = o.solve_miqp_enumerate(p, b.x0[i], b.ref[i], polys, planes)
        assert r1.status == r2.status
        if r1.status == o.OPTIMAL:
            assert abs(r1.obj - r2.obj) <= 1e-7 * max(1, abs(r2.obj))
            checked += 1
        if checked >= 3:
            break
    assert checked >= 1


def test_union_hull_rows_are_valid_for_every_member():
    b, _ = load_golden("config2_step8")
    rng = np.random.default_rng(1)
    for i in range(b.n):
        polys = b.polys_of(i)
        A, d = o.union_hull_rows(polys)
        assert len(d) >= 6  # the six axis faces are always shared
        for (Ap, bp) in polys:
            centre = np.array([np.mean([bp[-6], -bp[-5]]), np.mean([bp[-4], -bp[-3]]), np.mean([bp[-2], -bp[-1]])])
            pts = centre + rng.uniform(-3, 3, size=(200, 3))
            inside = np.all(pts @ Ap.T <= bp, axis=1)
            assert np.all(pts[inside] @ A.T <= d + 1e-12)


@pytest.mark.parametrize("prune", [True, False])
def test_c_port_matches_golden(golden_names, prune):
    for name in golden_names:
        b, exp = load_golden(name)
        out = co.solve_batch(b, max_nodes=5000, prune=prune)
        r = out["res"]
        assert np.array_equal(r["status"], exp["status"]), name
        ok = exp["status"] == o.OPTIMAL
        gap = np.abs(r["obj"][ok] - exp["obj"][ok]) / np.maximum(1, np.abs(exp["obj"][ok]))
        assert gap.max() <= 1e-6, (name, gap.max())
        assert np.abs(out["traj"][ok][..., :3] - exp["traj"][ok][..., :3]).max() <= 1e-3, name  # positions, m
        assert np.abs(out["traj"][ok] - exp["traj"][ok]).max() <= 2e-2, name  # vel / acc are weakly determined
        assert np.abs(out["ctrl"][ok] - exp["ctrl"][ok]).max() <= 0.5, name  # jerk is weakly determined (r_u = 0.01)
        assert r["kkt"][ok].max() <= 1e-6


def test_c_port_fixed_assignment_and_node_limit():
    b, exp = load_golden("config2_step8")
    out = co.solve_batch(b, assign_in=exp["sigma"], max_nodes=1)
    ok = exp["status"] == o.OPTIMAL
    assert (out["res"]["status"][ok] == o.OPTIMAL).all() and (out["res"]["nodes"][ok] == 1).all()
    gap = np.abs(out["res"]["obj"][ok] - exp["obj"][ok]) / np.maximum(1, np.abs(exp["obj"][ok]))
    assert gap.max() <= 1e-6
    hard = int(np.argmax(exp["nodes"]))
    assert exp["nodes"][hard] > 3
    lim = co.solve_batch(b.take([hard]), max_nodes=2)["res"][0]
    assert lim["status"] == 4  # NODE_LIMIT
    assert not np.isfinite(lim["obj"]) or lim["obj"] >= exp["obj"][hard] * (1 - 1e-9)


def test_c_port_infeasible_cases():
    b, exp = load_golden("config2_step8")
    one = b.take([0])
    one.poly_rows = one.poly_rows.copy()
    one.poly_rows[:] = 0  # no polytope: sum of an empty set of binaries == 1 (agent_class.cpp:939-940)
    assert co.solve_batch(one)["res"]["status"][0] == o.INFEASIBLE
    two = b.take([0])
    two.poly_b = two.poly_b - 50.0  # start far outside every cell
    assert co.solve_batch(two)["res"]["status"][0] == o.INFEASIBLE
    three = b.take([0])
    three.x0 = three.x0.copy()
    three.x0[0, 3] = 80.0  # v_1 = v_0 + dt a_0 violates max_vel: constant box row
    assert co.solve_batch(three)["res"]["status"][0] == o.INFEASIBLE


def test_fallback_shift():
    t = np.arange(11 * 9.0).reshape(11, 9)
    c = np.arange(30.0).reshape(10, 3)
    t2, c2 = o.fallback_shift(t, c)
    assert np.array_equal(t2[:-1], t[1:]) and np.array_equal(t2[-1], t[-1]) and np.array_equal(c2[-1], c[-1])


def test_c_port_explicit_inverse_knob_agrees_with_cholesky():
    """ORC_LINSOLVE=inv runs the linear algebra of the CUDA kernel inside the C port (LDL' with the inverse of the unit
    factor accumulated on the side, two products instead of two triangular solves; ORC_LINTHR = the pivot threshold below
    which it falls back to substitution, as the kernel does).  With the kernel's threshold the closed loop must end like
    the Cholesky reference: same statuses and node counts, objectives to 1e-9."""
    import os
    from multi_agent_pkgs_b200 import scenarios as sc
    from oracle import c_oracle as co
    sw = sc.config5_random(seed=11, n_rob=200, side=50.0)
    for step in range(4):
        b = sw.make_batch()
        ref = co.solve_batch(b, max_nodes=64)
        os.environ["ORC_LINSOLVE"], os.environ["ORC_LINTHR"] = "inv", "1e-6"
        try:
            got = co.solve_batch(b, max_nodes=64)
        finally:
            del os.environ["ORC_LINSOLVE"], os.environ["ORC_LINTHR"]
        assert np.array_equal(got["res"]["status"], ref["res"]["status"]), step
        assert np.array_equal(got["res"]["nodes"], ref["res"]["nodes"]), step
        ok = ref["res"]["status"] == 0
        gap = np.abs(got["res"]["obj"][ok] - ref["res"]["obj"][ok]) / np.maximum(1, np.abs(ref["res"]["obj"][ok]))
        assert gap.max() <= 1e-9, (step, gap.max())
        st = ref["res"]["status"]
        sw.advance(ref["traj"], ref["ctrl"], (st == 0) | ((st == 4) & np.isfinite(ref["res"]["obj"])))
