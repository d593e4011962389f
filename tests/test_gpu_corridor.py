"""Parity of the CUDA corridor generator (hdsm_corridor_batch, csrc/hdsm_corridor.cu) through the C ABI.

Bar: BIT-EXACT.  Normals are small integers; points, offsets b = p . n and seeds are doubles computed
in the reference's evaluation order without FMA contraction, so every output array must equal the
checker's array exactly:
  * against tests/golden/corridor_ref.npz - GetPolyOcta3D outputs of the reference's own code;
  * against oracle/corridor_oracle.c (cor_safe_corridor) on scenario batches, first and follow-up steps.
"""
import numpy as np
import pytest

from multi_agent_pkgs_b200 import corridor as cr, scenarios as sc
from oracle import corridor as oc
from test_corridor_oracle import golden_cases

pytestmark = pytest.mark.gpu


def _gen(cb):
    G = cb.grids.shape[0]
    return cr.SafeCorridorGenerator(cb.poly_hor, cb.n_it, cb.voxel, cb.n, G, int(cb.grids[0].size), cb.prev_traj.shape[1],
                                    cb.path.shape[1], rmax=cb.rmax, use_cvx_new=cb.use_cvx_new)


def _assert_equal(out, ref):
    for k in ("poly_rows", "flags", "poly_A", "poly_b", "seeds"):
        assert np.array_equal(out[k], ref[k]), k


def test_single_polytopes_match_reference_golden():
    """One GetPolyOcta3D / GetPolyOcta3DNew per agent, seeded at the golden seed voxel (poly_hor = 1, a 0.7 m
    path).  The New cases run with use_cvx_new = 1, like the reference with that parameter set."""
    by_cfg = {}
    for c in golden_cases():
        by_cfg.setdefault((c["n_it"], c["res"], c["use_new"]), []).append(c)
    checked = n_new = 0
    for (n_it, res, use_new), cases in by_cfg.items():
        stride = max(c["grid"].size for c in cases)
        n = len(cases)
        grids = np.zeros((n, stride), np.int8)
        dims, origins, pos = np.zeros((n, 3), np.int32), np.zeros((n, 3)), np.zeros((n, 3))
        path, n_path = np.zeros((n, 2, 3)), np.ones(n, np.int32)
        for i, c in enumerate(cases):
            g = c["grid"]
            grids[i, :g.size] = g.ravel()
            dims[i] = (g.shape[2], g.shape[1], g.shape[0])
            origins[i] = c["origin"]
            pos[i] = (np.array(c["seed"]) + 0.5) * res + c["origin"]
            path[i, 0] = pos[i] + np.array([0.7, 0.0, 0.0])
        gen = cr.SafeCorridorGenerator(1, n_it, res, n, n, stride, 0, 2, rmax=18, use_cvx_new=use_new)
        cb = cr.CorridorBatch(1, n_it, 18, res, grids[:, None, None, :], None, dims, origins, pos, path, n_path,
                              np.zeros((n, 0, 3)))
        out = gen.generate(cb)
        gen.close()
        for i, c in enumerate(cases):
            r = len(c["pts"])
            assert out["poly_rows"][i, 0] == r, (i, out["poly_rows"][i], r, out["flags"][i])
            assert np.array_equal(out["poly_A"][i, 0, :r], c["nrm"])
            b = (c["pts"][:, 0] * c["nrm"][:, 0] + c["pts"][:, 1] * c["nrm"][:, 1]) + c["pts"][:, 2] * c["nrm"][:, 2]
            assert np.array_equal(out["poly_b"][i, 0, :r], b)
            checked += 1
            n_new += use_new
    assert checked >= 112 and n_new >= 48


@pytest.mark.parametrize("n_it", [42, 60])
def test_scenario_batch_matches_c_restatement(n_it):
    sw = sc.config2_circle(n_swarms=6)
    for i in range(sw.n):
        sw.state[i, :2] = sw.world.push_free(0.45 * sw.state[i, :2] + 0.55 * sw.goal[i, :2], 0.3)
    cb = cr.corridor_batch(sw, n_it=n_it)
    gen = _gen(cb)
    out = gen.generate(cb)
    ref = oc.c_safe_corridor(cb)
    _assert_equal(out, ref)
    assert gen.launch_count == 1
    # follow-up step: previous polytopes, used flags and previous plan decide what is kept
    rng = np.random.default_rng(5)
    used = (rng.random((cb.n, cb.poly_hor)) < 0.5).astype(np.uint8)
    N = sw.params["n_hor"]
    traj = cb.pos[:, None, :] + np.linspace(0, 1, N + 1)[None, :, None] * (cb.path[:, 0] - cb.pos)[:, None, :] * \
        rng.uniform(0.0, 0.4, (cb.n, 1, 1))
    sw.state[:, :3] = traj[:, 1]
    cb2 = cr.corridor_batch(sw, n_it=n_it).with_previous(out, used, traj)
    out2 = gen.generate(cb2)
    ref2 = oc.c_safe_corridor(cb2)
    _assert_equal(out2, ref2)
    assert (out2["poly_rows"] > 0).any()
    gen.close()


def test_squeezed_seeds_take_the_new_method():
    """Agents pulled next to the columns: squeezed seeds (flag bit 0) are grown with GetPolyOcta3DNew by the
    kernel and by the checker alike; also the all-New mode (use_cvx_new = 1)."""
    sw = sc.config2_circle(n_swarms=24)
    for i in range(sw.n):
        sw.state[i, :2] = sw.world.push_free(0.45 * sw.state[i, :2] + 0.55 * sw.goal[i, :2], 0.3)
    for use_new in (False, True):
        cb = cr.corridor_batch(sw)
        cb.use_cvx_new = use_new
        gen = _gen(cb)
        out = gen.generate(cb)
        gen.close()
        _assert_equal(out, oc.c_safe_corridor(cb))
        assert (out["flags"] & cr.FLAG_SQUEEZED).any()
        assert (out["flags"] & ~cr.FLAG_SQUEEZED == 0).all()


def test_shared_grid_and_empty_map():
    """Agents sharing one grid through grid_index; an empty map gives the 4.5 m cube."""
    sw = sc.config1_single_agent()
    cb = cr.corridor_batch(sw)
    n = 5
    cb5 = cr.CorridorBatch(cb.poly_hor, cb.n_it, cb.rmax, cb.voxel, cb.grids, np.zeros(n, np.int32), np.repeat(cb.dims, n, 0),
                           np.repeat(cb.origins, n, 0), np.repeat(cb.pos, n, 0) + np.arange(n)[:, None] * 0.31,
                           np.repeat(cb.path, n, 0), np.repeat(cb.n_path, n), np.zeros((n, 11, 3)))
    gen = cr.SafeCorridorGenerator(cb.poly_hor, cb.n_it, cb.voxel, n, 1, int(cb.grids[0].size), 11, cb.path.shape[1])
    out = gen.generate(cb5)
    cb5_own = cr.CorridorBatch(cb.poly_hor, cb.n_it, cb.rmax, cb.voxel, np.repeat(cb.grids, n, 0), None, cb5.dims, cb5.origins,
                               cb5.pos, cb5.path, cb5.n_path, cb5.prev_traj)
    _assert_equal(out, oc.c_safe_corridor(cb5_own))
    r = out["poly_rows"][0, 0]
    assert r == 6
    A, b = out["poly_A"][0, 0, :6], out["poly_b"][0, 0, :6]
    ext = b[1] + b[3]  # +x and -x faces: 15 voxels of 0.3 m
    assert abs(ext - 4.5) < 1e-9
    gen.close()


def test_corridor_feeds_the_optimisation():
    """Corridor rows go straight into hdsm_solve_batch: the agents get OPTIMAL plans inside them."""
    from multi_agent_pkgs_b200.planner import TrajectoryPlanner
    sw = sc.config2_circle(n_swarms=2)
    cb = cr.corridor_batch(sw)
    gen = _gen(cb)
    out = gen.generate(cb)
    gen.close()
    b = sw.make_batch()
    b.poly_A, b.poly_b, b.poly_rows = out["poly_A"], out["poly_b"], out["poly_rows"]
    pl = TrajectoryPlanner(sw.params, max_agents=b.n, max_neighbours=10)
    res = pl.solve_batch(b)
    pl.close()
    from oracle import c_oracle as co
    ref = co.solve_batch(b)
    assert np.array_equal(res["res"]["status"], ref["res"]["status"])
    ok = ref["res"]["status"] == 0
    assert ok.mean() > 0.8
    gap = np.abs(res["res"]["obj"][ok] - ref["res"]["obj"][ok]) / np.maximum(1.0, np.abs(ref["res"]["obj"][ok]))
    assert gap.max() <= 1e-6


def test_closed_loop_corridor_then_optimisation():
    """Ten replanning steps of 20 agents through both replaced calls on the GPU; every step the corridor rows
    must equal the checker's bit for bit (kept polytopes, used flags and previous plans now come from real
    solves) and the optimisation must agree with the C port on those rows."""
    from multi_agent_pkgs_b200.planner import TrajectoryPlanner
    from oracle import c_oracle as co
    sw = sc.config2_circle(n_swarms=2, seed=9)
    for i in range(sw.n):
        sw.state[i, :2] = sw.world.push_free(0.6 * sw.state[i, :2] + 0.4 * sw.goal[i, :2], 0.3)
    loop = cr.CorridorLoop(sw)
    cb = loop.corridor_inputs()
    gen = _gen(cb)
    pl = TrajectoryPlanner(sw.params, max_agents=sw.n, max_neighbours=10, max_nodes=5000)  # exact search on both sides
    kept = solved_ok = 0
    for step in range(10):
        cb = loop.corridor_inputs()
        out = gen.generate(cb)
        _assert_equal(out, oc.c_safe_corridor(cb))
        assert (out["flags"] & ~cr.FLAG_SQUEEZED == 0).all() and (out["poly_rows"][:, 0] >= 6).all()
        if cb.prev_n is not None:  # kept polytopes come first, unchanged
            for i in range(cb.n):
                keep = [p for p in range(cb.prev_n[i]) if cb.prev_used[i, p]]
                last = cb.prev_n[i] - 1
                r = cb.prev_rows[i, last]
                inside = (cb.prev_A[i, last, :r] @ cb.prev_traj[i].T - cb.prev_b[i, last, :r, None]).max() <= 0
                src = [last] if inside else keep
                for slot, p in enumerate(src):
                    assert np.array_equal(out["poly_A"][i, slot], cb.prev_A[i, p]) and np.array_equal(out["seeds"][i, slot], cb.prev_seeds[i, p])
                    kept += 1
        b = loop.solver_inputs(out)
        res = pl.solve_batch(b)
        ref = co.solve_batch(b, max_nodes=5000)
        assert np.array_equal(res["res"]["status"], ref["res"]["status"])
        ok = ref["res"]["status"] == 0
        if ok.any():
            gap = np.abs(res["res"]["obj"][ok] - ref["res"]["obj"][ok]) / np.maximum(1.0, np.abs(ref["res"]["obj"][ok]))
            assert gap.max() <= 1e-6
        solved_ok += int(ok.sum())
        loop.advance(out, res, ok)
    assert kept > 50 and solved_ok > 100
    gen.close()
    pl.close()


def _free_batch(n, dims=(40, 40, 16), voxel=0.3, n_it=42, poly_hor=3, max_path=4, n_traj=0):
    g = np.zeros((n, dims[2], dims[1], dims[0]), np.int8)
    d = np.tile(np.array(dims, np.int32), (n, 1))
    o = np.zeros((n, 3))
    pos = np.tile((np.array(dims) / 2 + 0.5) * voxel, (n, 1))
    path = np.zeros((n, max_path, 3))
    path[:, 0] = pos + np.array([1.0, 0.2, 0.0])
    return cr.CorridorBatch(poly_hor, n_it, 18, voxel, g, None, d, o, pos, path, np.ones(n, np.int32), np.zeros((n, n_traj, 3)))


def test_edge_cases_match_the_checker():
    """Empty path, a path that leaves the grid, the largest supported n_it_decomp on a big free grid, another
    voxel size, previous polytopes of which none survives, obstacles right at the grid boundary."""
    # (a) empty path: nothing is grown, outputs are zero, no flag
    cb = _free_batch(3)
    cb.n_path[:] = 0
    gen = _gen(cb)
    out = gen.generate(cb)
    _assert_equal(out, oc.c_safe_corridor(cb))
    assert (out["poly_rows"] == 0).all() and (out["flags"] == 0).all()
    # (b) the path leaves the grid: polytopes up to there, then HDSM_COR_SEED_OUTSIDE
    cb = _free_batch(3, poly_hor=8)
    cb.path[:, 0] = cb.pos + np.array([30.0, 0.0, 0.0])
    gen.close()
    gen = _gen(cb)
    out = gen.generate(cb)
    _assert_equal(out, oc.c_safe_corridor(cb))
    assert (out["flags"] & cr.FLAG_SEED_OUTSIDE).all() and (out["poly_rows"][:, 0] == 6).all() and (out["poly_rows"][:, 7] == 0).all()
    gen.close()
    # (c) n_it_decomp = 90 on a free 70^3 grid: 15 layers per face, the whole 32^3 window in use
    cb = _free_batch(2, dims=(70, 70, 70), n_it=90, poly_hor=2)
    gen = _gen(cb)
    out = gen.generate(cb)
    _assert_equal(out, oc.c_safe_corridor(cb))
    b = out["poly_b"][0, 0, :6]
    assert abs((b[1] + b[3]) - 31 * 0.3) < 1e-9  # 15 + 1 + 15 voxels wide
    gen.close()
    # (d) voxel 0.2, n_it 60, columns touching the grid boundary
    cb = _free_batch(4, dims=(36, 30, 12), voxel=0.2, n_it=60, poly_hor=3)
    cb.grids[:, :, :, 0] = 100
    cb.grids[:, :, -1, :] = 100
    cb.grids[:, :, 10:14, 22:24] = 100
    cb.grids[:, :3] = -1
    gen = _gen(cb)
    out = gen.generate(cb)
    _assert_equal(out, oc.c_safe_corridor(cb))
    assert (out["poly_rows"][:, 0] >= 6).all()
    # (e) previous polytopes, none used and the plan outside the last one: all are dropped and regrown
    N1 = 5
    cb2 = _free_batch(4, dims=(36, 30, 12), voxel=0.2, n_it=60, poly_hor=3, n_traj=N1)
    cb2.grids = cb.grids
    far = np.repeat(cb2.pos[:, None, :] + np.array([100.0, 0.0, 0.0]), N1, 1)
    cb2.with_previous(out, np.zeros((4, 3), np.uint8), far)
    gen.close()
    gen = _gen(cb2)
    out2 = gen.generate(cb2)
    _assert_equal(out2, oc.c_safe_corridor(cb2))
    for k in ("poly_A", "poly_b", "poly_rows", "seeds"):
        assert np.array_equal(out2[k], out[k]), k
    gen.close()


def test_corridor_error_codes():
    cb = _free_batch(4)
    gen = cr.SafeCorridorGenerator(cb.poly_hor, cb.n_it, cb.voxel, 2, 2, int(cb.grids[0].size), 0, cb.path.shape[1])
    with pytest.raises(RuntimeError, match="capacity"):
        gen.generate(cb)  # 4 agents on a handle made for 2
    small = _free_batch(2)
    small.grid_index = np.array([0, 5], np.int32)
    with pytest.raises(RuntimeError, match="grid_index"):
        gen.generate(small)
    small.grid_index = None
    small.dims[1] = (400, 400, 400)
    with pytest.raises(RuntimeError, match="grid_stride"):
        gen.generate(small)
    gen.close()


@pytest.mark.parametrize("shared", [False, True])
def test_chunked_host_pipeline_matches_the_checker(shared):
    """Batches of 2048 agents and more go through hdsm_corridor_batch in four chunks over two streams (pinned
    staging, H2D, kernel and D2H overlapped): 2 520 agents, per-agent grids and grids shared through grid_index,
    with previous polytopes - every output equal to the checker's."""
    sw = sc.config2_circle(n_swarms=2, seed=13)
    for i in range(sw.n):
        sw.state[i, :2] = sw.world.push_free(0.5 * sw.state[i, :2] + 0.5 * sw.goal[i, :2], 0.3)
    base = cr.corridor_batch(sw)
    first = oc.c_safe_corridor(base)
    rng = np.random.default_rng(3)
    N = sw.params["n_hor"]
    base.with_previous(first, (rng.random((base.n, base.poly_hor)) < 0.6).astype(np.uint8),
                       np.repeat(base.pos[:, None, :], N + 1, 1) + rng.normal(0, 0.3, (base.n, N + 1, 3)))
    reps = 126

    def tile(a):
        return np.ascontiguousarray(np.concatenate([a] * reps))
    grids = base.grids if shared else tile(base.grids)
    gi = np.tile(np.arange(base.n, dtype=np.int32), reps) if shared else None
    cb = cr.CorridorBatch(base.poly_hor, base.n_it, base.rmax, base.voxel, grids, gi, tile(base.dims), tile(base.origins),
                          tile(base.pos), tile(base.path), tile(base.n_path), tile(base.prev_traj), tile(base.prev_n),
                          tile(base.prev_A), tile(base.prev_b), tile(base.prev_rows), tile(base.prev_seeds), tile(base.prev_used))
    gen = _gen(cb)
    for _ in range(2):  # twice: the arenas are reused
        out = gen.generate(cb)
        small = oc.c_safe_corridor(base)
        for k in ("poly_rows", "flags", "poly_A", "poly_b", "seeds"):
            got = out[k].reshape(reps, base.n, *out[k].shape[1:])
            assert np.array_equal(got, np.broadcast_to(small[k], got.shape)), k
    assert gen.launch_count == 8
    gen.close()
