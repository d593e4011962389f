"""Host-side logic of the closed-loop harness and the search devices added to the checker (CPU only):
input producers spread over a process pool, the forest's spatial index, per-step cell dominance."""
import numpy as np

from multi_agent_pkgs_b200 import scenarios as sc
from oracle import c_oracle as co, hdsm_oracle as o


def test_pooled_inputs_equal_serial_inputs_and_shards_equal_slices():
    sw = sc.config5_random(seed=4, n_rob=96, side=40.0)
    whole = sw.make_batch()
    pool = sc.InputPool(sw, 3)
    try:
        ref, A, b, rows = sw.make_inputs(np.arange(96), pool=pool)
        pooled = sw.make_batch_pooled(pool)
    finally:
        pool.close()
    for got, want in ((ref, whole.ref), (A, whole.poly_A), (b, whole.poly_b), (rows, whole.poly_rows)):
        assert np.array_equal(got, want)
    for k in ("global_id", "nbr_begin", "nbr_end", "x0", "ref", "poly_A", "poly_b", "poly_rows", "prev_self_pos", "all_pos", "all_valid"):
        assert np.array_equal(getattr(pooled, k), getattr(whole, k)), k
    ids = np.arange(40, 70)
    part = sw.make_inputs(ids, pos=sw.state[ids, :3], step=sw.step_count)
    assert np.array_equal(part[0], whole.ref[40:70]) and np.array_equal(part[1], whole.poly_A[40:70])


def test_forest_index_returns_what_the_plain_scan_returns():
    rng = np.random.default_rng(0)
    f = sc.Forest.density(rng, (0.0, 0.0), (80.0, 80.0), 0.2)
    plain = sc.Forest(f.cols[:500])  # below the index threshold: plain scan
    assert len(f.cols) >= 512
    for _ in range(300):
        xy = rng.uniform(-5, 85, 2)
        rad = rng.uniform(0.3, 9.0)
        d = np.abs(f.cols - xy[None])
        want = f.cols[(d[:, 0] < rad) & (d[:, 1] < rad)]
        assert np.array_equal(f.near(xy, rad), want)
        assert f.is_free(xy, 0.2) == (not np.any((d[:, 0] < sc.KEEP_OUT + 0.2) & (d[:, 1] < sc.KEEP_OUT + 0.2)))
    assert np.array_equal(plain.near([40.0, 40.0], 5.0), plain.cols[(np.abs(plain.cols - [40.0, 40.0]) < 5.0).all(1)])


def _planes(p, b, i):
    lo, hi = b.nbr_begin[i], b.nbr_end[i]
    return o.time_aware_planes(p, b.prev_self_pos[i], b.all_pos[lo:hi], b.all_valid[lo:hi], b.global_id[i] - lo)


def test_cell_dominance_keeps_the_optimum():
    """The checker drops, per step, cells that another candidate covers wherever the segment can be.  Against the NumPy
    branch and bound (no dominance, un-condensed QPs) on agents with several cells - including exact duplicates and
    duplicates that differ by rows out of reach, the case the filter exists for - the optimum must be the same."""
    sw = sc.config5_random(seed=8, n_rob=48, side=30.0)
    for _ in range(2):
        b = sw.make_batch()
        r = co.solve_batch(b)
        sw.advance(r["traj"], r["ctrl"], r["res"]["status"] == 0)
    b = sw.make_batch()
    p = o.Params(**b.params)
    multi = [i for i in range(b.n) if (b.poly_rows[i] > 0).sum() >= 3][:6]
    assert len(multi) >= 3
    # agent multi[0]: cell 2 := cell 1 (exact duplicate); agent multi[1]: cell 2 := cell 1 plus a chamfer far out of reach
    i0, i1 = multi[0], multi[1]
    for i in (i0, i1):
        r1 = b.poly_rows[i, 1]
        b.poly_A[i, 2], b.poly_b[i, 2], b.poly_rows[i, 2] = b.poly_A[i, 1].copy(), b.poly_b[i, 1].copy(), r1
    r1 = b.poly_rows[i1, 1]
    assert r1 < b.rmax
    # faces are the last six rows (+x, -x, +y, -y, +z, -z): an x-z chamfer that only cuts the far top corner
    b.poly_A[i1, 2, r1] = (1.0, 0.0, 1.0)
    b.poly_b[i1, 2, r1] = b.poly_b[i1, 2, r1 - 6] + b.poly_b[i1, 2, r1 - 2] - 0.05
    b.poly_rows[i1, 2] = r1 + 1
    got = co.solve_batch(b.take(multi), max_nodes=5000)
    for r, i in enumerate(multi):
        want = o.solve_miqp_bnb(p, b.x0[i], b.ref[i], b.polys_of(i), _planes(p, b, i))
        assert got["res"]["status"][r] == want.status, i
        if want.status == o.OPTIMAL:
            assert abs(got["res"]["obj"][r] - want.obj) <= 1e-6 * max(1.0, abs(want.obj)), i
    # and the filter does what it is for: duplicates no longer multiply the search
    dup = co.solve_batch(b.take([i0]), max_nodes=5000)["res"][0]
    b.poly_rows[i0, 2] = 0
    b.poly_rows[i0, 3] = 0
    nodup = co.solve_batch(b.take([i0]), max_nodes=5000)["res"][0]
    assert dup["status"] == nodup["status"]
