"""The corridor checkers (no GPU needed).

* oracle/corridor_oracle.c::cor_poly_octa against tests/golden/corridor_ref.npz - fixtures produced by the
  reference's OWN GetPolyOcta3D (convex_decomp_util/src/convex_decomp.cpp:5-376, compiled unmodified into
  oracle/_ref/) - and, where oracle/_ref exists, live against it on fresh random grids.  Bar: bit-exact
  hyperplane points and normals (in the reference's order) and an identical set of marked voxels.
* cor_safe_corridor (Agent::GenerateSafeCorridor, agent_class.cpp:1236-1447): properties of the walk.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from multi_agent_pkgs_b200 import corridor as cr, scenarios as sc
from oracle import corridor as oc


def golden_cases():
    z = np.load(os.path.join(GOLDEN, "corridor_ref.npz"))
    for k in range(int(z["n_cases"])):
        x, y, zz, n_it, conv, use_new = (int(v) for v in z[f"call{k}"])
        fp = z[f"fp{k}"]
        yield dict(grid=z[f"grid{k}"], seed=(x, y, zz), n_it=n_it, conv=conv, res=float(fp[0]), origin=fp[1:4], use_new=bool(use_new),
                   pts=z[f"pts{k}"], nrm=z[f"nrm{k}"], marks=z[f"marks{k}"])


def test_c_restatement_matches_reference_golden():
    n = 0
    for c in golden_cases():
        pts, nrm, marked = oc.c_poly(c["grid"], c["seed"], c["n_it"], c["res"], c["conv"], c["origin"], use_new=c["use_new"])
        assert pts.shape == c["pts"].shape
        assert np.array_equal(pts, c["pts"]) and np.array_equal(nrm, c["nrm"])  # bit-exact, same order
        assert np.array_equal(np.flatnonzero(marked.ravel() == c["conv"]), c["marks"])
        n += 1
    assert n >= 112


def test_golden_planes_bound_the_marked_voxels():
    """The six face planes (the last six rows) bound every marked voxel and the seed voxel's centre.  (Chamfer rows may cut marked voxels in the reference: a chamfer is fitted to the
    staircase of layer ends, convex_decomp.cpp:217-301, not to the voxel corners.)"""
    for c in golden_cases():
        g = c["grid"]
        b = (c["pts"] * c["nrm"]).sum(1)
        idx = c["marks"]
        dz, dy, dx = g.shape
        zc, yc, xc = idx // (dx * dy), (idx // dx) % dy, idx % dx
        ctr = (np.stack([xc, yc, zc], 1) + 0.5) * c["res"] + c["origin"]
        assert (ctr @ c["nrm"][-6:].T - b[None, -6:]).max() <= -0.5 * c["res"] + 1e-9
        seed = (np.array(c["seed"]) + 0.5) * c["res"] + c["origin"]
        assert (c["nrm"][-6:] @ seed - b[-6:]).max() <= 1e-9
        assert 6 <= len(b) <= 18


@pytest.mark.skipif(not oc.have_ref(), reason="oracle/_ref not built (reference checkout absent)")
def test_c_restatement_matches_reference_live():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_corridor_golden", os.path.join(GOLDEN, "make_corridor_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    random_grid = mod.random_grid
    rng = np.random.default_rng(77)
    done = 0
    while done < 400:
        dx, dy, dz = (int(v) for v in (rng.integers(12, 70), rng.integers(12, 70), rng.integers(8, 24)))
        g = random_grid(rng, done % 4, dx, dy, dz)
        free = np.argwhere(g < 100)
        if len(free) == 0:
            continue
        z, y, x = (int(v) for v in free[rng.integers(len(free))])
        n_it, res, conv = int(rng.choice([6, 17, 42, 60, 90])), float(rng.choice([0.3, 0.2])), -int(rng.integers(1, 5))
        origin = rng.uniform(-20, 20, 3)
        use_new = done % 2 == 1  # GetPolyOcta3D and GetPolyOcta3DNew alternate; every fourth seed is squeezed
        if done % 4 == 3:
            ax = int(rng.integers(3))
            for sgn in (-1, 1):
                c = [x, y, z]
                c[ax] += sgn
                if 0 <= c[0] < dx and 0 <= c[1] < dy and 0 <= c[2] < dz:
                    g[c[2], c[1], c[0]] = 100
        a = oc.ref_poly(g, (x, y, z), n_it, res, conv, origin, use_new=use_new)
        b = oc.c_poly(g, (x, y, z), n_it, res, conv, origin, use_new=use_new)
        assert all(np.array_equal(u, v) for u, v in zip(a, b)), (done, use_new, (dx, dy, dz), (x, y, z), n_it)
        done += 1


@pytest.fixture(scope="module")
def forest_batch():
    sw = sc.config2_circle(n_swarms=3)
    for i in range(sw.n):  # pull the agents into the forest so that the corridors have chamfers
        sw.state[i, :2] = sw.world.push_free(0.45 * sw.state[i, :2] + 0.55 * sw.goal[i, :2], 0.3)
    return sw, cr.corridor_batch(sw)


def test_safe_corridor_properties(forest_batch):
    sw, cb = forest_batch
    out = oc.c_safe_corridor(cb)
    assert (out["flags"] & ~oc.FLAG_SQUEEZED == 0).all()
    assert (out["poly_rows"][:, 0] >= 6).all() and out["poly_rows"].max() <= 18
    for i in range(cb.n):
        for p in range(cb.poly_hor):
            r = out["poly_rows"][i, p]
            if r == 0:
                assert (out["poly_rows"][i, p:] == 0).all()  # leading slots are filled first (agent_class.cpp:913)
                break
            A, b, s = out["poly_A"][i, p, :r], out["poly_b"][i, p, :r], out["seeds"][i, p]
            assert (A[-6:] @ s - b[-6:]).max() < 0             # the seed voxel centre is inside the six faces ...
            # ... while a chamfer may clip it by a fraction of a voxel (reference behaviour, pinned by the golden test)
            assert ((A @ s - b) / np.linalg.norm(A, axis=1)).max() <= cb.voxel
            assert np.array_equal(A[-6:], np.array([[0, -1, 0], [1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1.0]]))
            assert (np.abs(A) == np.round(np.abs(A))).all()    # small-integer normals (convex_decomp.cpp:341-357)
            # seeds sit on voxel centres of the agent's grid
            k = (s - cb.origins[i]) / cb.voxel - 0.5
            assert np.abs(k - np.round(k)).max() < 1e-9
        # the first polytope contains the agent (seed = the voxel of state_curr_)
        r = out["poly_rows"][i, 0]
        assert (out["poly_A"][i, 0, :r] @ cb.pos[i] - out["poly_b"][i, 0, :r]).max() <= 0.3


@pytest.mark.skipif(not oc.have_ref(), reason="oracle/_ref not built (reference checkout absent)")
def test_safe_corridor_polytopes_match_reference_per_seed(forest_batch):
    """Every polytope of the walk equals the reference's GetPolyOcta3D run on a fresh OccupyUnknown'ed copy
    of the agent's grid at the same seed voxel (agent_class.cpp:1405-1417).  Catches state leaking from
    one polytope to the next (an occupied seed voxel is marked, i.e. freed, in the working copy)."""
    sw, cb = forest_batch
    out = oc.c_safe_corridor(cb)
    checked = n_new = 0
    for i in range(cb.n):
        g = cb.grids[i].copy()
        g[g == cr.UNKNOWN] = cr.OCC
        for p in range(cb.poly_hor):
            r = out["poly_rows"][i, p]
            if r == 0:
                continue
            sv = np.round((out["seeds"][i, p] - cb.origins[i]) / cb.voxel - 0.5).astype(int)
            squeezed = False  # agent_class.cpp:1385-1395 (IsOccupied is false outside the grid)
            for ax in range(3):
                lo, hi = sv.copy(), sv.copy()
                lo[ax] -= 1
                hi[ax] += 1
                if lo[ax] >= 0 and hi[ax] < g.shape[2 - ax]:
                    squeezed |= g[lo[2], lo[1], lo[0]] == cr.OCC and g[hi[2], hi[1], hi[0]] == cr.OCC
            n_new += squeezed
            assert bool(out["flags"][i] & oc.FLAG_SQUEEZED) >= squeezed
            pts, nrm, _ = oc.ref_poly(g, sv, cb.n_it, cb.voxel, -(p + 1), cb.origins[i], use_new=squeezed)
            b = (pts[:, 0] * nrm[:, 0] + pts[:, 1] * nrm[:, 1]) + pts[:, 2] * nrm[:, 2]
            assert len(b) == r and np.array_equal(nrm, out["poly_A"][i, p, :r]) and np.array_equal(b, out["poly_b"][i, p, :r])
            checked += 1
    assert checked >= 3 * cb.n and n_new >= 1


def test_safe_corridor_keeps_previous_polytopes(forest_batch):
    sw, cb = forest_batch
    first = oc.c_safe_corridor(cb)
    n, P = cb.n, cb.poly_hor
    N = sw.params["n_hor"]
    used = np.zeros((n, P), np.uint8)
    used[:, 1] = 1                                     # the optimisation used polytope 1 only
    traj = np.repeat(cb.pos[:, None, :], N + 1, 1)     # hovering plan: inside the last polytope only by luck
    cb2 = cr.corridor_batch(sw).with_previous(first, used, traj)
    second = oc.c_safe_corridor(cb2)
    for i in range(n):
        nprev = int((first["poly_rows"][i] > 0).sum())
        last = nprev - 1
        r = first["poly_rows"][i, last]
        inside_last = (first["poly_A"][i, last, :r] @ cb.pos[i] - first["poly_b"][i, last, :r]).max() <= 0
        keep = last if inside_last else (1 if nprev > 1 else None)
        if keep is None:
            continue
        assert np.array_equal(second["poly_A"][i, 0], first["poly_A"][i, keep])
        assert np.array_equal(second["poly_b"][i, 0], first["poly_b"][i, keep])
        assert np.array_equal(second["seeds"][i, 0], first["seeds"][i, keep])
        assert len({tuple(s) for s, r in zip(second["seeds"][i], second["poly_rows"][i]) if r > 0}) == \
            int((second["poly_rows"][i] > 0).sum())    # a seed is never used twice (:1355-1372)


def test_local_grid_shape_and_floor():
    sw = sc.config2_circle()
    g, o = cr.local_grid(sw.world, sw.state[0, :3])
    assert g.shape == (20, 66, 66) and g.dtype == np.int8
    assert np.abs(o / 0.3 - np.round(o / 0.3)).max() < 1e-9
    assert (g[:5] == cr.UNKNOWN).all() and (g[6:] != cr.UNKNOWN).all()  # z origin = 1.5 - 3.0: five layers below ground
