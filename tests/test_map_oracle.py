"""The local-map checkers (no GPU needed): oracle/map_oracle.c against the reference's OWN VoxelGrid
(voxel_grid_util/src/voxel_grid.cpp compiled unmodified into oracle/_ref/libref_voxel.so) where that exists,
and structural properties everywhere.  Bar: byte-exact grids and stencils."""
import numpy as np
import pytest

from multi_agent_pkgs_b200 import mapping as mp, scenarios as sc
from oracle import mapping as om


def _random_grid(rng, t):
    dx, dy, dz = (int(v) for v in rng.integers(6, 40, 3))
    g = np.zeros((dz, dy, dx), np.int8)
    g[rng.random(g.shape) < rng.choice([0.002, 0.02, 0.1])] = 100
    g[rng.random(g.shape) < rng.choice([0.0, 0.05])] = -1
    return g


@pytest.mark.skipif(not om.have_ref(), reason="oracle/_ref not built (reference checkout absent)")
def test_masks_and_passes_match_the_reference():
    for vox, dist, pw in ((0.3, 0.3, 1), (0.3, 1.5, 4), (0.2, 0.5, 1), (0.25, 1.0, 2), (0.3, 0.29, 1), (0.3, 0.0, 1)):
        a, b = om.ref_mask(vox, dist, pw), om.c_mask(vox, dist, pw)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (vox, dist, pw)
    rng = np.random.default_rng(0)
    for t in range(80):
        g = _random_grid(rng, t)
        vox, infl = float(rng.choice([0.3, 0.2])), float(rng.choice([0.3, 0.5, 0.0]))
        pot, pw = float(rng.choice([1.5, 0.9, 0.0])), int(rng.choice([4, 2, 1]))
        assert np.array_equal(om.ref_inflate_potential(g, vox, infl, pot, pw), om.c_inflate_potential(g, vox, infl, pot, pw)), t


def test_default_stencils():
    off, val = om.c_mask(0.3, 0.3, 1)      # inflation 0.3 m at 0.3 m voxels: the 26 neighbours, centre excluded
    assert len(val) == 26 and not (off == 0).all(1).any() and np.abs(off).max() == 1
    off, val = om.c_mask(0.3, 1.5, 4)      # potential field: centre = 100, everything else below
    centre = (off == 0).all(1)
    assert centre.sum() == 1 and val[centre][0] == 100 and val[~centre].max() < 100 and (val >= 0).all()


def test_process_properties():
    sw = sc.config2_circle()
    grids = np.stack([mp.raw_local_grid(sw.world, sw.state[i, :3])[0] for i in range(4)])
    out = om.c_process(grids, 0.3, 0.3, 1.5, 4)
    assert ((grids == 100) <= (out == 100)).all()                    # occupied stays occupied
    assert (out == 100).sum() > 5 * (grids == 100).sum()             # and grows by the inflation stencil
    assert ((grids == -1) <= ((out == -1) | (out == 100))).all()     # unknown stays unknown unless an obstacle inflates into it
    known_free = (out != 100) & (out != -1)
    assert out[known_free].min() >= 0 and out[known_free].max() < 100 and (out[known_free] > 0).any()
    again = om.c_process(out, 0.3, 0.0, 0.0, 4)                      # no inflation, no potential: identity
    assert np.array_equal(again, out)


def _distance_transform_potential(g, tab, rn):
    """The kernel's potential pass in NumPy: squared distance to the nearest occupied voxel by three one-axis passes with
    windows of +-rn (capped at 127 like the kernel's bytes), then the table."""
    occ = g == 100
    Z, Y, X = g.shape
    far = 10000

    def shifted(a, d, axis, fill):
        out = np.full(a.shape, fill, a.dtype)
        n = a.shape[axis]
        if abs(d) >= n:
            return out
        src = [slice(None)] * 3
        dst = [slice(None)] * 3
        src[axis] = slice(0, n - d) if d >= 0 else slice(-d, n)
        dst[axis] = slice(d, n) if d >= 0 else slice(0, n + d)
        out[tuple(dst)] = a[tuple(src)]
        return out

    g1 = np.full(g.shape, far, int)
    for d in range(-rn, rn + 1):
        g1 = np.where(shifted(occ, d, 2, False), np.minimum(g1, d * d), g1)
    g1 = np.minimum(g1, 127)
    for axis in (1, 0):
        nxt = np.full(g.shape, far, int)
        for d in range(-rn, rn + 1):
            nxt = np.minimum(nxt, shifted(g1, d, axis, far) + d * d)
        g1 = np.minimum(nxt, 127)
    val = tab[g1].astype(int)
    out = g.astype(int)
    upd = (g != -1) & (val > out)
    out[upd] = val[upd]
    return out.astype(np.int8)


def test_distance_table_and_the_distance_transform_form_of_the_potential_field():
    """hdsm_map_distance_table (host code of the library, no device needed) against the checker's stencil, and the
    algorithm the kernel runs with it - an exact separable squared-distance transform - against the checker's potential
    pass: byte for byte, grids with pre-existing values and unknown voxels included."""
    import ctypes as C
    from multi_agent_pkgs_b200 import _lib

    class MapParams(C.Structure):
        _fields_ = [("voxel_size", C.c_double), ("inflation_dist", C.c_double), ("potential_dist", C.c_double),
                    ("potential_pow", C.c_int32), ("reserved", C.c_int32)]

    L = _lib.load(build=False)
    L.hdsm_map_distance_table.restype = C.c_int
    rng = np.random.default_rng(5)
    # (0.3, 2.1): radius of 8 voxels, beyond the kernel's windows; (0.3, 0.3): CreateMask leaves the centre out when the
    # distance equals the voxel size, so the mask is not complete from distance 0 on - both keep the stencil walk
    for vox, pot, pw, expect in ((0.3, 1.5, 4, 1), (0.2, 0.9, 2, 1), (0.3, 1.5, 1, 1), (0.25, 1.0, 3, 1), (0.3, 2.1, 2, 0), (0.3, 0.3, 1, 0)):
        tab = np.zeros(128, np.int8)
        prm = MapParams(vox, 0.3, pot, pw, 0)
        ok = L.hdsm_map_distance_table(C.byref(prm), tab.ctypes.data_as(C.c_void_p))
        rn = int(np.ceil(pot / vox))
        assert ok == expect, (vox, pot, pw)
        if not ok:
            continue
        off, val = om.c_mask(vox, pot, pw)
        d2 = (off.astype(int) ** 2).sum(1)
        assert np.array_equal(tab[d2], val)                                   # every stencil entry is its table value
        assert set(np.nonzero(tab != -128)[0]) == set(d2.tolist())             # and nothing else is in the table
        for shape in ((20, 66, 66), (12, 30, 41), (7, 9, 8)):
            g = np.zeros((4,) + shape, np.int8)
            g[rng.random(g.shape) < 0.02] = 100
            g[rng.random(g.shape) < 0.03] = -1
            g[1][:] = 0
            g[2][:] = 100
            g[3][rng.random(shape) < 0.3] = 37
            want = om.c_process(g, vox, 0.0, pot, pw)                        # inflation 0: the potential pass alone
            got = np.stack([_distance_transform_potential(x, tab, rn) for x in g])
            assert np.array_equal(got, want), (vox, pot, pw, shape)
    prm = MapParams(0.3, 0.3, 0.0, 4, 0)
    assert L.hdsm_map_distance_table(C.byref(prm), np.zeros(128, np.int8).ctypes.data_as(C.c_void_p)) == 0   # no stencil
