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
