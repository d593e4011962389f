"""Parity of the CUDA local-map post-processing (hdsm_map_batch, csrc/hdsm_map.cu) through the C ABI against
oracle/map_oracle.c, whose inflation / potential passes and stencils are pinned to the reference's own VoxelGrid.
Bar: byte-exact grids."""
import numpy as np
import pytest

from multi_agent_pkgs_b200 import corridor as cr, mapping as mp, scenarios as sc
from oracle import mapping as om

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("vox,infl,pot,pw", [(0.3, 0.3, 1.5, 4), (0.2, 0.5, 0.9, 2), (0.3, 0.0, 1.5, 1), (0.3, 0.3, 0.0, 4)])
def test_random_grids_match_the_checker(vox, infl, pot, pw):
    rng = np.random.default_rng(11)
    for shape in ((20, 66, 66), (12, 30, 41), (7, 9, 8)):
        g = np.zeros((6,) + shape, np.int8)
        g[rng.random(g.shape) < 0.02] = 100
        g[rng.random(g.shape) < 0.03] = -1
        g[0, :, :, :3] = -1
        g[1][:] = 0                                   # an empty grid
        g[2][:] = 100                                 # a full one
        gen = mp.MapProcessor(vox, 6, g[0].size, infl, pot, pw)
        out = gen.process(g)
        assert np.array_equal(out, om.c_process(g, vox, infl, pot, pw)), shape
        assert gen.launch_count == 1
        gen.close()


def test_forest_grids_feed_corridor_and_reference():
    """Raw forest grids -> hdsm_map_batch -> the grids the corridor generator takes; equal to the checker's, and
    the agent's own voxel stays usable (the corridor grows polytopes in them)."""
    from oracle import corridor as oc
    sw = sc.config2_circle(n_swarms=2)
    raw = np.stack([mp.raw_local_grid(sw.world, sw.state[i, :3])[0] for i in range(sw.n)])
    gen = mp.MapProcessor(0.3, sw.n, raw[0].size)
    out = gen.process(raw)
    gen.close()
    assert np.array_equal(out, om.c_process(raw, 0.3, 0.3, 1.5, 4))
    cb = cr.corridor_batch(sw)
    cb.grids = out
    cg = cr.SafeCorridorGenerator(cb.poly_hor, cb.n_it, cb.voxel, cb.n, cb.n, int(out[0].size), cb.prev_traj.shape[1], cb.path.shape[1])
    got = cg.generate(cb)
    cg.close()
    want = oc.c_safe_corridor(cb)
    for k in ("poly_rows", "poly_A", "poly_b", "seeds", "flags"):
        assert np.array_equal(got[k], want[k]), k
    assert (got["poly_rows"][:, 0] >= 6).all()


def test_map_error_codes():
    with pytest.raises(RuntimeError):
        mp.MapProcessor(0.3, 4, 200_000)             # two copies do not fit into shared memory
    gen = mp.MapProcessor(0.3, 2, 1000)
    with pytest.raises(RuntimeError, match="max_grids"):
        gen.process(np.zeros((3, 10, 10, 10), np.int8))
    gen.close()
