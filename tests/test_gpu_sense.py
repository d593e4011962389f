"""Parity of the CUDA local-map acquisition (hdsm_sense_batch, csrc/hdsm_sense.cu) through the C ABI against
oracle/sense_oracle.c, the sequential restatement of mapping_util/src/map_builder.cpp:80-205 (crop, RaycastAndClear,
MergeVoxelGrids) whose ray traversal is pinned to the reference's own raycast.cpp.  Bar: byte-exact grids and origins."""
import os

import numpy as np
import pytest

from multi_agent_pkgs_b200 import mapping as mp, sensing as sn
from oracle import mapping as om, sensing as osn
from test_sense_oracle import RANGE, VOX, forest_env, positions

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [0, 1])
def test_closed_loop_updates_match_the_checker(seed):
    """Five consecutive updates of 24 moving agents (kept grids merged in), more agents than resident blocks' worth of
    scratch is exercised by test_many_agents below."""
    env, org = forest_env(seed, floor=seed == 0)
    pos = positions(20 + seed, 24)
    rng = np.random.default_rng(seed)
    mb = sn.LocalMapBuilder(VOX, 24, RANGE)
    old_g = old_o = None
    for step in range(5):
        got_g, got_o = mb.update(env, org, pos)
        want_g, want_o = osn.c_update(env, org, pos, VOX, RANGE, old_grids=old_g, old_origin=old_o)
        assert np.array_equal(got_o, want_o), step
        assert np.array_equal(got_g, want_g), (step, int((got_g != want_g).sum()))
        old_g, old_o = want_g, want_o
        pos = pos + rng.uniform(-1.5, 1.5, pos.shape) * [1, 1, 0.1]
        if step == 2:
            pos[0, :2] += 25.0                      # a jump of more than a grid: nothing of the kept grid overlaps
    assert mb.launch_count == 5
    mb.close()


def test_known_map_limited_fov_and_other_shapes():
    env, org = forest_env(5)
    pos = positions(6, 9)
    mb = sn.LocalMapBuilder(VOX, 9, RANGE, free_grid=True)
    g, o = mb.update(env, org, pos)
    w, wo = osn.c_update(env, org, pos, VOX, RANGE, free_grid=True)
    assert np.array_equal(g, w) and np.array_equal(o, wo)
    mb.close()
    rng = np.random.default_rng(3)
    rot = np.zeros((9, 3, 3))
    for a in range(9):
        yaw = rng.uniform(-np.pi, np.pi)
        rot[a] = [[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]]
    for fov in ((1.57, 1.57), (1.0, 0.6)):
        mb = sn.LocalMapBuilder(VOX, 9, RANGE, limited_fov=True, fov_x=fov[0], fov_y=fov[1])
        g, o = mb.update(env, org, pos, rot)
        w, wo = osn.c_update(env, org, pos, VOX, RANGE, rot=rot, fov=fov)
        assert np.array_equal(g, w), fov
        g2, _ = mb.update(env, org, pos + 0.4, rot)
        w2, _ = osn.c_update(env, org, pos + 0.4, VOX, RANGE, rot=rot, fov=fov, old_grids=w, old_origin=wo)
        assert np.array_equal(g2, w2), fov
        mb.close()
    for vox, rng3 in ((0.2, (6.0, 5.0, 3.0)), (0.5, (10.0, 10.0, 0.5)), (0.3, (3.0, 20.0, 6.0))):
        mb = sn.LocalMapBuilder(vox, 9, rng3)
        g, o = mb.update(env, org, pos)
        w, wo = osn.c_update(env, org, pos, vox, rng3)
        assert np.array_equal(g, w) and np.array_equal(o, wo), (vox, rng3)
        mb.close()


def test_many_agents_reuse_the_key_scratch():
    """More agents than the kernel has resident blocks: every block's scratch is cleared between its agents."""
    env, org = forest_env(8)
    pos = positions(9, 700)
    mb = sn.LocalMapBuilder(VOX, 700, RANGE)
    g, o = mb.update(env, org, pos)
    w, wo = osn.c_update(env, org, pos, VOX, RANGE)
    assert np.array_equal(o, wo)
    assert np.array_equal(g, w), int((g != w).sum())
    g2, o2 = mb.update(env, org, pos + [0.7, -0.4, 0.05])
    w2, wo2 = osn.c_update(env, org, pos + [0.7, -0.4, 0.05], VOX, RANGE, old_grids=w, old_origin=wo)
    assert np.array_equal(g2, w2) and np.array_equal(o2, wo2)
    mb.close()


def test_acquired_grids_feed_the_post_processing():
    """hdsm_sense_batch -> hdsm_map_batch: the grid the planner receives (map_builder.cpp:80-216), byte for byte."""
    env, org = forest_env(4)
    pos = positions(5, 12)
    mb = sn.LocalMapBuilder(VOX, 12, RANGE)
    g, _ = mb.update(env, org, pos)
    mb.close()
    gen = mp.MapProcessor(VOX, 12, g[0].size)
    out = gen.process(g)
    gen.close()
    want = om.c_process(osn.c_update(env, org, pos, VOX, RANGE)[0], VOX, 0.3, 1.5, 4)
    assert np.array_equal(out, want)
    assert (out == -1).any() and (out == 100).any() and ((out > 0) & (out < 100)).any()


def test_golden_fixture_from_the_reference_node():
    """The CUDA path against tests/golden/mapbuilder_ref.npz, i.e. against the reference's own map-builder node: 18 agents x 3
    consecutive updates, 360 degree / limited field of view / known map, four grid shapes; byte-exact grids, bit-exact origins."""
    from test_sense_oracle import golden_cases
    n = 0
    for c in golden_cases():
        mb = sn.LocalMapBuilder(c["vox"], 1, c["rng3"], free_grid=c["free"], limited_fov=c["fov"] is not None,
                                fov_x=c["fov"][0] if c["fov"] else 1.57, fov_y=c["fov"][1] if c["fov"] else 1.57)
        for step in range(3):
            g, o = mb.update(c["env"], c["org"], c["pos"][step][None], c["rot"][None] if c["fov"] else None)
            assert np.array_equal(o[0], c["origin"][step]), (n, step)
            assert np.array_equal(g[0], c["cur"][step]), (n, step, int((g[0] != c["cur"][step]).sum()))
        mb.close()
        n += 1
    assert n == 18


def test_key_form_matches_the_checker(monkeypatch):
    """HDSM_SENSE_BITS=0: the kernel's first form (keys for every write), kept for A/B measurements."""
    monkeypatch.setenv("HDSM_SENSE_BITS", "0")
    env, org = forest_env(8)
    pos = positions(9, 300)
    mb = sn.LocalMapBuilder(VOX, 300, RANGE)
    g, o = mb.update(env, org, pos)
    w, wo = osn.c_update(env, org, pos, VOX, RANGE)
    assert np.array_equal(o, wo) and np.array_equal(g, w), int((g != w).sum())
    g2, o2 = mb.update(env, org, pos + [0.7, -0.4, 0.05])
    w2, wo2 = osn.c_update(env, org, pos + [0.7, -0.4, 0.05], VOX, RANGE, old_grids=w, old_origin=wo)
    assert np.array_equal(g2, w2) and np.array_equal(o2, wo2)
    mb.close()
    for vox, rng3 in ((0.2, (6.0, 5.0, 3.0)), (0.3, (3.0, 20.0, 6.0))):
        mb = sn.LocalMapBuilder(vox, 9, rng3)
        g, o = mb.update(env, org, pos[:9])
        w, wo = osn.c_update(env, org, pos[:9], vox, rng3)
        assert np.array_equal(g, w) and np.array_equal(o, wo), (vox, rng3)
        mb.close()


def test_sense_error_codes():
    with pytest.raises(ValueError):
        sn.grid_dims(0.0, RANGE)
    with pytest.raises(RuntimeError):
        sn.LocalMapBuilder(0.01, 2, (10.0, 10.0, 10.0))        # sides sum to more than 1400 voxels
    mb = sn.LocalMapBuilder(VOX, 2, RANGE)
    env, org = forest_env(0)
    with pytest.raises(RuntimeError, match="max_agents"):
        mb.update(env, org, positions(0, 3))
    mb.close()
    mb = sn.LocalMapBuilder(VOX, 2, RANGE, limited_fov=True)
    with pytest.raises(RuntimeError, match="camera rotation"):
        mb.update(env, org, positions(0, 2))
    mb.close()
