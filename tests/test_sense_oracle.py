"""Local-map acquisition checkers (oracle/sense_oracle.c: crop, RaycastAndClear, MergeVoxelGrids of
mapping_util/src/map_builder.cpp:80-205, restated in the reference's sequential order) and the kernel's own per-ray code
(csrc/hdsm_sense_core.h) compiled for the CPU and run in a scrambled ray order (oracle/sense_emu.cpp).

PINNED: the whole stage against the reference's own MapBuilder::EnvironmentVoxelGridCallback - map_builder.cpp, path_tools.cpp,
raycast.cpp and voxel_grid.cpp compiled unmodified on stand-in ROS / Eigen headers (oracle/_ref/libref_map.so) - through the
committed fixture tests/golden/mapbuilder_ref.npz and, where the reference is present, live."""
import os

import numpy as np
import pytest

from oracle import sensing as S

VOX, RANGE = 0.3, (20.0, 20.0, 6.0)


def forest_env(seed, n_cols=300, shape=(24, 140, 140), floor=True):
    rng = np.random.default_rng(seed)
    env = np.zeros(shape, np.int8)
    for _ in range(n_cols):
        env[:, rng.integers(0, shape[1]), rng.integers(0, shape[2])] = 100
    if floor:
        env[0] = 100
    return env, np.array([-3.0, -3.0, -0.3])


def positions(seed, n):
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(0, 38, n), rng.uniform(0, 38, n), rng.uniform(0.5, 4, n)], 1)


def test_frame_and_crop_known_map():
    """free_grid: the crop itself with unknown -> free (map_builder.cpp:89-153), checked against numpy slicing."""
    env, org = forest_env(1)
    pos = positions(2, 5)
    got, o = S.c_update(env, org, pos, VOX, RANGE, free_grid=True)
    assert got.shape == (5, 20, 66, 66)
    for a in range(5):
        want_o = np.round((pos[a] - np.array(RANGE) / 2 - org) / VOX) * VOX + org
        assert np.array_equal(o[a], want_o)
        s = np.round((want_o - org) / VOX).astype(int)
        pad = np.zeros((20 + 200, 66 + 200, 66 + 200), np.int8)          # outside the environment: unknown -> free
        pad[100:100 + env.shape[0], 100:100 + env.shape[1], 100:100 + env.shape[2]] = env
        want = pad[100 + s[2]:120 + s[2], 100 + s[1]:166 + s[1], 100 + s[0]:166 + s[0]]
        assert np.array_equal(got[a], want)
    emu, eo = S.emu_update(env, org, pos, VOX, RANGE, free_grid=True)
    assert np.array_equal(emu, got) and np.array_equal(eo, o)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_kernel_code_in_scrambled_order_equals_sequential_reference_order(seed):
    env, org = forest_env(seed, floor=seed != 2)
    pos = positions(10 + seed, 6)
    g1, o1 = S.c_update(env, org, pos, VOX, RANGE)
    g2, o2 = S.emu_update(env, org, pos, VOX, RANGE, seed=seed + 1)
    assert np.array_equal(g1, g2) and np.array_equal(o1, o2)
    assert set(np.unique(g1)) <= {-1, 0, 100}
    # second and third update: merge with the kept grid after the agents moved (also by more than a grid)
    rng = np.random.default_rng(seed)
    for step in range(2):
        pos = pos + rng.uniform(-1.5, 1.5, pos.shape) * [1, 1, 0.1]
        if step == 1:
            pos[0, :2] += 25.0
        g3, o3 = S.c_update(env, org, pos, VOX, RANGE, old_grids=g1, old_origin=o1)
        g4, o4 = S.emu_update(env, org, pos, VOX, RANGE, old_grids=g1, old_origin=o1, seed=7 * seed + step)
        assert np.array_equal(g3, g4) and np.array_equal(o3, o4)
        assert (g3 != -1).sum() >= (S.c_update(env, org, pos, VOX, RANGE)[0] != -1).sum() - 125 * len(pos)
        g1, o1 = g3, o3


def test_bits_form_of_the_kernel_logic_and_how_often_write_order_matters():
    """The kernel's second form (free / occupied bitmaps first, keys only for voxels written both ways) gives the same grids;
    and such voxels do occur in forest worlds - the reason the writes carry their position in the reference's ray order."""
    env, org = forest_env(3)
    pos = positions(4, 48)
    want, wo = S.c_update(env, org, pos, VOX, RANGE)
    conflicts = np.zeros(48, np.int64)
    got, go = S.emu_update(env, org, pos, VOX, RANGE, seed=11, bits_form=True, conflicts=conflicts)
    assert np.array_equal(got, want) and np.array_equal(go, wo)
    assert (conflicts > 0).any() and (conflicts == 0).any()
    pos2 = pos + 0.37
    want2, _ = S.c_update(env, org, pos2, VOX, RANGE, old_grids=want, old_origin=wo)
    got2, _ = S.emu_update(env, org, pos2, VOX, RANGE, old_grids=want, old_origin=wo, seed=12, bits_form=True)
    assert np.array_equal(got2, want2)
    for vox, rng3 in ((0.2, (6.0, 5.0, 3.0)), (0.5, (10.0, 10.0, 0.5)), (0.3, (3.0, 20.0, 6.0))):
        a, ao = S.c_update(env, org, pos[:8], vox, rng3)
        b, bo = S.emu_update(env, org, pos[:8], vox, rng3, seed=2, bits_form=True)
        assert np.array_equal(a, b) and np.array_equal(ao, bo)


def test_limited_field_of_view_and_other_shapes():
    env, org = forest_env(5)
    pos = positions(6, 4)
    rng = np.random.default_rng(3)
    rot = np.zeros((4, 3, 3))
    for a in range(4):
        yaw = rng.uniform(-np.pi, np.pi)
        rot[a] = [[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]]
    full, _ = S.c_update(env, org, pos, VOX, RANGE)
    for fov in ((1.57, 1.57), (1.0, 0.6)):
        g1, _ = S.c_update(env, org, pos, VOX, RANGE, rot=rot, fov=fov)
        g2, _ = S.emu_update(env, org, pos, VOX, RANGE, rot=rot, fov=fov, seed=4)
        assert np.array_equal(g1, g2)
        assert ((g1 == 0).sum() < (full == 0).sum()) and (g1 == 0).sum() > 125 * 4
    for vox, rng3 in ((0.2, (6.0, 5.0, 3.0)), (0.5, (10.0, 10.0, 0.5)), (0.3, (3.0, 20.0, 6.0))):
        g1, o1 = S.c_update(env, org, pos, vox, rng3)
        g2, o2 = S.emu_update(env, org, pos, vox, rng3, seed=2)
        assert g1.shape[1:] == S.local_dims(vox, rng3)[::-1]
        assert np.array_equal(g1, g2) and np.array_equal(o1, o2)


def test_properties_of_a_first_update():
    """An empty world is seen completely; a wall hides what is behind it; the agent's 5x5x5 neighbourhood is known free
    on the first update (ClearVoxelsCenter) even when every ray is blocked."""
    env = np.zeros((24, 140, 140), np.int8)
    org = np.array([-3.0, -3.0, -0.3])
    pos = np.array([[18.0, 18.0, 2.0]])
    g, o = S.c_update(env, org, pos, VOX, RANGE)
    assert (g == 0).all()
    env[:, :, 75] = 100                                   # a wall across x = -3 + 75 * 0.3 = 19.5 m
    g, o = S.c_update(env, org, pos, VOX, RANGE)
    ix = int(round((19.5 - o[0, 0]) / VOX))
    assert (g[0, :, :, ix] == 100).any() and (g[0, :, :, ix + 1:] == -1).all() and (g[0, :, :, :ix] == 0).all()
    boxed = np.full((24, 140, 140), 100, np.int8)         # the agent inside solid matter: start voxel occupied
    g, _ = S.c_update(boxed, org, pos, VOX, RANGE)
    assert (g != -1).sum() == 125 and (g == 100).sum() >= 1 and (g == 0).sum() > 100
    g2, _ = S.emu_update(boxed, org, pos, VOX, RANGE)
    assert np.array_equal(g, g2)


def test_rays_of_this_stage_against_the_reference_ray_caster():
    """The traversal under ClearLine (border-voxel centres from a grid-centre start) against the reference's own Raycast."""
    from oracle import reftraj as ort
    if not ort.have_ref():
        pytest.skip("oracle/_ref/libref_voxel.so not built (needs /root/reference)")
    rng = np.random.default_rng(0)
    g = np.zeros((20, 66, 66), np.int8)
    g[rng.random(g.shape) < 0.01] = 100
    g[:, 30:36, 30:36] = 0
    start = np.array([33.2, 32.7, 10.4])
    n = 0
    for i in range(0, 66, 5):
        for j in range(0, 66, 7):
            for k in (0, 19):
                end = np.array([i + 0.5, j + 0.5, k + 0.5])
                a = ort.c_raycast(g, start, end, float(np.linalg.norm(start - end)))
                b = ort.ref_raycast(g, start, end, float(np.linalg.norm(start - end)))
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
                n += 1
    assert n > 100


# ---- pinned against the reference's own map-builder node ------------------------------------------------------------
def golden_cases():
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "mapbuilder_ref.npz"))
    for k in range(int(z["n_cases"])):
        p = z[f"c{k}_prm"]
        yield dict(vox=float(p[0]), rng3=tuple(float(v) for v in p[1:4]), free=bool(p[4]), fov=(float(p[6]), float(p[7])) if p[5] else None,
                   infl=float(p[8]), pot=float(p[9]), pw=float(p[10]), env=z[f"c{k}_env"], org=z[f"c{k}_org"], rot=z[f"c{k}_rot"],
                   pos=z[f"c{k}_pos"], cur=z[f"c{k}_cur"], origin=z[f"c{k}_origin"], pub=z[f"c{k}_pub"])


def test_golden_fixture_from_the_reference_node():
    """tests/golden/mapbuilder_ref.npz was produced by the reference's OWN MapBuilder::EnvironmentVoxelGridCallback
    (map_builder.cpp compiled unmodified, tests/golden/make_mapbuilder_golden.py): 18 agents x 3 consecutive updates, 360 degree /
    limited field of view / known map, four grid shapes.  The sequential restatement, the kernel's code in scrambled order (both
    forms) and the post-processing restatement must reproduce voxel_grid_curr_, its origin and the published grid byte for byte."""
    from oracle import mapping as om
    n = 0
    for c in golden_cases():
        old_g = old_o = None
        for step in range(3):
            kw = dict(free_grid=c["free"], rot=c["rot"][None] if c["fov"] else None, fov=c["fov"],
                      old_grids=None if old_g is None else old_g[None], old_origin=None if old_o is None else old_o[None])
            g, o = S.c_update(c["env"], c["org"], c["pos"][step][None], c["vox"], c["rng3"], **kw)
            assert np.array_equal(o[0], c["origin"][step]) and np.array_equal(g[0], c["cur"][step]), (n, step)
            for bits in (False, True):
                e, eo = S.emu_update(c["env"], c["org"], c["pos"][step][None], c["vox"], c["rng3"], seed=n + step, bits_form=bits, **kw)
                assert np.array_equal(e, g) and np.array_equal(eo, o), (n, step, bits)
            pub = om.c_process(g, c["vox"], c["infl"], c["pot"], int(c["pw"]))
            assert np.array_equal(pub[0], c["pub"][step]), (n, step)
            old_g, old_o = g[0], o[0]
        n += 1
    assert n == 18


def test_live_against_the_compiled_reference_node():
    """Same comparison on fresh random inputs at the planner's grid size, where oracle/_ref/libref_map.so exists."""
    if not S.have_ref():
        pytest.skip("oracle/_ref/libref_map.so not built (needs /root/reference)")
    from oracle import mapping as om
    env, org = forest_env(21)
    pos = positions(22, 4)
    rot = np.array([[0.6, -0.8, 0], [0.8, 0.6, 0], [0, 0, 1.0]])
    for fov in (None, (1.0, 0.6)):
        w, wo = S.c_update(env, org, pos, VOX, RANGE, rot=None if fov is None else np.stack([rot] * 4), fov=fov)
        pos2 = pos + [0.5, -0.7, 0.04]
        w2, wo2 = S.c_update(env, org, pos2, VOX, RANGE, rot=None if fov is None else np.stack([rot] * 4), fov=fov, old_grids=w, old_origin=wo)
        pub2 = om.c_process(w2, VOX, 0.3, 1.5, 4)
        for a in range(4):
            cur, o, _ = S.ref_update(env, org, pos[a], VOX, RANGE, rot=rot if fov else None, fov=fov)
            assert np.array_equal(cur, w[a]) and np.array_equal(o, wo[a]), (fov, a)
            cur, o, pub = S.ref_update(env, org, pos2[a], VOX, RANGE, rot=rot if fov else None, fov=fov, old_grid=w[a], old_origin=wo[a])
            assert np.array_equal(cur, w2[a]) and np.array_equal(o, wo2[a]) and np.array_equal(pub, pub2[a]), (fov, a)


def test_mixed_first_and_follow_up_agents_in_one_call():
    """have_old is per agent: agents on their first update (all-unknown grid with the freed 5x5x5 cube) and agents with a kept grid
    in the same batch."""
    env, org = forest_env(7)
    pos = positions(8, 6)
    g0, o0 = S.c_update(env, org, pos, VOX, RANGE)
    have = np.array([1, 0, 1, 0, 0, 1], np.uint8)
    pos2 = pos + 0.45
    a, ao = S.c_update(env, org, pos2, VOX, RANGE, old_grids=g0, old_origin=o0, have_old=have)
    b, bo = S.emu_update(env, org, pos2, VOX, RANGE, old_grids=g0, old_origin=o0, have_old=have, seed=3, bits_form=True)
    assert np.array_equal(a, b) and np.array_equal(ao, bo)
    first, _ = S.c_update(env, org, pos2, VOX, RANGE)
    kept, _ = S.c_update(env, org, pos2, VOX, RANGE, old_grids=g0, old_origin=o0)
    for i in range(6):
        assert np.array_equal(a[i], kept[i] if have[i] else first[i]), i
