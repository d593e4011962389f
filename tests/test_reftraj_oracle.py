"""The reference-trajectory checkers (no GPU needed).

* oracle/reftraj_oracle.c::rt_raycast against tests/golden/raycast_ref.npz - fixtures produced by the reference's
  OWN voxel_grid_util::Raycast (raycast.cpp:21-186 + voxel_grid.cpp, compiled unmodified into oracle/_ref/) - and,
  where oracle/_ref exists, live against it.  Bar: bit-exact visited points (in order) and collision point.
* rt_generate (GenerateReferenceTrajectory, agent_class.cpp:1449-1553): properties of the sampling.
"""
import importlib.util
import os

import numpy as np
import pytest

from conftest import GOLDEN
from multi_agent_pkgs_b200 import reftraj as rtj, scenarios as sc
from oracle import reftraj as ort


def test_raycast_matches_reference_golden():
    z = np.load(os.path.join(GOLDEN, "raycast_ref.npz"))
    n = int(z["n_cases"])
    hits = 0
    for k in range(n):
        ray = z[f"ray{k}"]
        vis, col, cnt = ort.c_raycast(z[f"grid{k}"], ray[0:3], ray[3:6], float(ray[6]))
        assert np.array_equal(vis, z[f"vis{k}"], equal_nan=True) and np.array_equal(col, z[f"col{k}"], equal_nan=True), k
        hits += z[f"col{k}"][0] != -1
    assert n >= 160 and hits >= 20


@pytest.mark.skipif(not ort.have_ref(), reason="oracle/_ref not built (reference checkout absent)")
def test_raycast_matches_reference_live():
    spec = importlib.util.spec_from_file_location("mk", os.path.join(GOLDEN, "make_raycast_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    rng = np.random.default_rng(99)
    for t in range(1500):
        g, s, e, md = mk.random_case(rng, t)
        a, b = ort.ref_raycast(g, s, e, md), ort.c_raycast(g, s, e, md)
        assert a[2] == b[2] and np.array_equal(a[0], b[0], equal_nan=True) and np.array_equal(a[1], b[1], equal_nan=True), t


@pytest.fixture(scope="module")
def forest():
    sw = sc.config2_circle(n_swarms=3)
    for i in range(sw.n):
        sw.state[i, :2] = sw.world.push_free(0.45 * sw.state[i, :2] + 0.55 * sw.goal[i, :2], 0.3)
    return sw


def test_reference_trajectory_properties(forest):
    sw = forest
    rb = rtj.reftraj_batch(sw)
    out = ort.c_generate(rb)
    ref, vel = out["ref"], out["path_vel"]
    N = rb.n_hor
    assert ((vel >= rb.path_vel_min - 1e-12) & (vel <= rb.path_vel_max + 1e-12)).all() and vel.std() > 0.1
    assert np.array_equal(ref[:, 0, :3], rb.path[:, 0])                       # starts at the path start (:1472-1475)
    sp = np.linalg.norm(np.diff(ref[:, :, :3], axis=1), axis=2)
    assert (sp <= vel[:, None] * rb.dt + 1e-9).all()                           # never faster than path_vel (corners cut)
    assert (np.abs(sp - vel[:, None] * rb.dt) < 1e-9).mean() > 0.6             # exactly path_vel * dt on straight parts (rest: corners, repeats)
    for i in range(rb.n):                                                      # velocity reference: backwards along the path
        for k in range(N):
            d = ref[i, k, :3] - ref[i, k + 1, :3]
            if np.linalg.norm(d) > 1e-2:
                assert np.allclose(ref[i, k, 3:], vel[i] * d / np.linalg.norm(d), rtol=0, atol=1e-9)
        assert np.array_equal(ref[i, N, 3:], ref[i, N - 1, 3:])
        for k in range(N + 1):                                                 # samples lie in free voxels (:1665-1693)
            c = ((ref[i, k, :3] - rb.origins[i]) / rb.voxel).astype(int)
            assert rb.grids[i][c[2], c[1], c[0]] not in (100, -1)


def test_second_step_starts_from_the_previous_reference(forest):
    sw = forest
    first = ort.c_generate(rtj.reftraj_batch(sw))
    rb2 = rtj.reftraj_batch(sw, prev_ref=first["ref"][:, :, :3].copy())
    second = ort.c_generate(rb2)
    assert np.array_equal(second["ref"][:, 0, :3], first["ref"][:, 1, :3])     # increment_traj_ref_: start = old point 1
    rb2.increment[:] = 0
    assert np.array_equal(ort.c_generate(rb2)["ref"][:, 0, :3], first["ref"][:, 0, :3])


def test_neighbours_slow_the_agent_down(forest):
    sw = forest
    rb = rtj.reftraj_batch(sw)
    free = ort.c_generate(rb)["path_vel"].copy()
    rb.all_valid[:] = 1
    rb.all_pos[:] = rb.all_pos[0] + np.array([0.4, 0.0, 0.0])                  # everybody next to agent 0
    rb.all_pos[0] -= np.array([0.4, 0.0, 0.0])
    close = ort.c_generate(rb)["path_vel"]
    assert close[0] < free[0] - 0.5 or free[0] < rb.path_vel_min + 0.6
    rb.all_valid[:] = 0
    assert np.array_equal(ort.c_generate(rb)["path_vel"], free)
