"""Parity of the CUDA reference-trajectory generator (hdsm_reftraj_batch, csrc/hdsm_reftraj.cu) through the C ABI
against oracle/reftraj_oracle.c, whose ray caster is pinned to the reference's own Raycast.

Tolerances: the velocity goes through pow / exp (CUDA's math library vs the CPU's libm): path_vel to 1e-12
relative; everything else is +, -, *, /, sqrt in the reference's order without FMA contraction, so positions
and velocity references follow to 1e-11 m (they scale with path_vel).  Which sample is the first non-free one
(KeepOnlyFreeReference) must agree exactly.
"""
import numpy as np
import pytest

from multi_agent_pkgs_b200 import reftraj as rtj, scenarios as sc
from oracle import reftraj as ort

pytestmark = pytest.mark.gpu

VEL_RTOL = 1e-12
POS_ATOL = 1e-11


def _check(out, ref):
    assert np.abs(out["path_vel"] / ref["path_vel"] - 1).max() <= VEL_RTOL
    assert np.abs(out["ref"] - ref["ref"]).max() <= POS_ATOL * max(1.0, np.abs(ref["ref"]).max())


def _forest(n_swarms, seed=2):
    sw = sc.config2_circle(n_swarms=n_swarms, seed=seed)
    for i in range(sw.n):
        sw.state[i, :2] = sw.world.push_free(0.45 * sw.state[i, :2] + 0.55 * sw.goal[i, :2], 0.3)
    return sw


def test_first_and_second_step_match_the_checker():
    sw = _forest(6)
    rb = rtj.reftraj_batch(sw)
    gen = rtj.ReferenceTrajectoryGenerator(rb)
    out, ref = gen.generate(rb), ort.c_generate(rb)
    _check(out, ref)
    assert out["path_vel"].std() > 0.1 and gen.launch_count == 1
    rb2 = rtj.reftraj_batch(sw, prev_ref=ref["ref"][:, :, :3].copy())
    _check(gen.generate(rb2), ort.c_generate(rb2))
    rb2.increment[:] = 0
    _check(gen.generate(rb2), ort.c_generate(rb2))
    gen.close()


def test_neighbour_sweep_and_blocked_paths():
    sw = _forest(4, seed=5)
    rb = rtj.reftraj_batch(sw)
    rng = np.random.default_rng(1)
    rb.all_valid[:] = 1
    rb.all_pos[:] = rb.all_pos + rng.normal(0, 0.5, rb.all_pos.shape)  # plans close to each other inside a swarm
    for r in range(0, rb.n, 3):                                         # walls across every third path: collisions
        g = rb.grids[r]
        c = ((rb.path[r, 1] - rb.origins[r]) / rb.voxel).astype(int)
        g[:, max(c[1] - 8, 0):c[1] + 8, max(c[0] - 1, 0):c[0] + 1] = 100
    gen = rtj.ReferenceTrajectoryGenerator(rb)
    out, ref = gen.generate(rb), ort.c_generate(rb)
    _check(out, ref)
    assert (ref["path_vel"] < 0.5 * (rb.path_vel_min + rb.path_vel_max)).any()
    rb.sens_other_agents = 0.8                                          # weights decaying along the horizon
    gen.close()
    gen = rtj.ReferenceTrajectoryGenerator(rb)
    _check(gen.generate(rb), ort.c_generate(rb))
    gen.close()


def test_reference_feeds_the_optimisation():
    """The first N rows of `ref` are hdsm_solve_batch's ref input: same plans as with the checker's reference."""
    from multi_agent_pkgs_b200.planner import TrajectoryPlanner
    sw = sc.config2_circle(n_swarms=2)
    rb = rtj.reftraj_batch(sw)
    gen = rtj.ReferenceTrajectoryGenerator(rb)
    out, ref = gen.generate(rb), ort.c_generate(rb)
    gen.close()
    b = sw.make_batch()
    pl = TrajectoryPlanner(sw.params, max_agents=b.n, max_neighbours=10, max_nodes=5000)
    b.ref = np.ascontiguousarray(out["ref"][:, :sw.params["n_hor"]])
    a = pl.solve_batch(b)
    b.ref = np.ascontiguousarray(ref["ref"][:, :sw.params["n_hor"]])
    c = pl.solve_batch(b)
    pl.close()
    assert np.array_equal(a["res"]["status"], c["res"]["status"]) and (a["res"]["status"] == 0).mean() > 0.7
    ok = a["res"]["status"] == 0
    assert np.abs(a["traj"][ok] - c["traj"][ok]).max() <= 1e-6


def test_reftraj_error_codes():
    sw = sc.config2_circle(n_swarms=1)
    rb = rtj.reftraj_batch(sw)
    gen = rtj.ReferenceTrajectoryGenerator(rb, max_agents=4)
    with pytest.raises(RuntimeError, match="max_agents|capacity"):
        gen.generate(rb)
    gen.close()
    gen = rtj.ReferenceTrajectoryGenerator(rb)
    rb.n_path[3] = 0
    with pytest.raises(RuntimeError, match="n_path"):
        gen.generate(rb)
    gen.close()
