"""Regenerates tests/golden/reftraj_node_ref.npz from the REFERENCE ITSELF: the reference's own Agent::GenerateReferenceTrajectory
(multi_agent_planner/src/agent_class.cpp:1449-1553 with SamplePath, KeepOnlyFreeReference, ComputePathVelocity, GetVelocityLimit;
agent_class.cpp compiled unmodified into oracle/_ref/libref_agent.so, `make -C oracle ref`) on agents of the 10-agent forest scenario:
first call, follow-up call, follow-up call with increment_traj_ref_.

    python tests/golden/make_reftraj_node_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from multi_agent_pkgs_b200 import reftraj as rtj, scenarios as sc  # noqa: E402
from oracle import c_oracle as co, hdsm_oracle as ho, ref_agent as ra  # noqa: E402


def scenario_batch(steps=2):
    sw = sc.config2_circle(n_swarms=1)
    for _ in range(steps):
        b = sw.make_batch()
        out = co.solve_batch(b)
        sw.advance(out["traj"], out["ctrl"], out["res"]["status"] == 0)
    return rtj.reftraj_batch(sw)


def run_reference(rb, i, prev_ref, increment):
    ag = ra.RefAgent(ho.Params(), rb.all_pos.shape[0], int(rb.global_id[i]), [0.0] * 9)
    gi = i if rb.grid_index is None else int(rb.grid_index[i])
    return ag.reference_trajectory(rb.grids[gi], rb.origins[i], rb.voxel, rb.path[i, :rb.n_path[i]], prev_ref, increment,
                                   rb.traj[i] if rb.traj.shape[1] else None, rb.all_pos, rb.all_valid, rb.path_vel_min, rb.path_vel_max,
                                   rb.path_vel_dec, rb.sens_dist, rb.sens_pot, rb.sens_other_agents)


RB_ARRAYS = ("grids", "dims", "origins", "path", "n_path", "prev_ref", "have_prev", "increment", "traj", "global_id", "nbr_begin",
             "nbr_end", "all_pos", "all_valid")


def save_batch(rb):
    """The scenario batch itself goes into the fixture: it comes out of two closed-loop steps of the solver, whose tie-breaking
    may change from one version to the next, and the fixture must not move with it."""
    out = {f"rb_{k}": np.asarray(getattr(rb, k)) for k in RB_ARRAYS}
    out["rb_scalars"] = np.array([rb.n_hor, rb.dt, rb.voxel])
    return out


def load_batch(z):
    n_hor, dt, voxel = z["rb_scalars"]
    a = {k: z[f"rb_{k}"] for k in RB_ARRAYS}
    return rtj.RefTrajBatch(int(n_hor), float(dt), float(voxel), a["grids"], None, a["dims"], a["origins"], a["path"], a["n_path"], a["prev_ref"],
                            a["have_prev"], a["increment"], a["traj"], a["global_id"], a["nbr_begin"], a["nbr_end"], a["all_pos"], a["all_valid"])


def main():
    rb = scenario_batch()
    agents = [0, 3, 7]
    out = {"agents": np.array(agents)}
    out.update(save_batch(rb))
    for i in agents:
        first, v0 = run_reference(rb, i, None, 0)
        out[f"a{i}_first"], out[f"a{i}_first_vel"] = first, np.array(v0)
        for inc in (0, 1):
            nxt, v = run_reference(rb, i, first[:, :3], inc)
            out[f"a{i}_next{inc}"], out[f"a{i}_next{inc}_vel"] = nxt, np.array(v)
    path = os.path.join(ROOT, "tests", "golden", "reftraj_node_ref.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {len(agents)} agents x 3 calls to {path} ({os.path.getsize(path) / 1024:.0f} KB)")


if __name__ == "__main__":
    main()
