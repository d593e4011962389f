"""The reference's OWN planner node (agent_class.cpp compiled unmodified, oracle/_ref/libref_agent.so) running closed loop on an
external solver behind its Gurobi calls: model_.optimize() (agent_class.cpp:959) reaches the stand-in's solver hook, which packs
the node's members into the C-ABI arrays and calls hdsm_solve_batch (CUDA library, -m gpu) or the C port (CPU run); the
reference's own read-back (:962-987), fallback (:997-1019) and plane construction (:1086-1215) run on real numbers.  Ten nodes,
the config-2 circle, against the same loop driven from Python (Swarm.advance semantics): every trajectory bit for bit."""
import numpy as np
import pytest

from multi_agent_pkgs_b200 import scenarios as sc
from oracle import c_oracle as co, hdsm_oracle as ho, ref_agent as ra

STEPS = 30
BREAK = {(0, 3), (7, 3), (12, 5), (13, 5)}   # (step, agent): corridor moved away -> no solution -> fallback (step 0: no previous plan)


def _inputs(sw, i, pos, step):
    ref, polys = sw.agent_inputs(i, pos, step)
    if (step, i) in BREAK:
        polys = [(A, b - 50.0) for A, b in polys]
    return np.vstack([ref, ref[-1:]]), polys  # the node holds N + 1 reference points and reads the first N (:871-883)


def _run_nodes(solver_of):
    """solver_of(i) -> the `solver` argument of RefAgent.loop_step for node i."""
    sw = sc.config2_circle(n_swarms=1, seed=21)
    n, N = sw.n, sw.params["n_hor"]
    p = ho.Params(**sw.params)
    nodes = [ra.RefAgent(p, n, i, sw.state[i]) for i in range(n)]
    table, valid = np.zeros((n, N + 1, 3)), np.zeros(n, np.uint8)
    pos = sw.state[:, :3].copy()
    hist = []
    for step in range(STEPS):
        outs = []
        for i in range(n):
            ref, polys = _inputs(sw, i, pos[i], step)
            outs.append(nodes[i].loop_step(ref, polys, table, valid, solver_of(i)))
        for i, o in enumerate(outs):   # every node publishes after all have planned (PublishTrajectoryFull :185, :645-677)
            if o["have_traj"]:
                table[i], valid[i] = o["traj"][:, :3], 1
                pos[i] = o["traj"][1, :3]
        hist.append(outs)
    return sw, hist


def _run_python(solve_one):
    """The same closed loop with the read-back / fallback / advance written in Python (scenarios.Swarm.advance semantics)."""
    sw = sc.config2_circle(n_swarms=1, seed=21)
    n, N = sw.n, sw.params["n_hor"]
    state = sw.state.copy()
    traj, ctrl, have = np.zeros((n, N + 1, 9)), np.zeros((n, N, 3)), np.zeros(n, bool)
    table, valid = np.zeros((n, N + 1, 3)), np.zeros(n, np.uint8)
    hist = []
    for step in range(STEPS):
        new = []
        for i in range(n):
            if have[i]:
                state[i] = traj[i, 1]
            ref, polys = _inputs(sw, i, state[i, :3], step)
            b = sc.Batch(sw.params, np.array([i], np.int32), np.array([0], np.int32), np.array([n], np.int32), state[i:i + 1].copy(), ref[None, :N],
                         *sc.pack_polys([polys], sw.params["poly_hor"], 18), (traj[i, :, :3] if have[i] else np.repeat(sw.state[i, None, :3], N + 1, 0))[None],
                         table.copy(), np.where(np.arange(n) == i, 0, valid).astype(np.uint8), 18)
            new.append(solve_one(b))
        for i, o in enumerate(new):
            st = o["res"]["status"][0]
            ok = st == 0 or (st == 4 and np.isfinite(o["res"]["obj"][0]))
            if ok:
                traj[i], ctrl[i], have[i] = o["traj"][0], o["ctrl"][0], True
            elif have[i]:
                traj[i] = np.concatenate([traj[i, 1:], traj[i, -1:]])
                ctrl[i] = np.concatenate([ctrl[i, 1:], ctrl[i, -1:]])
            if have[i]:
                table[i], valid[i] = traj[i, :, :3], 1
        hist.append([dict(traj=traj[i].copy(), ctrl=ctrl[i].copy(), have=bool(have[i]), failed=not (new[i]["res"]["status"][0] == 0 or
                          (new[i]["res"]["status"][0] == 4 and np.isfinite(new[i]["res"]["obj"][0]))), x0=state[i].copy(),
                          used=new[i]["poly_used"][0].copy()) for i in range(n)])
    return hist


def _compare(hist_nodes, hist_py):
    n_failed = 0
    for step, (a, b) in enumerate(zip(hist_nodes, hist_py)):
        for i, (o, e) in enumerate(zip(a, b)):
            assert o["failed"] == e["failed"], (step, i)
            assert o["have_traj"] == e["have"], (step, i)
            assert np.array_equal(o["x0"], e["x0"]), (step, i)
            if e["have"]:
                assert np.array_equal(o["traj"], e["traj"]) and np.array_equal(o["ctrl"], e["ctrl"]), (step, i)
            if not e["failed"]:
                assert np.array_equal(o["poly_used"], e["used"]), (step, i)
            n_failed += e["failed"]
    assert n_failed >= len(BREAK)   # the fallback paths were exercised, with and without a previous plan
    assert not hist_py[0][3]["have"] and hist_py[1][3]["have"]


@pytest.mark.skipif(not ra.have_ref(), reason="oracle/_ref/libref_agent.so not built (needs /root/reference)")
def test_reference_node_closed_loop_on_the_c_port():
    sw0 = sc.config2_circle(n_swarms=1, seed=21)
    prm = co.make_params(sw0.params, max_nodes=64)
    _, hist = _run_nodes(lambda i: ("port", prm))
    _compare(hist, _run_python(lambda b: co.solve_batch(b, max_nodes=64, n_threads=1)))


@pytest.mark.gpu
@pytest.mark.skipif(not ra.have_ref(), reason="oracle/_ref/libref_agent.so not built (needs /root/reference)")
def test_reference_node_closed_loop_on_the_cuda_library():
    """F3 as far as it goes without ROS2: the reference's node, unmodified, planning on libhdsm through the C ABI."""
    from multi_agent_pkgs_b200.planner import TrajectoryPlanner
    sw0 = sc.config2_circle(n_swarms=1, seed=21)
    planners = [TrajectoryPlanner(sw0.params, 1, sw0.n, max_nodes=64) for _ in range(sw0.n)]   # one handle per node, like one GRBModel per node
    _, hist = _run_nodes(lambda i: ("hdsm", planners[i]))
    ref = TrajectoryPlanner(sw0.params, 1, sw0.n, max_nodes=64)
    _compare(hist, _run_python(ref.solve_batch))
    launches = sum(p.launch_count for p in planners)
    assert launches >= STEPS * sw0.n
    for p in planners + [ref]:
        p.close()
