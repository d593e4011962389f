"""Closed loop with the swarm state in HBM (swarm.ClosedLoop), the device half of the failure fallback
(agent_class.cpp:997-1019), K1's plane coefficients against the reference's libm chain (:1159-1170), a complete
4096-agent config-5 step, the handle-lifetime fixes and - on boxes with two GPUs - the NCCL exchange against the gloo path.
Run on the B200 box: -m gpu."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from multi_agent_pkgs_b200 import scenarios as sc
from multi_agent_pkgs_b200._lib import OPTIMAL, RESULT_DTYPE
from multi_agent_pkgs_b200.planner import TrajectoryPlanner

pytestmark = pytest.mark.gpu


def _ok(res):
    return (res["status"] == 0) | ((res["status"] == 4) & np.isfinite(res["obj"]))


def test_closed_loop_on_device_equals_host_closed_loop_with_the_c_port():
    """Six steps of a 300-agent config-5 swarm: ClosedLoop (solve, advance and table on the device) against the host
    loop Swarm.make_batch -> C port -> Swarm.advance.  Status exact, objective 1e-6, states 1e-3; replay == pre-roll."""
    import torch
    from multi_agent_pkgs_b200.swarm import ClosedLoop, table_checksum
    from oracle import c_oracle as co
    sw = sc.config5_random(seed=11, n_rob=300, side=60.0)
    host = sc.config5_random(seed=11, n_rob=300, side=60.0)
    loop = ClosedLoop(sw, 1, 0, "cuda:0", 64, None)
    loop.checker = lambda b: co.solve_batch(b, max_nodes=64)
    par = loop.preroll(6, parity_sample=300)
    assert par["agents"] == 6 * 300 and par["status_mismatches"] == 0 and par["max_rel_obj_gap"] <= 1e-6, par
    for s in range(6):
        b = host.make_batch()
        ref = co.solve_batch(b, max_nodes=64)
        assert np.array_equal(loop.stats[s]["status"], ref["res"]["status"]), s
        host.advance(ref["traj"], ref["ctrl"], _ok(ref["res"]))
    assert np.abs(loop.t["x0"].cpu().numpy() - host.state).max() <= 1e-3
    assert np.array_equal(loop.t["have_plan"].cpu().numpy(), host.have_plan)
    assert np.abs(loop.t["traj_curr"].cpu().numpy()[..., :3] - host.traj[..., :3]).max() <= 1e-3
    final = loop.sums[-1]
    loop.reset()
    with torch.cuda.stream(loop.stream):
        for s in range(6):
            loop.device_step(s)
    loop.stream.synchronize()
    assert table_checksum(loop.current_table()) == final  # deterministic: the device-only replay is the same closed loop
    loop.close()


def test_failure_fallback_on_the_device():
    """A7: agents whose optimisation fails get the previous plan shifted by one step with the last element duplicated
    (agent_class.cpp:1004-1013) - in pos_out (the solver's epilogue) and in traj_curr / ctrl_curr / x0
    (hdsm_advance_device); agents that never had a plan publish nothing and keep their state (:180-190)."""
    import torch
    sw = sc.config2_circle(n_swarms=2, seed=7)
    pl = TrajectoryPlanner(sw.params, max_agents=sw.n, max_neighbours=10)
    b = sw.make_batch()
    out = pl.solve_batch(b)
    sw.advance(out["traj"], out["ctrl"], _ok(out["res"]))
    b = sw.make_batch()
    n, N = b.n, 10
    bad = np.array([1, 4, 13])             # corridors these agents are not in: infeasible
    b.poly_b[bad] -= 50.0
    dev = torch.device("cuda:0")
    from multi_agent_pkgs_b200.swarm import DeviceBatch
    db = DeviceBatch(b, dev)
    t = db.t
    prev_traj, prev_ctrl = sw.traj.copy(), sw.ctrl.copy()
    have = np.ones(n, np.uint8)
    have[13] = 0                            # agent 13 fails without a previous plan
    t["traj_curr"] = torch.from_numpy(prev_traj).to(dev)
    t["ctrl_curr"] = torch.from_numpy(prev_ctrl).to(dev)
    t["have_plan"] = torch.from_numpy(have.copy()).to(dev)
    x0_before = b.x0.copy()
    stream = torch.cuda.Stream(device=dev)   # a real stream: 0 would select the handle's own one
    torch.cuda.synchronize()                 # uploads above ran on the default stream
    st = stream.cuda_stream
    pl.solve_batch_device(t, db.n_rob, st)
    prev_pos = t.pop("prev_self_pos")
    with torch.cuda.stream(stream):
        t["prev_self_pos"] = torch.zeros_like(prev_pos)
    pl.advance_device(t, st)
    torch.cuda.synchronize()
    res = db.results()
    assert (res["status"][bad] == 1).all() and (res["status"][np.setdiff1d(np.arange(n), bad)] == 0).all()
    pos_out, traj, ctrl = t["pos_out"].cpu().numpy(), t["traj"].cpu().numpy(), t["ctrl"].cpu().numpy()
    tc, cc, x0 = t["traj_curr"].cpu().numpy(), t["ctrl_curr"].cpu().numpy(), t["x0"].cpu().numpy()
    hv, pp = t["have_plan"].cpu().numpy(), t["prev_self_pos"].cpu().numpy()
    for i in range(n):
        if i in (1, 4):
            want_t = np.concatenate([prev_traj[i, 1:], prev_traj[i, -1:]])
            want_c = np.concatenate([prev_ctrl[i, 1:], prev_ctrl[i, -1:]])
            assert np.array_equal(pos_out[i], np.concatenate([b.prev_self_pos[i, 1:], b.prev_self_pos[i, -1:]]))
            assert np.array_equal(tc[i], want_t) and np.array_equal(cc[i], want_c)
            assert np.array_equal(x0[i], want_t[1]) and hv[i] == 1 and np.array_equal(pp[i], want_t[:, :3])
        elif i == 13:
            assert np.array_equal(tc[i], prev_traj[i]) and np.array_equal(x0[i], x0_before[i]) and hv[i] == 0
            assert np.array_equal(pp[i], np.repeat(x0_before[i, None, :3], N + 1, 0))
        else:
            assert np.array_equal(tc[i], traj[i]) and np.array_equal(cc[i], ctrl[i]) and np.array_equal(x0[i], traj[i, 1])
            assert np.array_equal(pos_out[i], traj[i, :, :3]) and np.array_equal(pp[i], traj[i, :, :3]) and hv[i] == 1
    pl.close()


def test_plane_coefficients_match_the_reference_chain():
    """K1 against the reference's acos / tan / atan / cos / sin / hypot chain (agent_class.cpp:1159-1170; the C port keeps
    the chain): random pairs, near-vertical pairs (where tan(ang) blows up), touching pairs (|n| < 2 s) - and the
    planes the reference's own GenerateTimeAwareSafeCorridor produced (tests/golden/agent_model_ref.npz)."""
    from oracle import c_oracle as co
    rng = np.random.default_rng(3)
    for params in (sc.agile_params(10), sc.crazyflie_params(9), dict(sc.agile_params(10), drone_radius=0.4, drone_z_offset=0.15)):
        pl = TrajectoryPlanner(params, max_agents=1, max_neighbours=1)
        pc = rng.uniform(-30, 30, (6000, 3))
        d = rng.normal(size=(6000, 3))
        d /= np.linalg.norm(d, axis=1)[:, None]
        dist = rng.uniform(0.05, 12.0, 6000)
        d[2000:3000] = [0, 0, 1]                                           # vertical and near-vertical
        d[2000:3000, :2] = rng.normal(size=(1000, 2)) * 10.0 ** rng.uniform(-17, -2, (1000, 1))
        d[2500:3000, 2] = -1
        dist[3000:4000] = rng.uniform(1e-6, 2 * max(params["drone_radius"], params["drone_z_offset"]), 1000)   # closer than 2 s
        d[4000:4500] = rng.permutation(np.eye(3))[0]                        # axis parallel
        po = pc + d * dist[:, None]
        got = pl.planes(pc, po)
        want = np.array([np.r_[co.plane(params, pc[i], po[i])] for i in range(len(pc))])
        scale = np.maximum(1.0, np.abs(want).max(axis=1, keepdims=True))
        assert np.isfinite(got).all()
        assert (np.abs(got - want) / scale).max() <= 1e-13, (np.abs(got - want) / scale).max()
        assert np.isnan(pl.planes(pc[:3], pc[:3])).all()                    # coincident points: flagged, not a plane
        pl.close()
    # the planes the reference's own GenerateTimeAwareSafeCorridor appended to its polytopes (recorded fixture):
    # rows after the static rows of final polytope 0 of step k, one per valid neighbour in id order
    from test_ref_agent import cases
    checked = 0
    for c in cases():
        p, N = c["p"], c["p"].n_hor
        prm = dict(sc.agile_params(N), drone_radius=p.drone_radius, drone_z_offset=p.drone_z_offset, tilt=p.tilt, poly_hor=p.poly_hor, dt=p.dt)
        pl = TrajectoryPlanner(prm, max_agents=1, max_neighbours=1)
        prev_pos = c["prev"][:, :3] if c["prev"] is not None else np.tile(c["state_ini"][:3], (N + 1, 1))
        others = [j for j in range(c["n_rob"]) if c["all_valid"][j] and j != c["id"]]
        R = len(c["polys"][0][1])
        for k in range(N):
            A, b = c["final"][k][0]
            got = pl.planes(np.tile(prev_pos[k + 1], (len(others), 1)), c["all_pos"][others, k + 1])
            want = np.c_[A[R:], b[R:]]
            assert got.shape == want.shape
            assert (np.abs(got - want) / np.maximum(1.0, np.abs(want).max(axis=1, keepdims=True))).max() <= 1e-13, k
            checked += len(others)
        pl.close()
    assert checked > 0


def test_config5_step_at_full_size():
    """All 4096 agents of BASELINE config 5, 4096 neighbour candidates each: first step (no planes) and second step
    (planes from the first step's plans) against the C port - status exact, objective 1e-6."""
    from oracle import c_oracle as co
    sw = sc.config5_random(n_rob=4096)
    pool = None
    pl = TrajectoryPlanner(sw.params, max_agents=4096, max_neighbours=4096, max_nodes=64)
    for step in range(2):
        b = sw.make_batch()
        out = pl.solve_batch(b)
        ref = co.solve_batch(b, max_nodes=64)
        assert np.array_equal(out["res"]["status"], ref["res"]["status"]), step
        ok = ref["res"]["status"] == OPTIMAL
        gap = np.abs(out["res"]["obj"][ok] - ref["res"]["obj"][ok]) / np.maximum(1, np.abs(ref["res"]["obj"][ok]))
        assert gap.max() <= 1e-6, (step, gap.max())
        assert np.abs(out["traj"][ok][..., :3] - ref["traj"][ok][..., :3]).max() <= 1e-3
        sw.advance(ref["traj"], ref["ctrl"], _ok(ref["res"]))
    pl.close()


def test_two_live_handles_of_different_size_and_changing_batch_sizes():
    """A small handle created while a large one is alive must not lower the kernel's shared-memory limit under it, and
    the cached dispatch order must not be reused across batch sizes whose pipeline chunks start elsewhere."""
    from oracle import c_oracle as co
    sw = sc.config5_random(seed=3, n_rob=1500, side=120.0)
    b = sw.make_batch()
    ref = co.solve_batch(b, max_nodes=64)
    sw.advance(ref["traj"], ref["ctrl"], _ok(ref["res"]))
    b = sw.make_batch()
    ref = co.solve_batch(b, max_nodes=64)
    big = TrajectoryPlanner(sw.params, max_agents=1500, max_neighbours=1500, max_nodes=64)
    small = TrajectoryPlanner(sw.params, max_agents=4, max_neighbours=2, max_nodes=64, rmax=18)
    small.solve_batch(b.take([0, 1]))          # configures the kernel for the small handle
    for n in (1500, 1499, 1500, 1200):         # large tiers of the big handle still launch; order cache keyed on (n, offset)
        os.environ["HDSM_CHUNKS"] = "2"
        try:
            out = big.solve_batch(b.take(np.arange(n)))
        finally:
            del os.environ["HDSM_CHUNKS"]
        assert np.array_equal(out["res"]["status"], ref["res"]["status"][:n]), n
        ok = ref["res"]["status"][:n] == OPTIMAL
        gap = np.abs(out["res"]["obj"][ok] - ref["res"]["obj"][:n][ok]) / np.maximum(1, np.abs(ref["res"]["obj"][:n][ok]))
        assert gap.max() <= 1e-6, (n, gap.max())
    big.close()
    small.close()


def test_search_rounds_depend_on_the_width_only():
    """search_width 1 / 2 / 4 / 8 against the C port run with the same width (status exact, objective 1e-6), and for each width
    the same outputs bit for bit whether a round's nodes run in one block (then in two passes) or on a cluster of 2, 4 or 8."""
    from oracle import c_oracle as co
    sw = sc.config5_random(seed=13, n_rob=400, side=60.0)
    for _ in range(3):
        b = sw.make_batch()
        r = co.solve_batch(b, max_nodes=64)
        sw.advance(r["traj"], r["ctrl"], _ok(r["res"]))
    b = sw.make_batch()
    assert co.solve_batch(b, max_nodes=64)["res"]["nodes"].max() >= 8      # the slice does branch
    for width in (1, 2, 4, 8):
        ref = co.solve_batch(b, max_nodes=64, width=width)
        outs = []
        for csize in (1, 2, 4, 8):
            if csize > width:
                continue
            os.environ["HDSM_CLUSTER"] = str(csize)
            try:
                pl = TrajectoryPlanner(sw.params, max_agents=b.n, max_neighbours=b.n, max_nodes=64, width=width)
                outs.append(pl.solve_batch(b))
                pl.close()
            finally:
                del os.environ["HDSM_CLUSTER"]
        out = outs[0]
        assert np.array_equal(out["res"]["status"], ref["res"]["status"]), width
        ok = ref["res"]["status"] == OPTIMAL
        gap = np.abs(out["res"]["obj"][ok] - ref["res"]["obj"][ok]) / np.maximum(1, np.abs(ref["res"]["obj"][ok]))
        assert gap.max() <= 1e-6, (width, gap.max())
        assert np.array_equal(out["res"]["nodes"], ref["res"]["nodes"]), width
        for o in outs[1:]:
            for k in ("traj", "ctrl", "assign", "poly_used"):
                assert np.array_equal(o[k], out[k]), (width, k)
            assert np.array_equal(o["res"]["obj"], out["res"]["obj"]) and np.array_equal(o["res"]["nodes"], out["res"]["nodes"])
    # the warm start of the search (previous plan's assignment solved in the second round): same optimum, same search as the port's
    ref = co.solve_batch(b, max_nodes=64, width=4, warm_start=True)
    cold = co.solve_batch(b, max_nodes=64, width=4)
    pl = TrajectoryPlanner(sw.params, max_agents=b.n, max_neighbours=b.n, max_nodes=64, width=4, warm_start=True)
    out = pl.solve_batch(b)
    pl.close()
    assert np.array_equal(out["res"]["status"], ref["res"]["status"]) and np.array_equal(out["res"]["nodes"], ref["res"]["nodes"])
    ok = (ref["res"]["status"] == OPTIMAL) & (cold["res"]["status"] == OPTIMAL)
    for other in (ref, cold):
        gap = np.abs(out["res"]["obj"][ok] - other["res"]["obj"][ok]) / np.maximum(1, np.abs(other["res"]["obj"][ok]))
        assert gap.max() <= 1e-6, gap.max()
    assert not np.array_equal(ref["res"]["nodes"], cold["res"]["nodes"])   # it does change the search


def test_single_warp_variant_gives_the_same_results():
    """HDSM_WARPS=1 selects the one-warp-per-agent instantiation of the kernel (same code, W = 1): identical statuses and
    node counts, objectives to 1e-9, for the depth-first search and for rounds of 4 with the warm start."""
    sw = sc.config5_random(seed=17, n_rob=200, side=45.0)
    from oracle import c_oracle as co
    for _ in range(2):
        b = sw.make_batch()
        r = co.solve_batch(b, max_nodes=64)
        sw.advance(r["traj"], r["ctrl"], _ok(r["res"]))
    b = sw.make_batch()
    for kw in (dict(width=1), dict(width=4, warm_start=True)):
        outs = []
        for warps in ("4", "1"):
            os.environ["HDSM_WARPS"] = warps
            try:
                pl = TrajectoryPlanner(sw.params, max_agents=b.n, max_neighbours=b.n, max_nodes=64, **kw)
                outs.append(pl.solve_batch(b))
                pl.close()
            finally:
                del os.environ["HDSM_WARPS"]
        a, c = outs
        assert np.array_equal(a["res"]["status"], c["res"]["status"]) and np.array_equal(a["res"]["nodes"], c["res"]["nodes"]), kw
        ok = a["res"]["status"] == OPTIMAL
        assert (np.abs(a["res"]["obj"][ok] - c["res"]["obj"][ok]) / np.maximum(1, np.abs(a["res"]["obj"][ok]))).max() <= 1e-9, kw


def test_head_start_and_dispatch_order_never_change_a_result():
    """From the second call on a large batch is solved in dispatch order (last call's iteration counts, longest first)
    and the head of that order skips the first pass: a cluster launch on a second stream (HDSM_EARLY_DIV, default
    n / 16).  Which launch solves an agent, and when, must not change a single bit of any output."""
    sw = sc.config5_random(seed=23, n_rob=1536, side=120.0)
    from oracle import c_oracle as co
    for _ in range(2):
        b = sw.make_batch()
        r = co.solve_batch(b, max_nodes=64, width=4)
        sw.advance(r["traj"], r["ctrl"], _ok(r["res"]))
    b = sw.make_batch()
    runs = {}
    for div in ("0", "16", "4"):
        os.environ["HDSM_EARLY_DIV"] = div
        try:
            pl = TrajectoryPlanner(sw.params, max_agents=b.n, max_neighbours=b.n, max_nodes=64, width=4)
            runs[div] = [pl.solve_batch(b) for _ in range(3)]   # call 1: natural order; calls 2, 3: ordered (+ head start)
            pl.close()
        finally:
            del os.environ["HDSM_EARLY_DIV"]
    base = runs["0"][0]
    assert (base["res"]["status"] == OPTIMAL).sum() > b.n // 2
    for div, outs in runs.items():
        for c, o in enumerate(outs):
            for k in ("traj", "ctrl", "assign", "poly_used"):
                assert np.array_equal(o[k], base[k]), (div, c, k)
            for f in ("status", "nodes", "iters", "obj", "kkt_res"):
                assert np.array_equal(o["res"][f], base["res"][f]), (div, c, f)


def test_host_entry_point_rejects_bad_indices():
    sw = sc.config2_circle(n_swarms=1)
    b = sw.make_batch()
    pl = TrajectoryPlanner(sw.params, max_agents=b.n, max_neighbours=10)
    from multi_agent_pkgs_b200.planner import HdsmError
    import copy
    for field, val in (("poly_rows", 19), ("nbr_end", 11), ("global_id", 10)):
        bb = copy.copy(b)
        a = getattr(b, field).copy()
        a.flat[0] = val
        setattr(bb, field, a)
        with pytest.raises(HdsmError):
            pl.solve_batch(bb)
    with pytest.raises(HdsmError):
        pl.solve_batch(b, assign_in=np.full((b.n, 10), 4, np.int32))
    pl.solve_batch(b)
    pl.close()


def test_nccl_exchange_equals_gloo_exchange_on_two_gpus():
    """Sharded closed loop over two GPUs (torchrun, NCCL group of positions + flags) against the same loop on one GPU:
    identical table checksums after every step.  Skips itself on a box with one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = os.path.join(ROOT, "tests", "closed_loop_worker.py")
    outs = []
    for world in (1, 2):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(29611 + world), script]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("SUMS ")][-1]
        outs.append(line)
    assert outs[0] == outs[1], outs
