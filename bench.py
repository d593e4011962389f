#!/usr/bin/env python
"""Benchmark of the replaced hot path: agent-QP solves per second at horizon N = 10.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (BASELINE.json configs[4], the north star's split): 4096 agents in a 200 x 200 m random forest,
closed loop, the swarm block-partitioned over the N GPUs (strong scaling: 4096 / N agents per GPU).  One "step" is
one replanning step of the whole swarm: per rank one solver launch on its shard (inter-agent planes from the
all-gathered neighbour table, exact assignment search, interior-point solves, position pack), the on-device
read-back / fallback / state advance, and ONE NCCL group that all-gathers the new plans' positions and validity
flags into the table every rank's NEXT step reads.  Reference trajectories and corridor cells of every step come from
the host producers of an untimed pre-roll of the same closed loop and are resident in HBM; everything else is live.

`value`   agent-QP solves/s, CUDA-event timed on the launching stream, max over ranks, L2 flushed between steps.
`e2e`     the same closed loop with every step's exogenous inputs uploaded from page-locked host memory and the
          step's results (trajectories, controls, statuses, assignments) downloaded, wall clock per step.
`weak`    (N > 1) 4096 agents PER GPU in a forest of N times the area, same measurement.
`config4` BASELINE.json configs[3]: the 256-agent circle, closed loop on one GPU, C port beside it.
`latency` what the unchanged ROS node sees: hdsm_solve_batch(n_local = 1), wall clock, p50 / p99.
`config2_replicas`  round 1's headline (configs[1] as 4096 independent 10-agent swarms), device-resident and through
          hdsm_solve_batch with page-locked host buffers.
`--impl reference`  the CPU arm: the reference's own solver is Gurobi 10 (closed source, absent), so this times the
          C port of the oracle (oracle/hdsm_oracle.c, kind "port") on all host cores on the same closed loop.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# The contract is ONE json line on stdout, but libraries write there too (NCCL prints its version banner to
# stdout at every debug level above NONE).  File descriptor 1 is therefore pointed at stderr for the whole
# run and the result line is written to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

from multi_agent_pkgs_b200 import scenarios as sc  # noqa: E402

SNAP_STEPS = (1, 6, 12, 18)
DISTINCT_SWARMS = 24
MAX_NODES = 64
WIDTH = 4          # nodes per round of the assignment search in the closed-loop workloads (hdsm_params.search_width)
WARM = True        # hdsm_params.warm_start in the same workloads (the previous plan's assignment is tried in the second round)
METRIC = "agent-QP solves/sec (horizon N=10)"
N_AGENTS = 4096


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------
def tile_batch(b, reps):
    """Replicate a batch of independent swarm instances `reps` times (neighbour ranges re-based)."""
    n, n_rob = b.n, b.all_pos.shape[0]
    idx = np.tile(np.arange(n), reps)
    off = np.repeat(np.arange(reps) * n_rob, n).astype(np.int32)
    t = b.take(idx)
    t.global_id = (t.global_id + off).astype(np.int32)
    t.nbr_begin = (t.nbr_begin + off).astype(np.int32)
    t.nbr_end = (t.nbr_end + off).astype(np.int32)
    t.all_pos = np.tile(b.all_pos, (reps, 1, 1))
    t.all_valid = np.tile(b.all_valid, reps)
    return t


def make_snapshots(solve, seed, n_swarms):
    """Closed-loop simulation of DISTINCT_SWARMS instances; returns the frozen input batches at
    SNAP_STEPS, each tiled up to n_swarms instances.  `solve(batch)` -> dict(traj, ctrl, res)."""
    sw = sc.config2_circle(seed=seed, n_swarms=min(DISTINCT_SWARMS, n_swarms))
    reps = -(-n_swarms // (sw.n // 10))
    snaps = []
    for step in range(max(SNAP_STEPS) + 1):
        b = sw.make_batch()
        if step in SNAP_STEPS:
            t = tile_batch(b, reps)
            snaps.append(t.take(np.arange(n_swarms * 10)) if t.n > n_swarms * 10 else t)
        out = solve(b)
        st = out["res"]["status"]
        sw.advance(out["traj"], out["ctrl"], (st == 0) | ((st == 4) & np.isfinite(out["res"]["obj"])))
    for s in snaps:  # take() keeps the full table; trim it to the instances actually used
        n_rob = int(s.nbr_end.max())
        s.all_pos, s.all_valid = s.all_pos[:n_rob], s.all_valid[:n_rob]
    return snaps


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], 0.0, set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arm (oracle port): bench.py may execute oracle/ only here
# ----------------------------------------------------------------------------------------------
def cpu_solve_rate(snaps, min_seconds, max_agents=None):
    from oracle import c_oracle as co
    cores = co.max_threads()
    n = t = 0
    i = 0
    while t < min_seconds:
        b = snaps[i % len(snaps)]
        if max_agents and b.n > max_agents:
            b = b.take(np.arange(max_agents))
        t0 = time.perf_counter()
        co.solve_batch(b, max_nodes=MAX_NODES, n_threads=cores)
        t += time.perf_counter() - t0
        n += b.n
        i += 1
    return n / t, cores, n, t


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import c_oracle as co
    co.build()
    n_swarms = args.swarms
    # bounded sample of the workload per step: the first 2000 swarm instances (20 000 agent QPs, the same
    # sample size as the cpu_baseline leg of the GPU arm) - enough agents per thread to amortise the few
    # expensive ones, about half a second per step on 16 threads
    snaps = make_snapshots(lambda b: co.solve_batch(b, max_nodes=MAX_NODES), args.seed, min(n_swarms, 2000))
    sample = snaps[0].n
    for _ in range(args.warmup):
        co.solve_batch(snaps[0].take(np.arange(sample)), max_nodes=MAX_NODES)
    times = []
    for k in range(args.steps):
        b = snaps[k % len(snaps)].take(np.arange(sample))
        t0 = time.perf_counter()
        co.solve_batch(b, max_nodes=MAX_NODES)
        times.append(time.perf_counter() - t0)
    total = float(np.sum(times))
    value = sample * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, world),
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": co.max_threads(), "kind": "port",
                             "sample": f"{sample} agent QPs per step (first instances of the workload, snapshots rotated), "
                                       f"C port of the oracle on all host threads, Gurobi not available"},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------------------------
# secondary measurement: the corridor producer of the path, reported under "corridor"
# ----------------------------------------------------------------------------------------------
def corridor_measure(n_agents=4096, steps=10, local_rank=0, cpu=True):
    """Throughput of hdsm_corridor_batch_device / hdsm_corridor_batch next to the CPU checkers.
    Workload: agents of the config-2 circle swap pulled into the forest, one 66 x 66 x 20 int8 local voxel
    grid per agent (1.4 GB for the default 16 384 agents = 9 waves of one-warp blocks), poly_hor 4, n_it_decomp 42; tiled from 240
    distinct agents.  L2 flushed between timed launches."""
    import torch
    from multi_agent_pkgs_b200 import corridor as cr
    sw = sc.config2_circle(n_swarms=DISTINCT_SWARMS)
    for i in range(sw.n):
        sw.state[i, :2] = sw.world.push_free(0.45 * sw.state[i, :2] + 0.55 * sw.goal[i, :2], 0.3)
    base = cr.corridor_batch(sw)
    reps = -(-n_agents // base.n)

    def tile(a):
        return np.ascontiguousarray(np.concatenate([a] * reps)[:n_agents])
    cb = cr.CorridorBatch(base.poly_hor, base.n_it, base.rmax, base.voxel, tile(base.grids), None, tile(base.dims),
                          tile(base.origins), tile(base.pos), tile(base.path), tile(base.n_path), tile(base.prev_traj))
    dev = torch.device(f"cuda:{local_rank}")
    gen = cr.SafeCorridorGenerator(cb.poly_hor, cb.n_it, cb.voxel, cb.n, cb.n, int(cb.grids[0].size), cb.prev_traj.shape[1],
                                   cb.path.shape[1], device=local_rank)
    db = cr.DeviceCorridorBatch(cb, dev)
    stream = torch.cuda.Stream(device=dev)  # a real stream: NULL would select the handle's own stream, unseen by the events
    torch.cuda.set_stream(stream)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        gen.generate_device(db.t, cb.n, stream.cuda_stream)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.fill_(k & 0xFF)
        ev[k][0].record(stream)
        gen.generate_device(db.t, cb.n, stream.cuda_stream)
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    rows = db.t["poly_rows"].cpu().numpy()
    npoly = int((rows > 0).sum())
    gen.generate(cb)
    t0 = time.perf_counter()
    out = gen.generate(cb)
    e2e = time.perf_counter() - t0
    # steady state: the reference keeps the polytopes the last optimisation used (:1252-1281) and only grows
    # the missing ones, so a replanning step in flight costs far less than the cold start timed above.  Five
    # closed-loop steps of the distinct agents (corridor and optimisation on the GPU) give a batch with
    # previous polytopes / used flags / plans; it is tiled and timed the same way.
    from multi_agent_pkgs_b200.planner import TrajectoryPlanner
    loop = cr.CorridorLoop(sw)
    g1 = cr.SafeCorridorGenerator(base.poly_hor, base.n_it, base.voxel, base.n, base.n, int(base.grids[0].size),
                                  base.prev_traj.shape[1], base.path.shape[1], device=local_rank)
    pl1 = TrajectoryPlanner(sw.params, max_agents=sw.n, max_neighbours=10, device=local_rank, max_nodes=MAX_NODES)
    for _ in range(5):
        o1 = g1.generate(loop.corridor_inputs())
        r1 = pl1.solve_batch(loop.solver_inputs(o1))
        st = r1["res"]["status"]
        loop.advance(o1, r1, (st == 0) | ((st == 4) & np.isfinite(r1["res"]["obj"])))
    sb = loop.corridor_inputs()
    o_sb = g1.generate(sb)
    b_sb = loop.solver_inputs(o_sb)  # the optimisation inputs of the same step (rows = this corridor's output)
    g1.close()
    pl1.close()
    cbs = cr.CorridorBatch(sb.poly_hor, sb.n_it, sb.rmax, sb.voxel, tile(sb.grids), None, tile(sb.dims), tile(sb.origins),
                           tile(sb.pos), tile(sb.path), tile(sb.n_path), tile(sb.prev_traj), tile(sb.prev_n), tile(sb.prev_A),
                           tile(sb.prev_b), tile(sb.prev_rows), tile(sb.prev_seeds), tile(sb.prev_used))
    dbs = cr.DeviceCorridorBatch(cbs, dev)
    for _ in range(3):
        gen.generate_device(dbs.t, cbs.n, stream.cuda_stream)
    torch.cuda.synchronize()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.fill_(k & 0xFF)
        ev2[k][0].record(stream)
        gen.generate_device(dbs.t, cbs.n, stream.cuda_stream)
        ev2[k][1].record(stream)
    torch.cuda.synchronize()
    ms_steady = float(np.mean([a.elapsed_time(b) for a, b in ev2]))
    # the chained replanning step on the device: corridor rows are written straight into the optimisation's
    # polytope inputs (same stream, no host round trip), then every agent is solved
    from multi_agent_pkgs_b200.swarm import DeviceBatch
    big = tile_batch(b_sb, -(-n_agents // b_sb.n))
    big = big.take(np.arange(n_agents)) if big.n > n_agents else big
    dsolve = DeviceBatch(big, dev)
    for k in ("poly_A", "poly_b", "poly_rows"):
        dsolve.t[k] = dbs.t[k]
    plc = TrajectoryPlanner(sw.params, max_agents=n_agents, max_neighbours=10, device=local_rank, max_nodes=MAX_NODES)
    for _ in range(3):
        gen.generate_device(dbs.t, cbs.n, stream.cuda_stream)
        plc.solve_batch_device(dsolve.t, dsolve.n_rob, stream.cuda_stream)
    torch.cuda.synchronize()
    ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.fill_(k & 0xFF)
        ev3[k][0].record(stream)
        gen.generate_device(dbs.t, cbs.n, stream.cuda_stream)
        plc.solve_batch_device(dsolve.t, dsolve.n_rob, stream.cuda_stream)
        ev3[k][1].record(stream)
    torch.cuda.synchronize()
    ms_chain = float(np.mean([a.elapsed_time(b) for a, b in ev3]))
    chain_res = dsolve.results()
    chain_status = np.bincount(chain_res["status"], minlength=6)
    chain_iters, chain_nodes = float(chain_res["iters"].mean()), float(chain_res["nodes"].mean())
    plc.close()
    kept = int(cbs.prev_n.sum()) if cbs.prev_n is not None else 0
    new_steady = int((dbs.t["poly_rows"].cpu().numpy() > 0).sum()) - 0
    launches = gen.launch_count
    smem = gen.smem_bytes
    gen.close()
    balg = cr.corridor_algorithmic_bytes(cb, rows)
    line = {"workload": f"{cb.n} agents, one 66x66x20 int8 local grid each, poly_hor {cb.poly_hor}, n_it_decomp {cb.n_it}; "
                        f"value = cold start (all {cb.poly_hor} polytopes grown), steady_state = a step in flight",
            "metric": "corridor updates/sec (agents/s)", "value": cb.n / (ms * 1e-3), "kernel_ms": ms,
            "polytopes_per_s": npoly / (ms * 1e-3), "dtype": "int8 grid -> f64 rows",
            "e2e": {"value": cb.n / e2e, "unit": "agents/s", "h2d_bytes_per_step": cb.input_bytes(),
                    "d2h_bytes_per_step": int(sum(out[k].nbytes for k in out))},
            "steady_state": {"value": cbs.n / (ms_steady * 1e-3), "unit": "agents/s", "kernel_ms": ms_steady,
                             "polytopes_out": new_steady, "polytopes_in": kept,
                             "note": "closed-loop step 5: previous polytopes, used flags and plans supplied; "
                                     "only the missing polytopes are grown"},
            "chained_step": {"value": cbs.n / (ms_chain * 1e-3), "unit": "agents/s", "ms": ms_chain,
                             "status_counts": [int(v) for v in chain_status], "ipm_iters_per_agent": chain_iters,
                             "qp_relaxations_per_agent": chain_nodes,
                             "note": "steady-state corridor kernel + optimisation kernels on one stream, corridor rows "
                                     "consumed in place (no host round trip); status_counts = optimal, infeasible, "
                                     "max_iter, numerical, node_limit (64 relaxations), row_overflow - cells grown from "
                                     "voxels overlap much more than the synthetic cells of the headline workload, so the "
                                     "assignment search is deeper here"},
            "gpu_launches": int(launches), "algorithmic_bytes_per_agent": balg, "smem_bytes_per_block": smem,
            "squeezed_seed_agents": int((out["flags"] & 1 != 0).sum())}
    if cpu:
        from oracle import corridor as oc
        t0 = time.perf_counter()
        ref = oc.c_safe_corridor(cb)
        t_cpu = time.perf_counter() - t0
        line["bit_exact_vs_cpu_port"] = bool(all(np.array_equal(out[k], ref[k]) for k in out))
        line["cpu_baseline"] = {"value": cb.n / t_cpu, "unit": "agents/s", "cores": oc.max_threads(), "kind": "port",
                                "sample": f"all {cb.n} agents once, C restatement on all host threads"}
        if oc.have_ref():  # the reference's own GetPolyOcta3D (compiled unmodified), one thread, same seeds
            t_ref, cnt = 0.0, 0
            for i in range(min(base.n, 120)):
                g = base.grids[i].copy()
                g[g == -1] = 100
                for p in range(base.poly_hor):
                    if ref["poly_rows"][i, p] == 0:
                        continue
                    sv = np.round((ref["seeds"][i, p] - base.origins[i]) / base.voxel - 0.5).astype(np.int32)
                    t0 = time.perf_counter()
                    oc.ref_poly(g, sv, base.n_it, base.voxel, -(p + 1), base.origins[i])
                    t_ref += time.perf_counter() - t0
                    cnt += 1
            line["cpu_reference"] = {"value": cnt / t_ref, "unit": "polytopes/s", "cores": 1, "kind": "reference",
                                     "sample": f"{cnt} GetPolyOcta3D calls of oracle/_ref (the reference's own code) on the same seeds"}
    return line, balg * cb.n / (ms * 1e-3) / 1e9


def reftraj_measure(n_agents=16384, steps=10, local_rank=0):
    """Secondary measurement of the reference-trajectory producer: hdsm_reftraj_batch_device on
    agents of the config-2 swarms inside the forest (66 x 66 x 20 grids with a potential field, 10 neighbours each
    with plans), tiled from 240 distinct agents; the C port of the checker on all host threads beside it."""
    import torch
    from multi_agent_pkgs_b200 import reftraj as rtj
    sw = sc.config2_circle(n_swarms=DISTINCT_SWARMS)
    for i in range(sw.n):
        sw.state[i, :2] = sw.world.push_free(0.45 * sw.state[i, :2] + 0.55 * sw.goal[i, :2], 0.3)
    base = rtj.reftraj_batch(sw)
    base.all_valid[:] = 1
    reps = -(-n_agents // base.n)
    n_rob = base.all_pos.shape[0]
    off = np.repeat(np.arange(reps, dtype=np.int32) * n_rob, base.n)[:n_agents]

    def tile(a):
        return np.ascontiguousarray(np.concatenate([a] * reps)[:n_agents])
    rb = rtj.RefTrajBatch(base.n_hor, base.dt, base.voxel, tile(base.grids), None, tile(base.dims), tile(base.origins),
                          tile(base.path), tile(base.n_path), tile(base.prev_ref), tile(base.have_prev), tile(base.increment),
                          tile(base.traj), tile(base.global_id) + off, tile(base.nbr_begin) + off, tile(base.nbr_end) + off,
                          np.ascontiguousarray(np.concatenate([base.all_pos] * reps)), np.tile(base.all_valid, reps))
    dev = torch.device(f"cuda:{local_rank}")
    gen = rtj.ReferenceTrajectoryGenerator(rb, device=local_rank)
    t = {}
    for k in ("grids", "dims", "origins", "path", "n_path", "prev_ref", "have_prev", "increment", "traj", "global_id", "nbr_begin",
              "nbr_end", "all_pos", "all_valid"):
        a = getattr(rb, k)
        t[k] = torch.from_numpy(np.ascontiguousarray(a.reshape(a.shape[0], -1) if k == "grids" else a)).to(dev)
    N1 = rb.n_hor + 1
    t["ref"] = torch.zeros((rb.n, N1, 6), dtype=torch.float64, device=dev)
    t["ref_solver"] = torch.zeros((rb.n, rb.n_hor, 6), dtype=torch.float64, device=dev)
    t["path_vel"] = torch.zeros(rb.n, dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        gen.generate_device(t, rb.n, rb.all_pos.shape[0], stream.cuda_stream)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.fill_(k & 0xFF)
        ev[k][0].record(stream)
        gen.generate_device(t, rb.n, rb.all_pos.shape[0], stream.cuda_stream)
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    got_ref, got_vel = t["ref"].cpu().numpy(), t["path_vel"].cpu().numpy()
    solver_rows_equal = bool(np.array_equal(t["ref_solver"].cpu().numpy(), got_ref[:, :rb.n_hor]))
    gen.close()
    from oracle import reftraj as ort
    t0 = time.perf_counter()
    want = ort.c_generate(rb)
    t_cpu = time.perf_counter() - t0
    # compulsory traffic per agent: path, previous reference, own plan, neighbours' plans, outputs (the voxels a
    # ray touches are a few hundred bytes out of an 87 KB grid and are left out)
    balg = float(np.mean(rb.n_path * 24 + N1 * 24 + N1 * 24 + (rb.nbr_end - rb.nbr_begin) * N1 * 24 + N1 * 48 + rb.n_hor * 48 + 8))
    return {"workload": f"{rb.n} agents, 66x66x20 grids with potential field, {int((rb.nbr_end - rb.nbr_begin)[0])} neighbour plans each",
            "metric": "reference trajectories/sec (agents/s)", "value": rb.n / (ms * 1e-3), "kernel_ms": ms, "dtype": "f64",
            "max_rel_path_vel_diff_vs_cpu_port": float(np.abs(got_vel / want["path_vel"] - 1).max()),
            "max_abs_ref_diff_vs_cpu_port": float(np.abs(got_ref - want["ref"]).max()), "solver_layout_rows_equal": solver_rows_equal,
            "algorithmic_bytes_per_agent": balg,
            "cpu_baseline": {"value": rb.n / t_cpu, "unit": "agents/s", "cores": ort.max_threads(), "kind": "port",
                             "sample": f"all {rb.n} agents once, C restatement on all host threads"}}


def map_measure(n_grids=4096, steps=10, local_rank=0):
    """Secondary measurement of the local-map post-processing: hdsm_map_batch_device on raw
    66 x 66 x 20 forest grids (tiled from 240 distinct ones; 357 MB in + 357 MB out per launch for 4096 grids, larger
    than L2), the C port on all host threads beside it.  Algorithmic bytes = one read and one write per voxel."""
    import torch
    from multi_agent_pkgs_b200 import mapping as mp
    sw = sc.config2_circle(n_swarms=DISTINCT_SWARMS)
    for i in range(sw.n):
        sw.state[i, :2] = sw.world.push_free(0.45 * sw.state[i, :2] + 0.55 * sw.goal[i, :2], 0.3)
    base = np.stack([mp.raw_local_grid(sw.world, sw.state[i, :3])[0] for i in range(sw.n)])
    reps = -(-n_grids // base.shape[0])
    grids = np.ascontiguousarray(np.concatenate([base] * reps)[:n_grids])
    dev = torch.device(f"cuda:{local_rank}")
    gen = mp.MapProcessor(0.3, n_grids, grids[0].size, device=local_rank)
    t_in = torch.from_numpy(grids.reshape(n_grids, -1)).to(dev)
    t_out = torch.empty_like(t_in)
    t_dims = torch.from_numpy(np.tile(np.array([grids.shape[3], grids.shape[2], grids.shape[1]], np.int32), (n_grids, 1))).to(dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        gen.process_device(t_in, t_dims, t_out, stream.cuda_stream)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.fill_(k & 0xFF)
        ev[k][0].record(stream)
        gen.process_device(t_in, t_dims, t_out, stream.cuda_stream)
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    got = t_out.cpu().numpy().reshape(grids.shape)
    gen.close()
    from oracle import mapping as om
    m = min(n_grids, 960)  # bounded CPU sample
    t0 = time.perf_counter()
    want = om.c_process(grids[:m], 0.3, 0.3, 1.5, 4)
    t_cpu = time.perf_counter() - t0
    balg = 2.0 * grids[0].size
    return {"workload": f"{n_grids} raw 66x66x20 int8 grids; SetUncertainToUnknown, InflateObstacles 0.3 m, CreatePotentialField 1.5 m / pow 4",
            "metric": "grids/sec", "value": n_grids / (ms * 1e-3), "kernel_ms": ms, "dtype": "int8",
            "byte_exact_vs_cpu_port": bool(np.array_equal(got[:m], want)), "algorithmic_bytes_per_grid": balg,
            "cpu_baseline": {"value": m / t_cpu, "unit": "grids/s", "cores": om.max_threads(), "kind": "port",
                             "sample": f"the first {m} grids once, C restatement on all host threads"}}


def sense_measure(n_agents=4096, steps=10, local_rank=0, cpu_agents=512):
    """Secondary measurement of the local-map acquisition (first half): hdsm_sense_batch_device, one
    steady-state update (kept grids merged in) of n_agents agents spread over one 120 x 120 m forest environment grid
    (400 x 400 x 20 voxels of 0.3 m, shared by all agents), 66 x 66 x 20 local grids, 360 degree ray casting (13 992 rays
    per agent).  Algorithmic bytes = the kept grid read once and the new grid written once per agent."""
    import torch
    from multi_agent_pkgs_b200 import sensing as sn
    rng = np.random.default_rng(4)
    vox, rng3 = 0.3, (20.0, 20.0, 6.0)
    world = sc.Forest.density(rng, (0.0, 0.0), (114.0, 114.0), 0.2)
    env, org = sn.environment_grid(world, vox)
    pos0 = np.stack([rng.uniform(5, 109, n_agents), rng.uniform(5, 109, n_agents), rng.uniform(1.0, 3.0, n_agents)], 1)
    pos1 = pos0 + rng.uniform(-0.6, 0.6, pos0.shape) * [1, 1, 0.1]
    dev = torch.device(f"cuda:{local_rank}")
    mb = sn.LocalMapBuilder(vox, n_agents, rng3, device=local_rank)
    cells = mb.grid_stride
    dim_env = (env.shape[2], env.shape[1], env.shape[0])
    t_env = torch.from_numpy(env.reshape(-1)).to(dev)
    t_p0, t_p1 = torch.from_numpy(pos0).to(dev), torch.from_numpy(pos1).to(dev)
    t_g0 = torch.empty((n_agents, cells), dtype=torch.int8, device=dev)
    t_g1 = torch.empty_like(t_g0)
    t_o0 = torch.empty((n_agents, 3), dtype=torch.float64, device=dev)
    t_o1 = torch.empty_like(t_o0)
    t_have = torch.ones(n_agents, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    ev0 = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    ev0[0].record(stream)
    mb.update_device(t_env, dim_env, org, t_p0, None, None, None, None, t_g0, t_o0, stream.cuda_stream)  # first update
    ev0[1].record(stream)
    for _ in range(3):
        mb.update_device(t_env, dim_env, org, t_p1, None, t_g0, t_o0, t_have, t_g1, t_o1, stream.cuda_stream)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.fill_(k & 0xFF)
        ev[k][0].record(stream)
        mb.update_device(t_env, dim_env, org, t_p1, None, t_g0, t_o0, t_have, t_g1, t_o1, stream.cuda_stream)
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    first_ms = float(ev0[0].elapsed_time(ev0[1]))
    m = min(n_agents, cpu_agents)  # bounded CPU sample
    got0 = t_g0[:m].cpu().numpy()
    got1, goto1 = t_g1[:m].cpu().numpy(), t_o1[:m].cpu().numpy()
    launches = mb.launch_count
    mb.close()
    from oracle import sensing as osn
    want0, wo0 = osn.c_update(env, org, pos0[:m], vox, rng3)
    t0 = time.perf_counter()
    want1, wo1 = osn.c_update(env, org, pos1[:m], vox, rng3, old_grids=want0, old_origin=wo0)
    t_cpu = time.perf_counter() - t0
    exact = bool(np.array_equal(got0.reshape(want0.shape), want0) and np.array_equal(got1.reshape(want1.shape), want1)
                 and np.array_equal(goto1, wo1))
    ref_node = None
    if osn.have_ref():  # the reference's OWN node (map_builder.cpp compiled unmodified), one core, its own timers
        tms, acc = np.zeros(3), np.zeros(3)
        k = min(m, 6)
        for a in range(k):
            osn.ref_update(env, org, pos1[a], vox, rng3, old_grid=want0[a], old_origin=wo0[a], times_ms=tms)
            acc += tms
        ref_node = {"kind": "reference", "cores": 1, "raycast_ms": float(acc[0] / k), "merge_ms": float(acc[1] / k), "callback_total_ms": float(acc[2] / k),
                    "value": float(1e3 * k / (acc[0] + acc[1])), "unit": "agents/s",
                    "sample": f"{k} agents, MapBuilder::EnvironmentVoxelGridCallback of the compiled reference node; value = 1 / (its ray-cast + "
                              f"merge timers); callback_total_ms also holds its grid post-processing"}
    balg = 2.0 * cells
    return {"reference_node": ref_node,
            "workload": f"{n_agents} agents in one {dim_env[0]}x{dim_env[1]}x{dim_env[2]} forest environment grid (0.3 m voxels), "
                        f"66x66x20 local grids, 360 degree ray casting (13 992 rays per agent), merge with the kept grids",
            "metric": "map updates/sec (agents/s)", "value": n_agents / (ms * 1e-3), "kernel_ms": ms, "first_update_ms": first_ms,
            "dtype": "int8 grids, f64 ray traversal", "byte_exact_vs_cpu_port": exact, "algorithmic_bytes_per_agent": balg,
            "gpu_launches": int(launches),
            "cpu_baseline": {"value": m / t_cpu, "unit": "agents/s", "cores": osn.max_threads(), "kind": "port",
                             "sample": f"the first {m} agents once (steady-state update), C restatement on all host threads"}}


def full_step_measure(n_agents=2048, steps=10, local_rank=0):
    """The complete replanning step of the reference's agent and map-builder nodes on the device, one stream, every stage
    reading its predecessor's output in place (TrajPlanningIteration agent_class.cpp:157-258 behind EnvironmentVoxelGridCallback
    map_builder.cpp:80-240): local-map acquisition -> post-processing -> safe corridor -> reference trajectory -> optimisation ->
    read-back / advance.  Only the agent poses are uploaded per step; paths come from the (out-of-scope) host path planner and
    stay resident.  n_agents agents in one 120 x 120 m forest environment grid, first planning step of every agent (corridor grown
    from scratch, no previous plan), every other agent a static neighbour (planes and velocity sweep active)."""
    import torch
    from multi_agent_pkgs_b200 import sensing as sn, mapping as mp, corridor as cr, reftraj as rtj
    from multi_agent_pkgs_b200.planner import TrajectoryPlanner
    from multi_agent_pkgs_b200._lib import RESULT_DTYPE
    from oracle import c_oracle as co
    rng = np.random.default_rng(7)
    vox, rng3, N, P, R = 0.3, (20.0, 20.0, 6.0), 10, 4, 18
    params = sc.agile_params(N)
    world = sc.Forest.density(rng, (0.0, 0.0), (114.0, 114.0), 0.2)
    env, org = sn.environment_grid(world, vox)
    n = n_agents
    pos0 = np.stack([rng.uniform(12, 102, n), rng.uniform(12, 102, n), rng.uniform(1.2, 2.8, n)], 1)
    for i in range(n):
        pos0[i, :2] = world.push_free(pos0[i, :2], 0.35)
    ang = rng.uniform(0, 2 * np.pi, n)
    goal = pos0 + np.stack([30 * np.cos(ang), 30 * np.sin(ang), np.zeros(n)], 1)
    pos1 = pos0.copy()
    path, n_path = np.zeros((n, 16, 3)), np.zeros(n, np.int32)
    for i in range(n):
        pts = np.vstack([pos1[i][None], cr.clipped_path(pos1[i], goal[i], world)])[:16]
        path[i, :len(pts)], n_path[i] = pts, len(pts)
    dev = torch.device(f"cuda:{local_rank}")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    mb = sn.LocalMapBuilder(vox, n, rng3, device=local_rank)
    cells = mb.grid_stride
    gd = sn.grid_dims(vox, rng3)
    mproc = mp.MapProcessor(vox, n, cells, device=local_rank)
    gen = cr.SafeCorridorGenerator(P, 42, vox, n, n, cells, N + 1, 16, device=local_rank)
    rb0 = rtj.RefTrajBatch(N, 0.1, vox, np.zeros((1, gd[2], gd[1], gd[0]), np.int8), None, np.zeros((n, 3), np.int32), np.zeros((n, 3)), path, n_path,
                           np.zeros((n, N + 1, 3)), np.zeros(n, np.uint8), np.ones(n, np.uint8), np.zeros((n, N + 1, 3)), np.arange(n, dtype=np.int32),
                           np.zeros(n, np.int32), np.full(n, n, np.int32), np.zeros((n, N + 1, 3)), np.ones(n, np.uint8))
    rgen = rtj.ReferenceTrajectoryGenerator(rb0, max_agents=n, max_grids=n, device=local_rank)
    pl = TrajectoryPlanner(params, n, n, local_rank, max_nodes=MAX_NODES, width=WIDTH, warm_start=WARM)
    f64, T = torch.float64, torch.from_numpy
    dim_env = (env.shape[2], env.shape[1], env.shape[0])
    t_env = T(env.reshape(-1)).to(dev)
    h_pos = T(pos1.copy()).pin_memory()
    t_pos = torch.empty((n, 3), dtype=f64, device=dev)
    g_old, g_new, g_map = (torch.empty((n, cells), dtype=torch.int8, device=dev) for _ in range(3))
    o_old, o_new = torch.empty((n, 3), dtype=f64, device=dev), torch.empty((n, 3), dtype=f64, device=dev)
    have_old = torch.ones(n, dtype=torch.uint8, device=dev)
    dims = T(np.tile(np.array(gd, np.int32), (n, 1))).to(dev)
    plan = torch.empty((n, N + 1, 3), dtype=f64, device=dev)            # every agent's "plan": its position repeated
    ones = torch.ones(n, dtype=torch.uint8, device=dev)
    cor = {"grids": g_map, "dims": dims, "origins": o_new, "pos": t_pos, "path": T(path).to(dev), "n_path": T(n_path).to(dev),
           "poly_A": torch.zeros((n, P, R, 3), dtype=f64, device=dev), "poly_b": torch.zeros((n, P, R), dtype=f64, device=dev),
           "poly_rows": torch.zeros((n, P), dtype=torch.int32, device=dev), "seeds": torch.zeros((n, P, 3), dtype=f64, device=dev),
           "flags": torch.zeros(n, dtype=torch.int32, device=dev)}
    ids = torch.arange(n, dtype=torch.int32, device=dev)
    zeros_i, full_i = torch.zeros(n, dtype=torch.int32, device=dev), torch.full((n,), n, dtype=torch.int32, device=dev)
    ref = {"grids": g_map, "dims": dims, "origins": o_new, "path": cor["path"], "n_path": cor["n_path"],
           "prev_ref": torch.zeros((n, N + 1, 3), dtype=f64, device=dev), "have_prev": torch.zeros(n, dtype=torch.uint8, device=dev),
           "increment": ones, "traj": plan, "global_id": ids, "nbr_begin": zeros_i, "nbr_end": full_i, "all_pos": plan, "all_valid": ones,
           "ref": torch.zeros((n, N + 1, 6), dtype=f64, device=dev), "ref_solver": torch.zeros((n, N, 6), dtype=f64, device=dev),
           "path_vel": torch.zeros(n, dtype=f64, device=dev)}
    sol = {"global_id": ids, "nbr_begin": zeros_i, "nbr_end": full_i, "x0": torch.zeros((n, 9), dtype=f64, device=dev), "ref": ref["ref_solver"],
           "poly_A": cor["poly_A"], "poly_b": cor["poly_b"], "poly_rows": cor["poly_rows"], "prev_self_pos": plan, "all_pos": plan, "all_valid": ones,
           "traj": torch.zeros((n, N + 1, 9), dtype=f64, device=dev), "ctrl": torch.zeros((n, N, 3), dtype=f64, device=dev),
           "poly_used": torch.zeros((n, P), dtype=torch.uint8, device=dev), "assign_out": torch.zeros((n, N), dtype=torch.int32, device=dev),
           "res": torch.zeros((n, 32), dtype=torch.uint8, device=dev), "pos_out": torch.zeros((n, N + 1, 3), dtype=f64, device=dev),
           "traj_curr": torch.zeros((n, N + 1, 9), dtype=f64, device=dev), "ctrl_curr": torch.zeros((n, N, 3), dtype=f64, device=dev),
           "have_plan": torch.zeros(n, dtype=torch.uint8, device=dev)}
    h_traj = torch.empty((n, N + 1, 9), dtype=f64).pin_memory()
    h_res = torch.empty((n, 32), dtype=torch.uint8).pin_memory()
    # the kept grids of the previous map update (taken a little earlier on the way)
    t_prev = T(pos0 - 0.4 * (goal - pos0) / 30.0).to(dev)
    mb.update_device(t_env, dim_env, org, t_prev, None, None, None, None, g_old, o_old, sp)
    names = ("upload", "acquisition", "post-processing", "corridor", "reference", "optimisation", "advance")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]

    def step(record=False):
        if record:
            ev[0].record(stream)
        t_pos.copy_(h_pos, non_blocking=True)
        plan.copy_(t_pos[:, None, :].expand(-1, N + 1, -1))
        sol["x0"].zero_()
        sol["x0"][:, :3].copy_(t_pos)
        sol["have_plan"].zero_()
        if record:
            ev[1].record(stream)
        mb.update_device(t_env, dim_env, org, t_pos, None, g_old, o_old, have_old, g_new, o_new, sp)
        if record:
            ev[2].record(stream)
        mproc.process_device(g_new, dims, g_map, sp)
        if record:
            ev[3].record(stream)
        gen.generate_device(cor, n, sp)
        if record:
            ev[4].record(stream)
        rgen.generate_device(ref, n, n, sp)
        if record:
            ev[5].record(stream)
        pl.solve_batch_device(sol, n, sp)
        if record:
            ev[6].record(stream)
        prev = sol.pop("prev_self_pos")
        pl.advance_device(sol, sp)
        sol["prev_self_pos"] = prev
        if record:
            ev[7].record(stream)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    per = np.zeros(len(names))
    for _ in range(steps):
        step(record=True)
        torch.cuda.synchronize()
        per += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))])
    per /= steps
    t_wall = 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        h_traj.copy_(sol["traj_curr"], non_blocking=True)
        h_res.copy_(sol["res"], non_blocking=True)
        stream.synchronize()
        t_wall += time.perf_counter() - t0
    res = np.frombuffer(h_res.numpy().tobytes(), dtype=RESULT_DTYPE)
    # the optimisation of the first 256 agents against the C port on the rows / references the device stages produced
    m = min(n, 256)
    hb = sc.Batch(params, np.arange(m, dtype=np.int32), np.zeros(m, np.int32), np.full(m, n, np.int32), np.c_[pos1[:m], np.zeros((m, 6))],
                  ref["ref_solver"][:m].cpu().numpy(), cor["poly_A"][:m].cpu().numpy(), cor["poly_b"][:m].cpu().numpy(), cor["poly_rows"][:m].cpu().numpy(),
                  np.repeat(pos1[:m, None, :], N + 1, 1), np.repeat(pos1[:, None, :], N + 1, 1), np.ones(n, np.uint8), R)
    want = co.solve_batch(hb, max_nodes=MAX_NODES, width=WIDTH, warm_start=WARM)["res"]
    both = (res["status"][:m] == 0) & (want["status"] == 0)
    gap = float((np.abs(res["obj"][:m][both] - want["obj"][both]) / np.maximum(1.0, np.abs(want["obj"][both]))).max()) if both.any() else 0.0
    launches = mb.launch_count + mproc.launch_count + gen.launch_count + rgen.launch_count + pl.launch_count
    for h in (mb, mproc, gen, rgen, pl):
        h.close()
    total = float(per.sum())
    return {"workload": f"{n} agents in one {dim_env[0]}x{dim_env[1]}x{dim_env[2]} forest environment grid, 66x66x20 local grids; acquisition (13 992 rays), "
                        f"post-processing, corridor (4 polytopes grown), reference trajectory, optimisation ({n} neighbour candidates), read-back",
            "metric": "full replanning steps/sec (agents/s)", "value": n / (total * 1e-3), "unit": "agents/s", "ms_per_step": total,
            "stage_ms": {k: float(v) for k, v in zip(names, per)},
            "e2e": {"value": n * steps / t_wall, "unit": "agents/s", "h2d_bytes_per_step": int(n * 24), "d2h_bytes_per_step": int(n * ((N + 1) * 72 + 32)),
                    "note": "wall clock per step: agent poses up from page-locked memory, all stages, trajectories and statuses down"},
            "status_counts": [int(v) for v in np.bincount(res["status"], minlength=6)[:6]],
            "optimisation_vs_c_port": {"agents": int(m), "status_mismatches": int((res["status"][:m] != want["status"]).sum()), "max_rel_obj_gap": gap},
            "gpu_launches_total": int(launches)}


def config_dict(args, world):
    return {"workload": f"config5: {args.agents} agents, random forest {SIDE:.0f} x {SIDE:.0f} m (0.2 columns/m^2), N=10, "
                        f"closed loop, sharded {args.agents // world} agents per GPU over {world} GPU(s), NCCL all-gather of plan "
                        f"positions consumed as the next step's neighbour table",
            "n_hor": 10, "poly_hor": 4, "n_agents": args.agents, "agents_per_gpu": args.agents // world,
            "neighbour_candidates_per_agent": args.agents, "max_nodes": MAX_NODES, "search_width": WIDTH, "warm_start": WARM, "seed": args.seed,
            "l2": "flushed between timed steps (512 MiB write)",
            "inputs": "ref / corridor cells of every step from the host producers of an untimed pre-roll of the same closed "
                      "loop, resident in HBM; x0, previous plans and the neighbour table advance on the device",
            "parallelism": f"agents block-partitioned by id over {world} GPU(s) (strong scaling), one NCCL group per step"}


SIDE = 200.0


def okmask(res):
    return (res["status"] == 0) | ((res["status"] == 4) & np.isfinite(res["obj"]))


# ----------------------------------------------------------------------------------------------
# CPU arm: the same closed loop with the C port of the oracle on all host threads
# ----------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import c_oracle as co
    co.build()
    sw = sc.config5_random(seed=args.seed, n_rob=args.agents, side=SIDE)
    pool = sc.InputPool(sw)
    cores = co.max_threads()
    times = []
    for s in range(args.warmup + args.steps):
        b = sw.make_batch_pooled(pool)
        t0 = time.perf_counter()
        out = co.solve_batch(b, max_nodes=MAX_NODES, width=WIDTH, warm_start=WARM)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
        sw.advance(out["traj"], out["ctrl"], okmask(out["res"]))
    pool.close()
    total = float(np.sum(times))
    value = args.agents * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, world),
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port",
                             "sample": f"all {args.agents} agents of every closed-loop step ({args.steps} timed steps after "
                                       f"{args.warmup}), C port of the oracle on all host threads; Gurobi not available"},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def timed_replay(loop, W, T, flush, dist, torch):
    """Device-timed closed loop: steps [0, W) untimed, steps [W, T) each between two events on the launching stream,
    L2 flushed before each.  Returns per-step ms (this rank)."""
    loop.reset()
    with torch.cuda.stream(loop.stream):
        for s in range(W):
            loop.device_step(s)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(T - W)]
    with torch.cuda.stream(loop.stream):
        for s in range(W, T):
            flush.fill_(s & 0xFF)
            ev[s - W][0].record(loop.stream)
            loop.device_step(s)
            ev[s - W][1].record(loop.stream)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    return np.array([a.elapsed_time(b) for a, b in ev])


def e2e_replay(loop, W, T, dist, torch):
    """Wall-clock closed loop with host buffers: every step uploads its exogenous inputs from page-locked memory and
    downloads its results.  Returns (seconds for steps [W, T), h2d bytes, d2h bytes per step)."""
    dev = loop.device
    stage = {k: torch.empty_like(v) for k, v in loop.rec[0].items()}
    outs = ("traj", "ctrl", "res", "assign_out", "poly_used")
    host = {k: torch.empty(loop.t[k].shape, dtype=loop.t[k].dtype).pin_memory() for k in outs}
    host_table = torch.empty(loop.current_table().shape, dtype=torch.float64).pin_memory()

    def step(s):
        with torch.cuda.stream(loop.stream):
            for k, v in loop.rec_host[s].items():
                stage[k].copy_(v, non_blocking=True)
            loop.device_step(s, inputs=stage)
            for k in outs:
                host[k].copy_(loop.t[k], non_blocking=True)
            host_table.copy_(loop.current_table(), non_blocking=True)
        loop.stream.synchronize()

    loop.reset()
    for s in range(W):
        step(s)
    if dist:
        dist.barrier()
    t = 0.0
    for s in range(W, T):
        t0 = time.perf_counter()
        step(s)
        t += time.perf_counter() - t0
    h2d = loop.input_bytes_per_step()
    d2h = int(sum(v.numel() * v.element_size() for v in host.values()) + host_table.numel() * 8)
    return t, h2d, d2h


def max_over_ranks(x, dist, torch, dev):
    if not dist:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def quality_of(stats, W):
    st = np.zeros(6, int)
    kkt, iters, nodes, mx = 0.0, 0, 0, 0
    per_agent = []
    for r in stats[W:]:
        st += np.bincount(r["status"], minlength=6)[:6]
        good = r["status"] == 0
        kkt = max(kkt, float(r["kkt_res"][good].max()) if good.any() else 0.0)
        iters += int(r["iters"].sum())
        nodes += int(r["nodes"].sum())
        per_agent.append(r["iters"])
    it = np.concatenate(per_agent) if per_agent else np.zeros(1)
    tot = max(1, int(st.sum()))
    return {"status_counts": {k: int(v) for k, v in zip(("optimal", "infeasible", "max_iter", "numerical", "node_limit", "row_overflow"), st)},
            "max_kkt_residual": kkt, "ipm_iters_per_solve": iters / tot, "qp_relaxations_per_solve": nodes / tot,
            "ipm_iters_p50_p99_max": [float(np.percentile(it, 50)), float(np.percentile(it, 99)), int(it.max())]}


def config4_measure(args, local_rank, flush, torch):
    """BASELINE.json configs[3]: 256 agents on a 60 m circle, forest, closed loop on one GPU; the C port on all host
    threads beside it on the same per-step inputs; every agent of every step compared with the port."""
    from multi_agent_pkgs_b200.swarm import ClosedLoop
    from oracle import c_oracle as co
    sw = sc.config4_circle256()
    W, T = 3, 3 + args.steps
    loop = ClosedLoop(sw, 1, 0, f"cuda:{local_rank}", MAX_NODES, None, width=WIDTH, warm_start=WARM)
    loop.checker = lambda b: co.solve_batch(b, max_nodes=MAX_NODES, width=WIDTH, warm_start=WARM)
    loop.keep_steps, loop.keep_agents = set(range(W, T)), sw.n
    par = loop.preroll(T, parity_sample=sw.n)
    ms = timed_replay(loop, W, T, flush, None, torch)
    same = loop.sums[T - 1] == __import__("multi_agent_pkgs_b200.swarm", fromlist=["table_checksum"]).table_checksum(loop.current_table())
    t_cpu = 0.0
    for s in range(W, T):
        t0 = time.perf_counter()
        co.solve_batch(loop.host_batches[s], max_nodes=MAX_NODES, width=WIDTH, warm_start=WARM)
        t_cpu += time.perf_counter() - t0
    q = quality_of(loop.stats, W)
    loop.close()
    return {"workload": "config4: 256 agents, circle radius 60 m, forest 0.2 columns/m^2, N=10, closed loop, 256 neighbour candidates per agent, 1 GPU",
            "ms_per_step": float(ms.mean()), "ms_per_step_max": float(ms.max()), "value": sw.n / (ms.mean() * 1e-3), "unit": "solves/s",
            "steps": int(T - W), "replay_equals_preroll": bool(same), "parity_vs_c_port": par, "quality": q,
            "cpu_baseline": {"ms_per_step": 1e3 * t_cpu / (T - W), "value": sw.n * (T - W) / t_cpu, "unit": "solves/s",
                             "cores": co.max_threads(), "kind": "port", "sample": "the same 256 agents of the same timed steps"}}


def latency_measure(local_rank):
    """Single-call latency of the drop-in as the unchanged ROS node would see it (agent_class.cpp:137: 10 Hz, :952: 80 ms
    cap): hdsm_solve_batch with n_local = 1 and ordinary (pageable) host buffers, wall clock around the call."""
    from multi_agent_pkgs_b200.planner import TrajectoryPlanner
    from oracle import c_oracle as co
    out = {}
    for name, sw, steps in (("config1_single_agent", sc.config1_single_agent(), 60), ("config2_one_of_10_agents", sc.config2_circle(n_swarms=1), 12)):
        pl = TrajectoryPlanner(sw.params, 1, sw.n, local_rank, max_nodes=MAX_NODES)
        full = TrajectoryPlanner(sw.params, sw.n, sw.n, local_rank, max_nodes=MAX_NODES)
        t_gpu, t_cpu, its = [], [], []
        for s in range(steps):
            b = sw.make_batch()
            for i in range(sw.n):
                bi = b.take([i])
                if s == 0 and i == 0:
                    pl.solve_batch(bi)  # first call: lazy allocations
                t0 = time.perf_counter()
                r = pl.solve_batch(bi)
                t_gpu.append(time.perf_counter() - t0)
                its.append(int(r["res"]["iters"][0]))
                t0 = time.perf_counter()
                co.solve_batch(bi, max_nodes=MAX_NODES, n_threads=1)
                t_cpu.append(time.perf_counter() - t0)
            o = full.solve_batch(b)
            sw.advance(o["traj"], o["ctrl"], okmask(o["res"]))
        pl.close()
        full.close()
        g, c = 1e3 * np.array(t_gpu), 1e3 * np.array(t_cpu)
        out[name] = {"calls": len(g), "gpu_ms_p50": float(np.percentile(g, 50)), "gpu_ms_p99": float(np.percentile(g, 99)),
                     "gpu_ms_max": float(g.max()), "cpu_port_1thread_ms_p50": float(np.percentile(c, 50)),
                     "cpu_port_1thread_ms_p99": float(np.percentile(c, 99)), "ipm_iters_p50_max": [float(np.median(its)), int(max(its))]}
    out["note"] = ("wall clock of hdsm_solve_batch(n_local=1) incl. staging, H2D, kernels, D2H; the reference allows 80 ms per solve "
                   "(TimeLimit) in a 100 ms period; CPU column = the C port on one thread (Gurobi runs single-threaded too, :948)")
    return out


def config2_measure(args, local_rank, flush, torch, swarms, steps):
    """Round 1's headline: BASELINE.json configs[1] as `swarms` independent 10-agent swarm instances, frozen snapshots,
    device-resident and end to end through hdsm_solve_batch with page-locked host buffers."""
    from multi_agent_pkgs_b200.planner import TrajectoryPlanner
    from multi_agent_pkgs_b200.swarm import DeviceBatch, algorithmic_bytes
    from multi_agent_pkgs_b200._lib import RESULT_DTYPE
    dev = torch.device(f"cuda:{local_rank}")
    params = sc.agile_params(10)
    gen = TrajectoryPlanner(params, max_agents=DISTINCT_SWARMS * 10, max_neighbours=10, device=local_rank, max_nodes=MAX_NODES)
    snaps = make_snapshots(gen.solve_batch, 2, swarms)
    gen.close()
    n_local = swarms * 10
    pl = TrajectoryPlanner(params, max_agents=n_local, max_neighbours=10, device=local_rank, max_nodes=MAX_NODES)
    dbs = [DeviceBatch(s_, dev) for s_ in snaps]
    balg = float(np.mean([algorithmic_bytes(s_).mean() for s_ in snaps]))
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    for k in range(4):
        pl.solve_batch_device(dbs[k % len(dbs)].t, dbs[k % len(dbs)].n_rob, sp)
    torch.cuda.synchronize()
    stat, iters, nodes = np.zeros(6, int), 0, 0
    for db in dbs:
        r = db.results()
        stat += np.bincount(r["status"], minlength=6)[:6]
        iters += int(r["iters"].sum())
        nodes += int(r["nodes"].sum())
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.fill_(k & 0xFF)
        ev[k][0].record(stream)
        pl.solve_batch_device(dbs[k % len(dbs)].t, dbs[k % len(dbs)].n_rob, sp)
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    # end to end: every input array and every result array page-locked (no staging copy inside the library)
    IN = ("global_id", "nbr_begin", "nbr_end", "x0", "ref", "poly_A", "poly_b", "poly_rows", "prev_self_pos", "all_pos", "all_valid")
    keep = []
    for s_ in snaps:
        for k in IN:
            tt = torch.from_numpy(np.ascontiguousarray(getattr(s_, k))).pin_memory()
            keep.append(tt)
            setattr(s_, k, tt.numpy())
    N, P = 10, 4
    outs = {"traj": torch.empty((n_local, N + 1, 9), dtype=torch.float64).pin_memory(), "ctrl": torch.empty((n_local, N, 3), dtype=torch.float64).pin_memory(),
            "poly_used": torch.empty((n_local, P), dtype=torch.uint8).pin_memory(), "assign": torch.empty((n_local, N), dtype=torch.int32).pin_memory(),
            "res": torch.empty((n_local, RESULT_DTYPE.itemsize), dtype=torch.uint8).pin_memory()}
    out_np = {k: v.numpy() for k, v in outs.items()}
    out_np["res"] = out_np["res"].view(RESULT_DTYPE).reshape(n_local)
    for k in range(2):
        pl.solve_batch(snaps[k % len(snaps)], out=out_np)
    t_e2e = 0.0
    for k in range(steps):
        t1 = time.perf_counter()
        pl.solve_batch(snaps[k % len(snaps)], out=out_np)
        t_e2e += time.perf_counter() - t1
    h2d = dbs[0].input_bytes()
    d2h = int(sum(v.numel() * v.element_size() for v in outs.values()))
    pl.close()
    tot = max(1, int(stat.sum()))
    return {"workload": f"config2: 10-agent circular exchange, forest map, N=10, {swarms} independent swarm instances ({n_local} agent QPs per step), "
                        f"frozen closed-loop snapshots at steps {list(SNAP_STEPS)} rotated, L2 flushed",
            "value": n_local / (ms * 1e-3), "unit": "solves/s", "ms_per_step": ms, "steps": steps,
            "e2e": {"value": n_local * steps / t_e2e, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "hdsm_solve_batch with page-locked caller buffers: copied by the copy engine directly, no staging memcpy"},
            "algorithmic_bytes_per_solve": balg, "ipm_iters_per_solve": iters / tot, "qp_relaxations_per_solve": nodes / tot,
            "status_counts": [int(v) for v in stat]}


def run_ours(args, rank, world, local_rank):
    # host-side producers first: the fork pools must exist before this process touches CUDA
    if args.agents % world:
        raise SystemExit(f"bench.py: {args.agents} agents do not split evenly over {world} GPUs")
    procs = max(1, (os.cpu_count() or 1) // world)
    W, T = max(args.warmup, 3), max(args.warmup, 3) + args.steps
    t0 = time.time()
    sw = sc.config5_random(seed=args.seed, n_rob=args.agents, side=SIDE)
    pool = sc.InputPool(sw, procs)
    sw_weak = pool_weak = None
    if world > 1 and args.weak_steps > 0:
        sw_weak = sc.config5_random(seed=args.seed, n_rob=args.agents * world, side=SIDE * float(np.sqrt(world)))
        pool_weak = sc.InputPool(sw_weak, procs)
    log(f"[rank {rank}] scenarios built in {time.time() - t0:.1f}s, {procs} producer processes")

    import torch
    from multi_agent_pkgs_b200.swarm import ClosedLoop, table_checksum
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from oracle import c_oracle as co  # checker of the parity column and CPU baseline legs only
    if rank == 0:
        co.build()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    # ---- headline: config 5, strong scaling, closed loop
    loop = ClosedLoop(sw, world, rank, dev, MAX_NODES, pool, record_host=True, width=WIDTH, warm_start=WARM)
    if rank == 0:
        loop.checker = lambda b: co.solve_batch(b, max_nodes=MAX_NODES, width=WIDTH, warm_start=WARM)
        loop.keep_steps, loop.keep_agents = {W, W + (T - W) // 2, T - 1}, min(loop.n, 2048)
    par = loop.preroll(T, parity_sample=512 if rank == 0 else 0, log=log)
    pool.close()
    clocks = ClockSampler(local_rank)
    launches0 = loop.planner.launch_count
    ms = timed_replay(loop, W, T, flush, dist, torch)
    launches = loop.planner.launch_count - launches0
    clk = clocks.stop()
    replay_sum = table_checksum(loop.current_table())
    same = replay_sum == loop.sums[T - 1]
    total_ms = max_over_ranks(ms.sum(), dist, torch, dev)
    value = args.agents * (T - W) / (total_ms * 1e-3)
    t_e2e, h2d, d2h = e2e_replay(loop, W, T, dist, torch)
    same = same and table_checksum(loop.current_table()) == loop.sums[T - 1]
    t_e2e = max_over_ranks(t_e2e, dist, torch, dev)
    same_all = max_over_ranks(0.0 if same else 1.0, dist, torch, dev) == 0.0
    quality = quality_of(loop.stats, W)
    from multi_agent_pkgs_b200.swarm import algorithmic_bytes
    balg = float(np.mean([algorithmic_bytes(b).mean() for b in loop.host_batches.values()])) if rank == 0 else 0.0
    smem = loop.planner.smem_bytes

    # ---- weak scaling beside it (N > 1): 4096 agents per GPU
    weak = None
    if sw_weak is not None:
        Ww, Tw = 3, 3 + args.weak_steps
        lw = ClosedLoop(sw_weak, world, rank, dev, MAX_NODES, pool_weak, width=WIDTH, warm_start=WARM)
        lw.preroll(Tw, log=log)
        pool_weak.close()
        msw = timed_replay(lw, Ww, Tw, flush, dist, torch)
        okw = max_over_ranks(0.0 if table_checksum(lw.current_table()) == lw.sums[Tw - 1] else 1.0, dist, torch, dev) == 0.0
        tw = max_over_ranks(msw.sum(), dist, torch, dev)
        weak = {"workload": f"config5 at {args.agents} agents PER GPU: {sw_weak.n} agents in a {SIDE * np.sqrt(world):.0f} m forest, "
                            f"{sw_weak.n} neighbour candidates per agent, closed loop",
                "value": sw_weak.n * (Tw - Ww) / (tw * 1e-3), "unit": "solves/s", "ms_per_step": tw / (Tw - Ww), "steps": Tw - Ww,
                "scaling": "weak", "replay_equals_preroll": okw, "quality": quality_of(lw.stats, Ww)}
        lw.close()

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "of fallback (B200_PROFILING.md 6.65 TB/s)"
        kernel_ms = total_ms / (T - W)
        achieved = balg * loop.n / (kernel_ms * 1e-3) / 1e9
        traffic, pipes, tnote = None, None, "no ncu capture of this launch size committed"
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            key = f"config5_{loop.n}_agents"
            if key in tj:
                traffic = float(tj[key]["dram_bytes_per_launch"])
                pipes = {k: v for k, v in tj[key].items() if k != "dram_bytes_per_launch"}
                tnote = f"dram bytes and pipe / stall figures of one {loop.n}-agent launch of this workload: ncu --set full capture in profiles/ ({tj[key].get('source', '?')})"
        # CPU baseline: the C port on all host threads on the kept steps of the same closed loop (bounded sample)
        cpu_n = cpu_t = 0.0
        reps = 0
        while cpu_t < args.cpu_seconds and reps < 50:
            for hb in loop.host_batches.values():
                t1 = time.perf_counter()
                co.solve_batch(hb, max_nodes=MAX_NODES, width=WIDTH, warm_start=WARM)
                cpu_t += time.perf_counter() - t1
                cpu_n += hb.n
            reps += 1
    loop.close()

    secondary = {}
    if rank == 0:
        torch.cuda.set_stream(torch.cuda.Stream(device=dev))
        for name, fn in (("config4", lambda: config4_measure(args, local_rank, flush, torch)),
                         ("latency", lambda: latency_measure(local_rank)),
                         ("config2_replicas", lambda: config2_measure(args, local_rank, flush, torch, args.swarms, min(args.steps, 12)) if (world == 1 and args.swarms > 0) else None)):
            try:
                secondary[name] = fn()
            except Exception as e:  # a secondary object never costs the headline line
                secondary[name] = {"error": f"{type(e).__name__}: {e}"}
        producers = {}
        if world == 1 and args.corridor_agents > 0:
            def roof(gbs, note):
                return {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None, "note": note}
            try:
                corridor, cor_gbs = corridor_measure(args.corridor_agents, 10, local_rank)
                corridor["roofline"] = roof(cor_gbs, "serial list logic in shared memory: latency bound, not HBM bound")
                producers["corridor"] = corridor
                mapping = map_measure(min(args.corridor_agents, 4096), 10, local_rank)
                mapping["roofline"] = roof(mapping["algorithmic_bytes_per_grid"] * mapping["value"] / 1e9,
                                           "one read and one write per voxel reach HBM; the time goes into the stencil walks in shared memory")
                producers["local_map"] = mapping
                reftraj = reftraj_measure(args.corridor_agents, 10, local_rank)
                reftraj["roofline"] = roof(reftraj["algorithmic_bytes_per_agent"] * reftraj["value"] / 1e9,
                                           "serial voxel traversal and pow/exp per visited voxel: latency bound")
                producers["reference_trajectory"] = reftraj
                producers["full_step"] = full_step_measure(min(args.corridor_agents, 2048), 10, local_rank)
                sensing = sense_measure(min(args.corridor_agents, 4096), 10, local_rank)
                sensing["roofline"] = roof(sensing["algorithmic_bytes_per_agent"] * sensing["value"] / 1e9,
                                           "kept grid in, new grid out; the time goes into 14 k serial FP64 voxel traversals per agent")
                producers["local_map_acquisition"] = sensing
            except Exception as e:
                producers["error"] = f"{type(e).__name__}: {e}"
        line = {"metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": T - W,
                "warmup": W, "ms_per_step": total_ms / (T - W), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(args, world),
                "e2e": {"value": args.agents * (T - W) / t_e2e, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches),
                "closed_loop": {"replay_equals_preroll_on_every_rank": bool(same_all), "table_checksum": replay_sum,
                                "note": "checksum of the all-gathered plan table after the last timed step; identical on any number of GPUs"},
                "parity_vs_c_port": par,
                "quality": quality,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_solve": balg, "ncu": pipes,
                             "note": "algorithmic bytes per solve x agents per launch / launch time (SURVEY 8(d)); at 4096 neighbour candidates "
                                     "the 1.08 MB plan table is L2 resident and the solve itself is FP64-latency bound (DESIGN.md section 5). " + tnote},
                "cpu_baseline": {"value": cpu_n / cpu_t, "unit": "solves/s", "cores": co.max_threads(), "kind": "port",
                                 "sample": f"{int(cpu_n)} agent QPs ({len(loop.host_batches)} kept steps of the same closed loop, first "
                                           f"{loop.keep_agents} agents of rank 0's shard) in {cpu_t:.1f}s, C port of the oracle on all host "
                                           f"threads (Gurobi not available)"},
                "clocks": clk, "weak": weak, "smem_bytes_per_block": smem}
        line.update(secondary)
        line.update(producers)
        # the line is long; whoever keeps only its head or only its tail still gets the essentials
        def g(d, *ks):
            for k in ks:
                d = d.get(k) if isinstance(d, dict) else None
            return d
        line["digest"] = {
            "config5_ms_per_step": total_ms / (T - W), "config5_solves_per_s": value, "config5_e2e_solves_per_s": args.agents * (T - W) / t_e2e,
            "n_gpus": world, "table_checksum": replay_sum, "replay_equals_preroll": bool(same_all),
            "parity_status_mismatches": par.get("status_mismatches"), "parity_max_rel_obj_gap": par.get("max_rel_obj_gap"),
            "status_counts": quality["status_counts"], "weak_solves_per_s": g(weak, "value"), "weak_ms_per_step": g(weak, "ms_per_step"),
            "config4_ms_per_step": g(secondary, "config4", "ms_per_step"), "config4_cpu_port_ms_per_step": g(secondary, "config4", "cpu_baseline", "ms_per_step"),
            "latency_ms_p50_p99_config1": [g(secondary, "latency", "config1_single_agent", "gpu_ms_p50"), g(secondary, "latency", "config1_single_agent", "gpu_ms_p99")],
            "latency_ms_p50_p99_config2_agent": [g(secondary, "latency", "config2_one_of_10_agents", "gpu_ms_p50"), g(secondary, "latency", "config2_one_of_10_agents", "gpu_ms_p99")],
            "config2_replicas_solves_per_s": g(secondary, "config2_replicas", "value"), "config2_replicas_e2e": g(secondary, "config2_replicas", "e2e", "value"),
            "chained_step_status_counts": g(producers, "corridor", "chained_step", "status_counts"), "chained_step_ms": g(producers, "corridor", "chained_step", "ms"),
            "full_step_ms": g(producers, "full_step", "ms_per_step"), "full_step_stage_ms": g(producers, "full_step", "stage_ms"),
            "cpu_port_solves_per_s": cpu_n / cpu_t, "cpu_cores": co.max_threads()}
        emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--agents", type=int, default=N_AGENTS, help="agents of the config-5 swarm (whole job)")
    ap.add_argument("--weak-steps", type=int, default=8, help="timed steps of the weak-scaling object at N > 1 (0 = skip)")
    ap.add_argument("--swarms", type=int, default=4096, help="independent 10-agent swarm instances of the config-2 object (0 = skip)")
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--corridor-agents", type=int, default=8192,
                    help="agents of the secondary producer measurements at N = 1 (0 = skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
