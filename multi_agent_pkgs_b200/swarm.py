"""Synthetic-swarm harness: agents sharded over the GPUs of one box, one process per GPU.

In the reference every agent is its own OS process and broadcasts its new plan to all others
over ROS2 (publisher agent_class.cpp:645-677, subscribers :610-643).  Here rank r owns the agents
``shard_range(n_rob, world, r)``; after each replanning step the packed positions of the new plans
(written by the solver's epilogue) are exchanged with ONE all-gather, which rebuilds the table
``all_pos[n_rob][N+1][3]`` on every rank - the only cross-agent data the path needs
(SURVEY.md section 8(e)).  On GPUs the all-gather is ncclAllGather through the library's own
communicator (hdsm_allgather_positions); on CPU tensors (tests, gloo) it is
torch.distributed.all_gather_into_tensor, so the host-side logic is testable without a GPU.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Block partition by agent id: rank r owns [r*n/world, (r+1)*n/world) (sizes differ by <= 1)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world: int):
    return [shard_range(n, world, r)[1] - shard_range(n, world, r)[0] for r in range(world)]


class DeviceBatch:
    """One replanning step's inputs and outputs as contiguous torch tensors in the C-ABI layouts."""

    IN_KEYS = ("global_id", "nbr_begin", "nbr_end", "x0", "ref", "poly_A", "poly_b", "poly_rows", "prev_self_pos",
               "all_pos", "all_valid")

    def __init__(self, batch, device, with_outputs: bool = True, pin: bool = False):
        import torch
        self.n = int(batch.x0.shape[0])
        self.n_rob = int(batch.all_pos.shape[0])
        N, P = int(batch.params["n_hor"]), int(batch.params["poly_hor"])
        self.N, self.P = N, P
        dt = {"global_id": np.int32, "nbr_begin": np.int32, "nbr_end": np.int32, "poly_rows": np.int32,
              "all_valid": np.uint8}
        self.t: Dict[str, "torch.Tensor"] = {}
        self.host: Dict[str, "torch.Tensor"] = {}
        for k in self.IN_KEYS:
            a = np.ascontiguousarray(getattr(batch, k), dtype=dt.get(k, np.float64))
            h = torch.from_numpy(a)
            if pin:
                h = h.pin_memory()
            self.host[k] = h
            self.t[k] = h.to(device, non_blocking=pin)
        if with_outputs:
            f64 = torch.float64
            self.t["traj"] = torch.zeros((self.n, N + 1, 9), dtype=f64, device=device)
            self.t["ctrl"] = torch.zeros((self.n, N, 3), dtype=f64, device=device)
            self.t["poly_used"] = torch.zeros((self.n, P), dtype=torch.uint8, device=device)
            self.t["assign_out"] = torch.zeros((self.n, N), dtype=torch.int32, device=device)
            self.t["res"] = torch.zeros((self.n, 32), dtype=torch.uint8, device=device)  # hdsm_result[n]
            self.t["pos_out"] = torch.zeros((self.n, N + 1, 3), dtype=f64, device=device)

    def input_bytes(self) -> int:
        return int(sum(self.host[k].numel() * self.host[k].element_size() for k in self.IN_KEYS))

    def results(self):
        from ._lib import RESULT_DTYPE
        return np.frombuffer(self.t["res"].cpu().numpy().tobytes(), dtype=RESULT_DTYPE)


def algorithmic_bytes(batch) -> np.ndarray:
    """Compulsory HBM traffic per agent QP: every input read once, every
    output written once, FP64."""
    N, P = int(batch.params["n_hor"]), int(batch.params["poly_hor"])
    rows = np.asarray(batch.poly_rows).sum(axis=1)
    valid = np.asarray(batch.all_valid).astype(bool)
    csum = np.concatenate([[0], np.cumsum(valid)])
    lo, hi = np.asarray(batch.nbr_begin), np.asarray(batch.nbr_end)
    gid = np.asarray(batch.global_id)
    n_nb = csum[hi] - csum[lo] - valid[gid]
    words = 9 + 6 * N + 4 * rows + 3 * (N + 1) + 3 * N * n_nb + 9 * (N + 1) + 3 * N
    return 8 * words + 4 * N + P + 32


class Exchange:
    """All-gather of the packed plan positions.  ``table`` is [n_rob][N+1][3] on every rank."""

    def __init__(self, n_rob: int, n_hor: int, world: int, rank: int, device, planner=None):
        import torch
        self.world, self.rank, self.n_rob, self.N = world, rank, n_rob, n_hor
        self.lo, self.hi = shard_range(n_rob, world, rank)
        self.even = n_rob % world == 0
        self.device = device
        self.planner = planner
        self.table = torch.zeros((n_rob, n_hor + 1, 3), dtype=torch.float64, device=device)
        self._nccl = False
        if world > 1 and planner is not None and str(device).startswith("cuda"):
            import torch.distributed as dist
            uid = [planner.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            planner.comm_init(world, rank, uid[0])
            self._nccl = self.even  # ncclAllGather needs equal counts; ragged shards use torch's path

    def allgather(self, pos_out, stream_ptr: int = 0):
        """pos_out: [n_local][N+1][3] of this rank.  Fills self.table (stream ordered on GPU)."""
        import torch
        if self.world == 1:
            self.table.copy_(pos_out)
            return self.table
        if self._nccl:
            self.planner.allgather_positions(pos_out, self.table, self.hi - self.lo, stream_ptr)
            return self.table
        import torch.distributed as dist
        if self.even:
            dist.all_gather_into_tensor(self.table, pos_out.contiguous())
        else:  # ragged shards: pad every rank's block to the largest shard, gather, drop the padding
            sizes = shard_sizes(self.n_rob, self.world)
            mx = max(sizes)
            send = torch.zeros((mx, self.N + 1, 3), dtype=torch.float64, device=self.device)
            send[: pos_out.shape[0]] = pos_out
            recv = torch.empty((self.world * mx, self.N + 1, 3), dtype=torch.float64, device=self.device)
            dist.all_gather_into_tensor(recv, send)
            recv = recv.view(self.world, mx, self.N + 1, 3)
            off = 0
            for r, sz in enumerate(sizes):
                self.table[off:off + sz] = recv[r, :sz]
                off += sz
        return self.table


class ShardedSwarm:
    """Closed-loop driver: scenario state on the host (CPU producers of ref / corridors are out of
    scope), per-step solve of this rank's shard on the GPU, exchange, state advance."""

    def __init__(self, swarm, world: int = 1, rank: int = 0, device="cuda:0", max_nodes: int = 64, **kw):
        from .planner import TrajectoryPlanner
        self.swarm, self.world, self.rank, self.device = swarm, world, rank, device
        self.lo, self.hi = shard_range(swarm.n, world, rank)
        nn = int((swarm.group_end - swarm.group_begin).max())
        dev_index = int(str(device).split(":")[1]) if ":" in str(device) else 0
        self.planner = TrajectoryPlanner(swarm.params, self.hi - self.lo, nn, dev_index, max_nodes=max_nodes, **kw)
        self.exchange = Exchange(swarm.n, swarm.params["n_hor"], world, rank, device, self.planner)
        self.have = np.zeros(swarm.n, np.uint8)
        import torch
        self.stream = torch.cuda.Stream(device=device)  # non-NULL: NULL selects the handle's own stream

    def step(self):
        """One replanning step of the whole swarm; returns this rank's hdsm_result array."""
        import torch
        ids = np.arange(self.lo, self.hi)
        batch = self.swarm.make_batch(ids)
        db = DeviceBatch(batch, self.device)
        db.t["all_pos"] = self.exchange.table          # the table rebuilt by the previous exchange
        db.t["all_valid"] = torch.from_numpy(self.have.copy()).to(self.device)
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.planner.solve_batch_device(db.t, self.swarm.n, self.stream.cuda_stream)
            self.exchange.allgather(db.t["pos_out"], self.stream.cuda_stream)
        self.stream.synchronize()
        res = db.results()
        ok = (res["status"] == 0) | ((res["status"] == 4) & np.isfinite(res["obj"]))
        self.swarm.advance(db.t["traj"].cpu().numpy(), db.t["ctrl"].cpu().numpy(), ok, ids)
        # every rank needs every agent's state / plan flags to build the next step's inputs of its
        # shard only for the planes, and those come from the table; plan validity is exchanged too
        have_local = self.swarm.have_plan[self.lo:self.hi].copy()
        if self.world > 1:
            import torch.distributed as dist
            sizes = shard_sizes(self.swarm.n, self.world)
            outs = [torch.empty(s, dtype=torch.uint8) for s in sizes]
            gl = torch.from_numpy(have_local)
            if dist.get_backend() == "nccl":
                outs = [o.to(self.device) for o in outs]
                gl = gl.to(self.device)
            dist.all_gather(outs, gl)
            self.have = torch.cat([o.cpu() for o in outs]).numpy()
        else:
            self.have = have_local
        return res


def table_checksum(t) -> str:
    """Order-sensitive 64-bit checksum of a float64 tensor's bit patterns (xor of bits rotated by position parity,
    plus the wrapped integer sum): equal tables give equal checksums on any number of GPUs."""
    import torch
    bits = t.contiguous().view(torch.int64).flatten()
    odd = bits[1::2]
    x = int(torch.bitwise_xor(bits[0::2].sum(), (odd * 3).sum()).item()) & 0xFFFFFFFFFFFFFFFF
    return f"{x:016x}"


class ClosedLoop:
    """Closed-loop replanning of a sharded swarm with the swarm state resident in HBM.

    Per replanning step and rank, all on one stream and without host synchronisation:
      1. hdsm_solve_batch_device on the rank's shard - inter-agent planes from the neighbour table the previous
         exchange produced, assignment search, interior-point solves, position pack (agent_class.cpp:168, :174);
      2. hdsm_advance_device - read-back, failure fallback, state advance (:962-1019, :233-238);
      3. hdsm_exchange_plans - ONE NCCL group: all-gather of the packed plan positions into the table every rank
         reads in the next step, and of the "plan received" flags (the ROS2 broadcast, :645-677 / :629-643).
    The path's exogenous inputs - reference trajectory and corridor cells, whose producers stay on the host in this
    harness - are computed once per step by `preroll` (host producers, untimed) and kept in HBM; `replay` then runs the
    same closed loop from the same initial state with only device work.  The solver is deterministic, so the replay
    reproduces the pre-roll bit for bit (`checks` compares the tables), on any number of ranks.
    """

    def __init__(self, swarm, world: int = 1, rank: int = 0, device="cuda:0", max_nodes: int = 64, pool=None,
                 record_host: bool = False, **planner_kw):
        import torch
        from .planner import TrajectoryPlanner
        self.torch = torch
        self.swarm, self.world, self.rank, self.device, self.pool = swarm, world, rank, torch.device(device), pool
        self.lo, self.hi = shard_range(swarm.n, world, rank)
        self.n, self.n_rob = self.hi - self.lo, swarm.n
        if world > 1 and swarm.n % world:
            raise ValueError("ClosedLoop needs equal shards (ncclAllGather): n_rob must be a multiple of the world size")
        p = swarm.params
        self.N, self.P, self.R = int(p["n_hor"]), int(p["poly_hor"]), int(swarm.rmax)
        nn = int((swarm.group_end - swarm.group_begin).max())
        self.planner = TrajectoryPlanner(p, self.n, nn, self.device.index or 0, rmax=self.R, max_nodes=max_nodes, **planner_kw)
        if world > 1:
            import torch.distributed as dist
            uid = [self.planner.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            self.planner.comm_init(world, rank, uid[0])
        dev, f64, n, N, P, R = self.device, torch.float64, self.n, self.N, self.P, self.R
        ids = np.arange(self.lo, self.hi)
        self.ids = ids
        z = lambda *s, dt=f64: torch.zeros(s, dtype=dt, device=dev)  # noqa: E731
        self.t = {
            "global_id": torch.from_numpy(ids.astype(np.int32)).to(dev),
            "nbr_begin": torch.from_numpy(swarm.group_begin[ids].astype(np.int32)).to(dev),
            "nbr_end": torch.from_numpy(swarm.group_end[ids].astype(np.int32)).to(dev),
            "x0": z(n, 9), "traj_curr": z(n, N + 1, 9), "ctrl_curr": z(n, N, 3), "have_plan": z(n, dt=torch.uint8),
            "traj": z(n, N + 1, 9), "ctrl": z(n, N, 3), "poly_used": z(n, P, dt=torch.uint8),
            "assign_out": z(n, N, dt=torch.int32), "res": z(n, 32, dt=torch.uint8),
        }
        self.pos = [z(n, N + 1, 3), z(n, N + 1, 3)]           # ping-pong: prev_self_pos (read) / pos_out (written)
        self.table = z(self.n_rob, N + 1, 3) if world > 1 else None
        self.valid = z(self.n_rob, dt=torch.uint8) if world > 1 else None
        self.x0_init = torch.from_numpy(np.ascontiguousarray(swarm.state[ids])).to(dev)
        self.rec: list = []        # per step: dict(ref, poly_A, poly_b, poly_rows) on the device
        self.rec_host: list = []   # the same in pinned host memory (end-to-end timing)
        self.record_host = record_host
        self.sums: list = []       # per step: checksum of the table after the exchange
        self.stats: list = []      # per step: hdsm_result array of the shard
        self.step_index = 0
        self.stream = torch.cuda.Stream(device=dev)
        torch.cuda.synchronize(dev)  # the uploads above ran on another stream than the one every step is enqueued on
        self.reset()

    # ------------------------------------------------------------------------------------------------
    def reset(self):
        """Back to the swarm's initial state: no plans, no neighbour planes (agent_class.cpp:1134)."""
        t = self.t
        with self.torch.cuda.stream(self.stream):
            t["x0"].copy_(self.x0_init)
            t["have_plan"].zero_()
            t["traj_curr"].copy_(self.x0_init[:, None, :].expand(-1, self.N + 1, -1))
            t["ctrl_curr"].zero_()
            self.pos[0].copy_(self.x0_init[:, None, :3].expand(-1, self.N + 1, -1))
            self.pos[1].zero_()
            if self.world > 1:
                self.table.zero_()
                self.valid.zero_()
        self.stream.synchronize()
        self.step_index = 0

    def _bind(self, s):
        """Argument dictionary of step s: recorded exogenous inputs + live state."""
        t, r = self.t, self.rec[s]
        t["ref"], t["poly_A"], t["poly_b"], t["poly_rows"] = r["ref"], r["poly_A"], r["poly_b"], r["poly_rows"]
        t["prev_self_pos"], t["pos_out"] = self.pos[s & 1], self.pos[(s & 1) ^ 1]
        if self.world > 1:
            t["all_pos"], t["all_valid"] = self.table, self.valid
        else:  # one rank: the previous step's packed plans ARE the table
            t["all_pos"], t["all_valid"] = self.pos[s & 1], t["have_plan"]
        return t

    def device_step(self, s, inputs=None):
        """Enqueue step s on self.stream (no synchronisation).  `inputs`: dict(ref, poly_A, poly_b, poly_rows) of
        device tensors to use instead of the recorded ones (end-to-end timing uploads them every step)."""
        sp = self.stream.cuda_stream
        t = self._bind(s)
        if inputs is not None:
            t.update(inputs)
        self.planner.solve_batch_device(t, self.n_rob, sp)
        prev = t.pop("prev_self_pos")  # the advance call must not overwrite the buffer the table aliases
        self.planner.advance_device(t, sp)
        t["prev_self_pos"] = prev
        if self.world > 1:
            self.planner.exchange_plans(t["pos_out"], self.table, t["have_plan"], self.valid, self.n, sp)
        self.step_index = s + 1

    def current_table(self):
        return self.table if self.world > 1 else self.pos[self.step_index & 1]

    # ------------------------------------------------------------------------------------------------
    def preroll(self, steps: int, parity_sample: int = 0, log=None):
        """Run `steps` closed-loop steps with the host producers in the loop, recording their outputs.  With
        parity_sample > 0 the first that many agents of the shard are also solved by the caller-supplied checker
        (`self.checker(batch) -> dict(res=...)`, set by tests / bench) on the same inputs; returns the parity record."""
        import time
        torch, sw = self.torch, self.swarm
        par = {"agents": 0, "status_mismatches": 0, "max_rel_obj_gap": 0.0}
        t_host = t_dev = 0.0
        keep = getattr(self, "keep_steps", ())  # steps whose complete host-side inputs are kept (CPU baseline legs)
        if not hasattr(self, "host_batches"):
            self.host_batches = {}
        for s in range(len(self.rec), steps):
            t0 = time.perf_counter()
            x0 = self.t["x0"].cpu().numpy()
            ref, pA, pb, pr = sw.make_inputs(self.ids, pos=x0[:, :3], step=s, pool=self.pool)
            rec = {"ref": torch.from_numpy(ref), "poly_A": torch.from_numpy(pA), "poly_b": torch.from_numpy(pb),
                   "poly_rows": torch.from_numpy(pr)}
            if self.record_host:
                self.rec_host.append({k: v.pin_memory() for k, v in rec.items()})
            with torch.cuda.stream(self.stream):  # stream-ordered with the kernels that read them
                self.rec.append({k: v.to(self.device, non_blocking=False) for k, v in rec.items()})
            t1 = time.perf_counter()
            check = parity_sample > 0 and getattr(self, "checker", None) is not None
            if check or s in keep:
                m = min(max(parity_sample, getattr(self, "keep_agents", 0) if s in keep else 0), self.n)
                tab = self.current_table().cpu().numpy()
                val = (self.valid if self.world > 1 else self.t["have_plan"]).cpu().numpy()
                if self.world == 1 and s == 0:
                    tab, val = np.zeros((self.n_rob, self.N + 1, 3)), np.zeros(self.n_rob, np.uint8)
                from .scenarios import Batch
                hb = Batch(sw.params, self.ids[:m].astype(np.int32), sw.group_begin[self.ids[:m]].astype(np.int32),
                           sw.group_end[self.ids[:m]].astype(np.int32), x0[:m], ref[:m], pA[:m], pb[:m], pr[:m],
                           self.pos[s & 1][:m].cpu().numpy(), tab, val, self.R)
                if s in keep:
                    self.host_batches[s] = hb
                if check:
                    m = min(parity_sample, self.n)
                    want = self.checker(hb.take(np.arange(m)) if hb.n > m else hb)["res"]
            with torch.cuda.stream(self.stream):
                self.device_step(s)
            self.stream.synchronize()
            res = np.frombuffer(self.t["res"].cpu().numpy().tobytes(), dtype=_result_dtype()).copy()
            self.stats.append(res)
            self.sums.append(table_checksum(self.current_table()))
            if check:
                got = res[:m]
                par["agents"] += m
                par["status_mismatches"] += int((got["status"] != want["status"]).sum())
                both = (got["status"] == 0) & (want["status"] == 0)
                if both.any():
                    gap = np.abs(got["obj"][both] - want["obj"][both]) / np.maximum(1.0, np.abs(want["obj"][both]))
                    par["max_rel_obj_gap"] = max(par["max_rel_obj_gap"], float(gap.max()))
            t_host += t1 - t0
            t_dev += time.perf_counter() - t1
        if log:
            log(f"[rank {self.rank}] pre-roll: {steps} steps x {self.n} agents, host producers {t_host:.1f}s, rest {t_dev:.1f}s")
        return par

    def input_bytes_per_step(self) -> int:
        r = self.rec[0]
        return int(sum(v.numel() * v.element_size() for v in r.values()))

    def close(self):
        self.planner.close()


def _result_dtype():
    from ._lib import RESULT_DTYPE
    return RESULT_DTYPE
