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
