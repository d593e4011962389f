"""Host-side logic of the multi-GPU path on CPU: world_size 2, gloo backend.

Covers the agent partition, the all-gather that rebuilds the neighbour table on every rank (even and
ragged shards) and that per-shard input generation equals slicing the whole-swarm inputs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_agent_pkgs_b200 import scenarios as sc
from multi_agent_pkgs_b200.swarm import Exchange, algorithmic_bytes, shard_range, shard_sizes


def test_shard_range_partitions():
    for n in (1, 7, 10, 256, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(n, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n


def test_algorithmic_bytes_matches_survey_figures():
    """5.6 KB at 10 agents with 12-row polytopes (N = 10, P = 4)."""
    b = sc.config2_circle().make_batch()
    b.poly_rows[:] = 12
    b.all_valid[:] = 1
    assert abs(algorithmic_bytes(b)[0] - 5600) < 150


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rob, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N = 10
        ex = Exchange(n_rob, N, world, rank, torch.device("cpu"))
        lo, hi = shard_range(n_rob, world, rank)
        # every agent's "plan" is a function of its global id, so the gathered table is checkable
        ids = torch.arange(lo, hi, dtype=torch.float64)
        pos = ids[:, None, None] * 1000 + torch.arange(N + 1, dtype=torch.float64)[None, :, None] * 10 \
            + torch.arange(3, dtype=torch.float64)[None, None, :]
        table = ex.allgather(pos)
        gids = torch.arange(n_rob, dtype=torch.float64)
        want = gids[:, None, None] * 1000 + torch.arange(N + 1, dtype=torch.float64)[None, :, None] * 10 \
            + torch.arange(3, dtype=torch.float64)[None, None, :]
        ok = bool(torch.equal(table, want))
        # shard inputs == slice of the whole-swarm inputs (same seed on every rank)
        sw_all = sc.config2_circle(n_swarms=n_rob // 10 if n_rob % 10 == 0 else 1, seed=5)
        sw_me = sc.config2_circle(n_swarms=n_rob // 10 if n_rob % 10 == 0 else 1, seed=5)
        if sw_all.n == n_rob:
            full = sw_all.make_batch()
            mine = sw_me.make_batch(np.arange(lo, hi))
            for key in ("x0", "ref", "poly_rows", "prev_self_pos", "global_id", "nbr_begin", "nbr_end"):
                ok &= bool(np.array_equal(getattr(mine, key), getattr(full, key)[lo:hi]))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rob", [20, 23])
def test_allgather_rebuilds_table_world2(n_rob):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rob, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
