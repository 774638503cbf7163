"""N > 1 host logic on CPU: world_size-2 gloo, edge sharding + the one all-gather of the per-edge summary."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from plant_motion_planning_b200.distributed import (all_gather_summary, pack_summary, shard_bounds, shard_capacity,
                                                    unpack_summary)


def test_shard_bounds_cover_everything():
    for E in (0, 1, 7, 64, 1000003):
        for ws in (1, 2, 3, 8):
            spans = [shard_bounds(E, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == E
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == shard_capacity(E, ws) or E == 0


def test_pack_unpack_roundtrip():
    E, ws = 11, 3
    fh = torch.arange(E, dtype=torch.int32)
    ns = fh + 5
    cost = torch.linspace(0, 1, E)
    cap = shard_capacity(E, ws)
    parts = []
    for r in range(ws):
        lo, hi = shard_bounds(E, r, ws)
        parts.append(pack_summary(fh[lo:hi], ns[lo:hi], cost[lo:hi], cap))
    a, b, c = unpack_summary(torch.cat(parts), E, ws)
    assert torch.equal(a, fh) and torch.equal(b, ns) and torch.equal(c, cost)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, E, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        lo, hi = shard_bounds(E, rank, ws)
        fh = torch.arange(lo, hi, dtype=torch.int32) * 3
        ns = torch.arange(lo, hi, dtype=torch.int32) + 1
        cost = torch.arange(lo, hi, dtype=torch.float32) * 0.5
        a, b, c = all_gather_summary(fh, ns, cost, E)
        ok = (torch.equal(a, torch.arange(E, dtype=torch.int32) * 3) and torch.equal(b, torch.arange(E, dtype=torch.int32) + 1)
              and torch.equal(c, torch.arange(E, dtype=torch.float32) * 0.5))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("E", [9, 64])
def test_all_gather_summary_gloo_world2(E):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, E, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
