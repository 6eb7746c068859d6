"""Host logic of the N>1 path on CPU: gloo, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tsid_control_b200.sharding import gather_diagnostics, shard_range, status_histogram


def test_shard_range_partitions_exactly():
    for n, w in ((65536, 8), (1000003, 8), (10, 4), (7, 8)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 6
    status = torch.full((n,), rank, dtype=torch.int32)
    iters = torch.arange(n, dtype=torch.int32) + 100 * rank
    diag = gather_diagnostics(status, iters)
    q.put((rank, diag.tolist(), status_histogram(diag).tolist()))
    dist.destroy_process_group()


def test_gather_diagnostics_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, diag, hist in res:
        assert len(diag) == 12
        assert [d[0] for d in diag] == [0] * 6 + [1] * 6  # global env order: rank 0's shard then rank 1's
        assert [d[1] for d in diag[6:]] == [100 + i for i in range(6)]
        assert hist == [0, 6, 6, 0, 0, 0]


def test_single_process_passthrough():
    d = gather_diagnostics(torch.zeros(4, dtype=torch.int32), torch.ones(4, dtype=torch.int32))
    assert d.shape == (4, 2)
