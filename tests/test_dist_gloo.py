"""CPU, world_size 2 over gloo: the host side of the row-sharded path (shard layout and
the candidate all-gather) without a GPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from groopm_b200 import dist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.from_numpy(np.random.default_rng(7).standard_normal((n, d)))
        r0, r1 = dist.shard_bounds(n, world, rank)
        got = dist.gather_rows(full[r0:r1].clone(), n)
        ok = got.shape == (n, d) and torch.equal(got, full)
        # a rank that passes the wrong number of rows is told so
        try:
            dist.gather_rows(torch.zeros((r1 - r0 + 1, d), dtype=torch.float64), n)
            bad_caught = False
        except ValueError:
            bad_caught = True
        ret[rank] = bool(ok and bad_caught)
    finally:
        tdist.destroy_process_group()


def _run(world, n, d):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret[r] for r in range(world))


def test_gather_rows_world2_ragged():
    _run(2, 1000, 5)   # shards of 512 + 488 rows


def test_gather_rows_world2_tiny():
    _run(2, 40, 3)     # second rank owns nothing
