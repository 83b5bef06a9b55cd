"""Row-sharded exact kNN over the GPUs of one box.

One process per GPU (torchrun).  Every query row's top-k depends only on that row
and the full candidate set, so the contigs are block-sharded: rank g owns rows
[g*per, (g+1)*per), computes the neighbours of ITS rows against ALL rows, and no
cross-rank merge is needed.  The exchange is one all-gather of the float64
profile rows over NCCL/NVLink (N*D*8 bytes: 240 MB for 1 M x 30), after which each
rank runs the engine's default pipeline -- for large Euclidean problems the grouped
staging and the pruned tensor-core scan -- on its run of the rows.  The staging's one
pass over EVERY row, the nearest-leader search, is split as well: every rank searches
its own rows and two small all-gathers (4 + 8 bytes per row) put the keys together; the
leaders themselves come out identical on every rank (deterministic kernels on identical
rows).  Candidate indices are global row numbers, so the lowest-index tie rule is the
same for every GPU count.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import _lib

TILE = _lib.TILE_ROWS


def shard_rows(n: int, world: int) -> int:
    """Rows per shard: ceil(n / world) rounded up to whole 64-row profile tiles."""
    per = -(-n // world)
    return -(-per // TILE) * TILE


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Global row range [r0, r1) owned by ``rank`` (trailing ranks may be empty)."""
    per = shard_rows(n, world)
    r0 = min(n, rank * per)
    return r0, min(n, r0 + per)


def gather_rows(x_local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather the row shards into the full [n_total, ...] array on every rank.

    Works on NCCL (CUDA tensors) and gloo (CPU tensors, used by the CPU tests)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = shard_rows(n_total, world)
    r0, r1 = shard_bounds(n_total, world, rank)
    if x_local.shape[0] != r1 - r0:
        raise ValueError(f"rank {rank} holds {x_local.shape[0]} rows, its shard is [{r0}, {r1})")
    tail = tuple(x_local.shape[1:])
    full = torch.empty((world * per,) + tail, dtype=x_local.dtype, device=x_local.device)
    if x_local.shape[0] == per:
        mine = x_local.contiguous()
    else:
        mine = torch.zeros((per,) + tail, dtype=x_local.dtype, device=x_local.device)
        mine[: x_local.shape[0]] = x_local
    dist.all_gather_into_tensor(full, mine, group=group)
    # only the last non-empty shard is short, so the first n_total rows are exactly the
    # global array in row order
    return full[:n_total]


_ASSIGN_PIPES: dict = {}


def _nearest_leader_sharded(lead: torch.Tensor, full: torch.Tensor, r0: int, r1: int, group):
    """(leader index, float64 distance to it) of every row, every rank assigning its own rows [r0, r1) with the
    library's own assignment (gmd_assign_leaders_f64: the same call the single-GPU staging makes); two
    all-gathers put the shards together."""
    from . import _lib
    from .engine import _ptr, _stream_ptr

    n, d = full.shape
    dev = full.device
    m = r1 - r0
    idx = torch.empty(m, dtype=torch.int32, device=dev)
    dd = torch.empty(m, dtype=torch.float64, device=dev)
    if m > 0:
        lib = _lib.lib()
        lead = lead.contiguous()
        key = (m, lead.shape[0], d, str(dev))
        ws = _ASSIGN_PIPES.get(key)
        if ws is None:
            _ASSIGN_PIPES.clear()                     # the number of leaders is stable for a data set: keep one
            ws = _ASSIGN_PIPES[key] = torch.empty(lib.gmd_knn_workspace_bytes(m, lead.shape[0], d, 1),
                                                  dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.gmd_assign_leaders_f64(_ptr(full[r0:r1]), m, _ptr(lead), lead.shape[0], d, _ptr(idx),
                                                  _ptr(dd), _ptr(ws), ws.numel(), _stream_ptr()))
    return gather_rows(idx, n, group).contiguous(), gather_rows(dd, n, group).contiguous()


def sharded_knn(x_local: torch.Tensor, n_total: int, k: int, metric="euclidean", group=None,
                pipeline=None, form: Optional[int] = None):
    """Exact kNN graph rows for this rank's shard: (idx [n_local, k] int32 global
    candidate indices, dist [n_local, k] float64), ascending (distance, index).

    ``x_local``: this rank's contiguous float64 [n_local, D] CUDA rows.  Pass a
    ``KnnPipeline`` to reuse device buffers across calls (bench steps)."""
    from .engine import KnnPipeline

    full = gather_rows(x_local, n_total, group)
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    r0, r1 = shard_bounds(n_total, world, rank)
    if pipeline is None and r1 > r0:
        pipeline = KnnPipeline(r1 - r0, n_total, full.shape[1], k, metric, form, same=False,
                               device=x_local.device)
    # the split staging needs every rank in its collectives: agree on it once per pipeline (ranks without rows
    # go along with whatever the others do)
    mode = getattr(pipeline, "_shard_mode", None) if pipeline is not None else None
    if mode is None:
        mine = 1 if (pipeline is None or (pipeline.prune and 0 < r1 - r0 < n_total)) else 0
        t = torch.tensor([mine], dtype=torch.int32, device=x_local.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        mode = bool(t.item()) and world > 1
        if pipeline is not None:
            pipeline._shard_mode = mode
    if r1 == r0:
        if mode:
            # no rows here, but the others' all-gathers of the nearest-leader keys need this rank
            _nearest_leader_sharded(torch.empty((1, full.shape[1]), dtype=torch.float64, device=full.device), full, r0, r1,
                                    group)
        e = dict(device=x_local.device)
        return torch.empty((0, k), dtype=torch.int32, **e), torch.empty((0, k), dtype=torch.float64, **e)
    q = full[r0:r1]                                   # a run of the candidates' rows: the staging shares the keys
    pipeline.shard_assign = (lambda lead: _nearest_leader_sharded(lead, full, r0, r1, group)) if mode else None
    try:
        return pipeline.run(q, full)
    finally:
        pipeline.shard_assign = None
