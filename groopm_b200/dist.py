"""Row-sharded exact kNN over the GPUs of one box.

One process per GPU (torchrun).  Every query row's top-k depends only on that row
and the full candidate set, so the contigs are block-sharded: rank g owns rows
[g*per, (g+1)*per), computes the neighbours of ITS rows against ALL rows, and no
cross-rank merge is needed.  The only exchange is one all-gather of the float64
profile rows over NCCL/NVLink (N*D*8 bytes: 240 MB for 1 M x 30), after which each
rank stages the candidate tiles locally and scans them for its own rows.
Candidate indices are global row numbers, so the lowest-index tie rule is the same
for every GPU count.
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
    """All-gather the row shards into the full [n_total, D] matrix on every rank.

    Works on NCCL (CUDA tensors) and gloo (CPU tensors, used by the CPU tests)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = shard_rows(n_total, world)
    r0, r1 = shard_bounds(n_total, world, rank)
    if x_local.shape[0] != r1 - r0:
        raise ValueError(f"rank {rank} holds {x_local.shape[0]} rows, its shard is [{r0}, {r1})")
    d = x_local.shape[1]
    full = torch.empty((world * per, d), dtype=x_local.dtype, device=x_local.device)
    if x_local.shape[0] == per:
        mine = x_local.contiguous()
    else:
        mine = torch.zeros((per, d), dtype=x_local.dtype, device=x_local.device)
        mine[: x_local.shape[0]] = x_local
    dist.all_gather_into_tensor(full, mine, group=group)
    # only the last non-empty shard is short, so the first n_total rows are exactly the
    # global matrix in row order
    return full[:n_total]


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
    if r1 == r0:
        e = dict(device=x_local.device)
        return torch.empty((0, k), dtype=torch.int32, **e), torch.empty((0, k), dtype=torch.float64, **e)
    q = full[r0:r1]
    if pipeline is None:
        pipeline = KnnPipeline(r1 - r0, n_total, full.shape[1], k, metric, form, same=False,
                               device=x_local.device)
    return pipeline.run(q, full)
