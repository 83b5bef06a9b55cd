"""torchrun worker of tests/test_dist_gpu.py: every rank computes its shard of the exact kNN graph with
``dist.sharded_knn`` (NCCL all-gather of the rows, then the engine's default pipeline on the rank's run of the
rows) and checks it against the oracle on every row of the shard and, on rank 0, against the single-GPU
``engine.knn`` rows, bit for bit.  Prints ``RANK r OK`` on success."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as tdist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle  # noqa: E402
from groopm_b200 import dist, engine, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    tdist.init_process_group("nccl", device_id=dev)
    try:
        for n, d, k, force_prune in ((70_001, 10, 16, None), (9_000, 30, 32, True), (150, 4, 5, None)):
            x = synth.cfg_knn_rows(n, d, 11)
            r0, r1 = dist.shard_bounds(n, world, rank)        # ragged: the last rank's shard is short (or empty)
            xl = torch.from_numpy(x[r0:r1].copy()).to(dev)
            pipe = None
            if force_prune and r1 > r0:
                from groopm_b200 import _lib
                pipe = engine.KnnPipeline(r1 - r0, n, d, k, "euclidean", _lib.FORM_TC, same=False, device=dev, prune=True)
            idx, dd = dist.sharded_knn(xl, n, k, pipeline=pipe)
            assert idx.shape == (r1 - r0, k)
            if r1 > r0:
                rows = np.arange(r0, r1) if r1 - r0 <= 3000 else np.sort(
                    np.random.default_rng(rank).choice(np.arange(r0, r1), 3000, replace=False))
                od, oi = oracle.knn(x[rows], x, k)
                assert np.array_equal(idx[rows - r0].cpu().numpy(), oi), f"rank {rank}: indices differ (n={n})"
                assert np.array_equal(dd[rows - r0].cpu().numpy(), od), f"rank {rank}: distances differ (n={n})"
            if rank == 0:
                gi, gd = engine.knn(torch.from_numpy(x).to(dev), None, k=k)
                assert torch.equal(gi[r0:r1], idx) and torch.equal(gd[r0:r1], dd), "sharded rows != single-GPU rows"
        tdist.barrier()
        print(f"RANK {rank} OK", flush=True)
    finally:
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
