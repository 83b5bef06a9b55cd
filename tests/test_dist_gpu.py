"""The row-sharded path on real GPUs (north_star: "index-exact" at 2/4/8 GPUs):
``dist.sharded_knn`` under torchrun on every visible GPU (at least 2), each rank's rows against the oracle
and against the single-GPU result.  Skipped on a one-GPU box; ``bench.py --gpus N`` verifies sampled rows of
every rank on every multi-GPU line as well (config.verified_rows)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_sharded_knn_equals_single_gpu_and_oracle(cuda):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env={**os.environ, "MASTER_ADDR": "127.0.0.1"})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for rank in range(world):
        assert f"RANK {rank} OK" in r.stdout
