"""gmd_knn_host_f64 (pure C-ABI, pageable numpy buffers) timed call by call: python tools/cabi_time.py [N]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from groopm_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d, k = 30, 32
x = synth.cfg_knn_rows(n, d, 5)
outs = (np.zeros((n, k), dtype=np.int32), np.zeros((n, k), dtype=np.float64))
for i in range(6):
    t0 = time.perf_counter()
    engine.knn_host(x, None, k, "euclidean", out=outs)
    print("call %d: %.1f ms" % (i, (time.perf_counter() - t0) * 1e3), flush=True)
import torch
xs = torch.from_numpy(x).pin_memory().numpy()
for i in range(3):
    t0 = time.perf_counter()
    engine.knn(xs, k=k)
    print("engine.knn(numpy) %d: %.1f ms" % (i, (time.perf_counter() - t0) * 1e3), flush=True)
for i in range(3):
    t0 = time.perf_counter()
    engine.knn_host(x, None, k, "euclidean", out=outs)
    print("after torch, call %d: %.1f ms" % (i, (time.perf_counter() - t0) * 1e3), flush=True)
