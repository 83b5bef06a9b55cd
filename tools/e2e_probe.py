"""Where the end-to-end call's time goes: python tools/e2e_probe.py [N] (1M x 30, k = 32 by default)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from groopm_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d, k = 30, 32
x = torch.from_numpy(synth.cfg_knn_rows(n, d, 5)).pin_memory()
xn = x.numpy()
def t(f, reps=8):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
print("H2D only            %.2f ms" % t(lambda: x.to("cuda", non_blocking=True)))
xd = x.cuda()
oi = torch.empty((n, k), dtype=torch.int32, device="cuda"); od = torch.empty((n, k), dtype=torch.float64, device="cuda")
def d2h():
    a = torch.empty(oi.shape, dtype=oi.dtype, pin_memory=True); b = torch.empty(od.shape, dtype=od.dtype, pin_memory=True)
    a.copy_(oi, non_blocking=True); b.copy_(od, non_blocking=True); torch.cuda.current_stream().synchronize()
print("D2H only (+pinned alloc) %.2f ms" % t(d2h))
pipe = engine.KnnPipeline(n, n, d, k, same=True)
print("device step         %.2f ms" % t(lambda: pipe.run(xd, xd)))
for runs in (2, 3, 4, 6, 8):
    engine.clear_cache()
    print("streamed runs=%d on device rows %.2f ms" % (runs, t(lambda: engine._knn_streamed(xd, k, "euclidean", xd.device, runs=runs, min_rows=0))))
engine.clear_cache()
print("engine.knn(numpy)   %.2f ms" % t(lambda: engine.knn(xn, k=k)))
