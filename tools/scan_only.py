"""Time ONLY the FP32 scan kernel (no refine): python tools/scan_only.py N D k [nq] [metric]"""
import sys
import torch
sys.path.insert(0, ".")
from groopm_b200 import engine, synth
n, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
nq = int(sys.argv[4]) if len(sys.argv) > 4 else n
metric = sys.argv[5] if len(sys.argv) > 5 else "euclidean"
x = synth.cfg_knn_rows(n, d, 5)
xd = torch.from_numpy(x).cuda()
qd = xd if nq == n else xd[:nq].contiguous()
pipe = engine.KnnPipeline(nq, n, d, k, metric, same=(nq == n))
pipe.reset(); pipe.col_mean(xd); pipe.prep_candidates(xd); pipe.prep_queries(qd)
best = 1e9
for r in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pipe.scan(); e1.record(); torch.cuda.synchronize()
    if r: best = min(best, e0.elapsed_time(e1))
print(f"scan only N={n} D={d} k={k} nq={nq} nsplit={pipe.nsplit}: {best:.3f} ms  {2*d*nq*n/best/1e9:.1f} TFLOP/s")
