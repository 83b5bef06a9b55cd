"""Time the tensor-core scan only (optionally GMD_TC_DBG=1: MMA + copy pipeline without selection)."""
import sys, os
import torch
sys.path.insert(0, ".")
from groopm_b200 import _lib, engine, synth
n, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = synth.cfg_knn_rows(n, d, 5)
xd = torch.from_numpy(x).cuda()
pipe = engine.KnnPipeline(n, n, d, k, "euclidean", _lib.FORM_TC, same=True)
pipe.reset(); pipe.col_mean(xd); pipe.prep_candidates(xd); pipe.prep_queries(xd)
for dbg in (sys.argv[4:] or ("0", "1", "2", "4")):
    os.environ["GMD_TC_DBG"] = dbg
    best = 1e9
    for r in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pipe.scan(); e1.record(); torch.cuda.synchronize()
        if r: best = min(best, e0.elapsed_time(e1))
    print(f"N={n} D={d} k={k} dbg={dbg}: {best:.2f} ms = {n*n/best/1e9:.2f} Tpair/s")
