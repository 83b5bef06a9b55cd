"""All-pairs kNN with and without the locality-ordered sweep: whole-pipeline and scan times, results compared.
    python tools/order_time.py N D K"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from groopm_b200 import _lib, engine, synth
n, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = synth.cfg_knn_rows(n, d, 5) if d != 136 else synth.quantise(synth.kmer_sigs(n, 4), 34)
xd = torch.from_numpy(x).cuda()
res = {}
for order in (False, True):
    pipe = engine.KnnPipeline(n, n, d, k, "euclidean", _lib.FORM_TC, same=True, order=order)
    pipe.run(xd, xd); torch.cuda.synchronize()
    best = (1e9, 0, 0)
    for r in range(2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(); pipe.reset(); pipe.col_mean(xd); pipe.prep_candidates(xd); pipe.prep_queries(xd)
        ev[1].record(); pipe.scan(); ev[2].record(); pipe.refine(xd, xd); pipe.exact_rows(xd, xd); ev[3].record()
        torch.cuda.synchronize()
        t = ev[0].elapsed_time(ev[3])
        if t < best[0]:
            best = (t, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]))
    res[order] = (pipe.out_idx.clone(), pipe.out_dist.clone())
    print(f"N={n} D={d} k={k} ordered={order}: step {best[0]:.2f} ms (staging {best[1]:.2f}, scan {best[2]:.2f}), "
          f"{n * n / best[0] / 1e9:.2f} Tpair/s, exact rows {pipe.exact_row_count()}", flush=True)
print("identical results:", bool(torch.equal(res[False][0], res[True][0]) and torch.equal(res[False][1], res[True][1])))
