"""Developer check on a GPU box: engine vs oracle on a handful of shapes."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from groopm_b200 import engine, synth

def check(n, d, k, metric, nq=None, seed=0, form=None):
    x = synth.cfg_knn_rows(n, d, seed + 2)
    q = x if nq is None else x[:nq].copy()
    t0 = time.time()
    idx, dist, info = engine.knn(q, None if nq is None else x, k=k, metric=metric, form=form, return_info=True)
    t1 = time.time()
    od, oi = oracle.knn(q, x, k, metric)
    bad = oracle.knn_tie_aware_equal(q, x, idx, oi, od, metric)
    rel = np.max(np.abs(dist - od) / np.maximum(od, 1e-300) * (od > 0))
    exact = int((idx == oi).all(axis=1).sum())
    print(f"n={n} d={d} k={k} {metric} nq={nq} form={info['form']} nsplit={info['nsplit']} "
          f"exact_rows={info['exact_rows']} bad={bad} identical_rows={exact}/{len(q)} maxrel={rel:.2e} t={t1-t0:.3f}s")
    return bad

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    print("fp32 peak TFLOP/s:", engine.fp32_peak_tflops())
    bad = 0
    bad += check(1000, 4, 1, "euclidean", nq=300)
    bad += check(5000, 10, 16, "euclidean")
    bad += check(5000, 10, 16, "cityblock")
    bad += check(5000, 3, 11, "cityblock")
    bad += check(5000, 3, 6, "euclidean")
    bad += check(20000, 30, 32, "euclidean", nq=2000)
    bad += check(20000, 30, 32, "euclidean", nq=2000, form=1)
    bad += check(3000, 13, 6, "euclidean")
    bad += check(70, 5, 6, "euclidean")
    print("TOTAL BAD", bad)
    sys.exit(1 if bad else 0)
