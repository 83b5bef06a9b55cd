"""Tensor-core scan: parity against the FP32 CUDA-core path and timing.  python tools/tc_check.py"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from groopm_b200 import _lib, engine, synth

def run(n, d, k, nq=None):
    x = synth.cfg_knn_rows(n, d, 5)
    xd = torch.from_numpy(x).cuda()
    qd = xd if nq is None else xd[:nq].contiguous()
    res = {}
    import os
    for name, form in (("fp32", None), ("tc", _lib.FORM_TC)):
        os.environ["GMD_TC"] = "0" if form is None else "1"
        pipe = engine.KnnPipeline(qd.shape[0], n, d, k, "euclidean", form, same=nq is None)
        pipe.run(qd, xd); torch.cuda.synchronize()
        best = 1e9
        for r in range(3):
            pipe.reset(); pipe.col_mean(xd); pipe.prep_candidates(xd); pipe.prep_queries(qd)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e2 = torch.cuda.Event(enable_timing=True)
            e0.record(); pipe.scan(); e1.record()
            pipe.refine(qd, xd); pipe.exact_rows(qd, xd); e2.record(); torch.cuda.synchronize()
            if e0.elapsed_time(e1) < best:
                best, tail = e0.elapsed_time(e1), e1.elapsed_time(e2)
        res[name] = (pipe.out_idx.cpu().numpy().copy(), pipe.out_dist.cpu().numpy().copy(), best, pipe.exact_row_count(), tail)
    same = np.array_equal(res["fp32"][0], res["tc"][0]) and np.array_equal(res["fp32"][1], res["tc"][1])
    pairs = qd.shape[0] * n
    print(f"N={n} D={d} k={k} nq={qd.shape[0]}: fp32 scan {res['fp32'][2]:.2f} ms ({pairs/res['fp32'][2]/1e9:.2f} Tpair/s, exact rows {res['fp32'][3]}, refine+exact {res['fp32'][4]:.2f} ms) | "
          f"tc scan {res['tc'][2]:.2f} ms ({pairs/res['tc'][2]/1e9:.2f} Tpair/s, exact rows {res['tc'][3]}, refine+exact {res['tc'][4]:.2f} ms) | identical results: {same}", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[1].isdigit():
        a = [int(v) for v in sys.argv[1:] if v.isdigit()]
        for i in range(0, len(a) - 2, 3):
            run(a[i], a[i + 1], a[i + 2])
        sys.exit(0)
    run(20_000, 10, 16)
    run(100_000, 10, 16)
    run(250_000, 30, 32)
    if "--big" in sys.argv:
        run(1_000_000, 30, 32)
