"""Developer perf probe on a GPU box: per-stage CUDA-event timings of the pipeline."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from groopm_b200 import engine, synth

def ev():
    return torch.cuda.Event(enable_timing=True)

def stage_times(pipe, q, c, reps=3):
    names = ["col_mean", "prep_c", "prep_q", "scan", "refine", "exact"]
    best = {n: 1e30 for n in names}
    tot = 1e30
    for r in range(reps):
        evs = [ev() for _ in range(7)]
        pipe.reset()
        evs[0].record(); pipe.col_mean(c)
        evs[1].record(); pipe.prep_candidates(c)
        evs[2].record(); pipe.prep_queries(q)
        evs[3].record(); pipe.scan()
        evs[4].record(); pipe.refine(q, c)
        evs[5].record(); pipe.exact_rows(q, c)
        evs[6].record(); torch.cuda.synchronize()
        for i, n in enumerate(names):
            best[n] = min(best[n], evs[i].elapsed_time(evs[i + 1]))
        tot = min(tot, evs[0].elapsed_time(evs[6]))
    return best, tot

def run(name, q, c, k, metric, form=None, peak=None):
    same = c is None
    qd = torch.from_numpy(q).cuda()
    cd = qd if same else torch.from_numpy(c).cuda()
    pipe = engine.KnnPipeline(qd.shape[0], cd.shape[0], qd.shape[1], k, metric, form, same=same)
    best, tot = stage_times(pipe, qd, cd)
    pairs = qd.shape[0] * cd.shape[0]
    d = qd.shape[1]
    tf = 2 * d * pairs / (best["scan"] * 1e-3) / 1e12
    print(f"{name}: form={pipe.form} kk={pipe.kk} nsplit={pipe.nsplit} exact_rows={pipe.exact_row_count()} "
          f"total={tot:.2f}ms scan={best['scan']:.2f}ms ({pairs/best['scan']/1e6:.1f} Gpair/s, {tf:.1f} TFLOP/s"
          f"{'' if peak is None else f', {tf/peak*100:.1f}% of probe peak'}) "
          + " ".join(f"{n}={v:.2f}" for n, v in best.items() if n != "scan"), flush=True)

if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg2", "cfg3", "cfg5s"]
    print(torch.cuda.get_device_name(0))
    peak = engine.fp32_peak_tflops()
    print("fp32 probe peak TFLOP/s:", peak, flush=True)
    if "cfg2" in which:
        x = synth.cfg_knn_rows(100_000, 10, 2)
        run("cfg2 100k x10 L2 k16", x, None, 16, "euclidean", peak=peak)
        run("cfg2 100k x10 L1 k16", x, None, 16, "cityblock", peak=peak / 2)
    if "cfg2split" in which:
        import os
        x = synth.cfg_knn_rows(100_000, 10, 2)
        for ns in (1, 2, 3, 4, 6):
            os.environ["GMD_NSPLIT"] = str(ns)
            run(f"cfg2 100k x10 L2 k16 nsplit={ns}", x, None, 16, "euclidean", peak=peak)
        x = synth.cfg_knn_rows(250_000, 30, 5)
        for ns in (1, 2, 3):
            os.environ["GMD_NSPLIT"] = str(ns)
            run(f"cfg5 250k x30 L2 k32 nsplit={ns}", x, None, 32, "euclidean", peak=peak)
        del os.environ["GMD_NSPLIT"]
    if "cfg3" in which:
        b, n = synth.cfg_bmu(1_000_000, 64)
        run("cfg3 1M x4 BMU 4096", b, n.reshape(-1, 4), 1, "euclidean", peak=peak)
    if "cfg4" in which:
        x = synth.quantise(synth.kmer_sigs(100_000, 4), 34)
        run("cfg4 100k x136 L2 k16", x, None, 16, "euclidean", peak=peak)
        run("cfg4 100k x136 L1 k16", x, None, 16, "cityblock", peak=peak / 2)
    if "cfg5s" in which:
        x = synth.cfg_knn_rows(250_000, 30, 5)
        run("cfg5 250k x30 L2 k32", x, None, 32, "euclidean", peak=peak)
    if "cfg5" in which:
        x = synth.cfg_knn_rows(1_000_000, 30, 5)
        run("cfg5 1M x30 L2 k32", x, None, 32, "euclidean", peak=peak)
