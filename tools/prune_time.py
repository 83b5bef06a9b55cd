"""All-pairs (or row-sliced) kNN through the full sweep and through the pruned tensor-core scan:
whole-pipeline, staging and scan times, tiles visited, results compared bit for bit.

    python tools/prune_time.py N D K [slice_fraction] [variants...]
variants: plain order prune (default: plain prune)
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from groopm_b200 import _lib, engine, synth

n, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
frac = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
variants = sys.argv[5:] or ["plain", "prune"]
seed = {30: 5, 10: 2}.get(d, 4)
x = synth.cfg_knn_rows(n, d, seed) if d != 136 else synth.quantise(synth.kmer_sigs(n, 4), 34)
xd = torch.from_numpy(x).cuda()
same = frac >= 1.0
nq = n if same else int(n * frac)
r0 = 0 if same else (n - nq) // 2
qd = xd if same else xd[r0:r0 + nq]
res = {}
for v in variants:
    pipe = engine.KnnPipeline(nq, n, d, k, "euclidean", _lib.FORM_TC, same=same, order=(v == "order"), prune=(v == "prune"))
    pipe.run(qd, xd)
    torch.cuda.synchronize()
    best = (1e9, 0, 0)
    for r in range(3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(); pipe.reset(); pipe.col_mean(xd); pipe.prep_candidates(xd); pipe.prep_queries(qd)
        ev[1].record(); pipe.scan(); ev[2].record(); pipe.refine(qd, xd); pipe.exact_rows(qd, xd); ev[3].record()
        torch.cuda.synchronize()
        t = ev[0].elapsed_time(ev[3])
        if t < best[0]:
            best = (t, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]))
    res[v] = (pipe.out_idx.clone(), pipe.out_dist.clone())
    extra = ""
    if v == "prune":
        st = pipe.prune_stats()
        extra = (f"; {st['groups']} groups, {st['candidate_tiles']} candidate tiles, visited "
                 f"{st['tiles_visited'] / max(1, st['query_tiles']):.1f} per query tile ({st['visited_fraction']:.3%}), "
                 f"{st['groups_surviving_per_query_tile']:.1f} groups surviving")
    print(f"N={n} Nq={nq} D={d} k={k} {v}: step {best[0]:.2f} ms (staging {best[1]:.2f}, scan {best[2]:.2f}, "
          f"finish {best[0] - best[1] - best[2]:.2f}), {nq * n / best[0] / 1e9:.2f} Tpair/s, "
          f"exact rows {pipe.exact_row_count()}{extra}", flush=True)
base = variants[0]
for v in variants[1:]:
    same_i = bool(torch.equal(res[base][0], res[v][0]))
    same_d = bool(torch.equal(res[base][1], res[v][1]))
    print(f"{v} == {base}: indices {same_i}, distances {same_d}")
    if not (same_i and same_d):
        bad = (res[base][0] != res[v][0]).any(dim=1).nonzero().flatten()
        print("  rows differing:", bad.numel(), bad[:10].tolist())
        sys.exit(1)
