"""One short pipeline run for ncu captures: python tools/ncu_case.py N D k [metric] [nq]
(engine defaults: the pruned tensor-core scan for large Euclidean problems; GMD_TC_PRUNE=0 and
GMD_TC_ORDER=1 in the environment select the full sweep / its locality-ordered variant)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from groopm_b200 import engine, synth
n, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
metric = sys.argv[4] if len(sys.argv) > 4 else "euclidean"
nq = int(sys.argv[5]) if len(sys.argv) > 5 else n
x = synth.cfg_knn_rows(n, d, {30: 5, 10: 2}.get(d, 4)) if d != 136 else synth.quantise(synth.kmer_sigs(n, 4), 34)
xd = torch.from_numpy(x).cuda()
qd = xd if nq == n else xd[:nq]
pipe = engine.KnnPipeline(nq, n, d, k, metric, same=(nq == n))
for _ in range(2):
    pipe.run(qd, xd)
torch.cuda.synchronize()
print("ok", "pruned" if pipe.prune else ("ordered" if pipe.order else "plain"), pipe.exact_row_count(), pipe.nsplit)
