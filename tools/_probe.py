import sys, torch
sys.path.insert(0, ".")
from groopm_b200 import engine, synth
n, d, k = 1000000, 30, 32
xd = torch.from_numpy(synth.cfg_knn_rows(n, d, 5)).cuda()
pipe = engine.KnnPipeline(n, n, d, k, "euclidean", same=True)
for _ in range(2):
    pipe.run(xd, xd)
st = pipe.prune_stats()
print({k: round(v) for k, v in st["probes"].items()})
