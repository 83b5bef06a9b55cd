"""Cycle probes of the pruned scan (needs a library built with GMD_NVCC_FLAGS=-DGMD_TC_PROBES):
    GMD_NVCC_FLAGS=-DGMD_TC_PROBES python -m groopm_b200.build --force && python tools/probe_prune.py [N D k]
Counters of thread 0 of every CTA (warp 0, lane 0), averaged over the CTAs."""
import sys
import torch
sys.path.insert(0, ".")
from groopm_b200 import engine, synth
n, d, k = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (1_000_000, 30, 32)
xd = torch.from_numpy(synth.cfg_knn_rows(n, d, {30: 5, 10: 2}.get(d, 4))).cuda()
pipe = engine.KnnPipeline(n, n, d, k, same=True)
pipe.run(xd, xd)
st = pipe.prune_stats()
print({k_: (round(v, 1) if isinstance(v, float) else v) for k_, v in st.items() if k_ != "probes"})
print({k_: round(v, 1) for k_, v in st["probes"].items()})
