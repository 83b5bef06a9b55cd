"""(Not a pytest module.)  BASELINE configs[0]: 10 k contigs x 3 samples through the distance
calls of core clustering (cluster.py:430-590) and bin merging (refine.py:646-665, 1490-1503):
drop-ins on the GPU vs the oracle (the reference's scipy expressions) on this host."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import oracle
from groopm_b200 import dropin, synth

c = synth.cfg_core(10000, 3, 1)
tcp, kpc = c["transformedCP"], c["kmerPCs"]
dropin.two_way_contraction_step(tcp[:2000], kpc[:2000])          # warm-up
t0 = time.perf_counter(); got = dropin.two_way_contraction_step(tcp, kpc); t_gpu = time.perf_counter() - t0
t0 = time.perf_counter(); want = oracle.two_way_contraction_step(tcp, kpc); t_cpu = time.perf_counter() - t0
same = (got["c_radius"] == want["c_radius"] and got["k_radius"] == want["k_radius"]
        and got["k_noise"] == want["k_noise"] and np.allclose(got["k_dat"], want["k_dat"], rtol=1e-12, atol=1e-12))
print(f"twoWayContraction distance step, n=10000: gmdist {t_gpu*1e3:.1f} ms, oracle {t_cpu:.2f} s, x{t_cpu/t_gpu:.0f}, same={same}")
lab = np.random.default_rng(0).integers(0, 64, 10000)
bins = [kpc[lab == b] for b in range(64)]
t0 = time.perf_counter(); a = [dropin.k_dist(b) for b in bins]; t_gpu = time.perf_counter() - t0
t0 = time.perf_counter(); b_ = [oracle.k_dist(b) for b in bins]; t_cpu = time.perf_counter() - t0
print(f"kDist over 64 bins: gmdist {t_gpu*1e3:.1f} ms, oracle {t_cpu*1e3:.1f} ms, same={a == b_}")
rows = [np.flatnonzero(lab == b) for b in range(64)]
dropin.k_dist_many(kpc, rows[:2])
t0 = time.perf_counter(); m = dropin.k_dist_many(kpc, rows); t_b = time.perf_counter() - t0
print(f"kDist over 64 bins, one launch (k_dist_many): {t_b*1e3:.2f} ms, same={m == b_}")
