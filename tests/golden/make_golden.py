"""Generates tests/golden/*.npz from the oracle (scipy/numpy float64, the reference's
own arithmetic) on the seeded synthetic profiles of SURVEY.md section 8d.

Run in the build container:  python tests/golden/make_golden.py
The reference ships no fixtures, so these pin the oracle's
outputs as produced by the scipy/numpy versions recorded in each file.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np
import scipy

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from groopm_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent
VERS = np.array([f"numpy {np.__version__}", f"scipy {scipy.__version__}"])


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def csr_rows(dm, radius):
    counts = (dm <= radius).sum(axis=1).astype(np.int32)
    idxsum = np.array([np.flatnonzero(r <= radius).sum() for r in dm], dtype=np.int64)
    return counts, idxsum


def cfg1():
    c = synth.cfg_core(10000, 3, 1)
    tcp, kpc = c["transformedCP"], c["kmerPCs"]
    eps = 10
    cw = oracle.whiten(tcp)
    c_dm = oracle.allpairs_cityblock(cw)
    c_radius, c_kth = oracle.kth_neighbour_radius(c_dm, eps)
    k_dm = oracle.allpairs_cityblock(kpc)
    k_radius, k_kth = oracle.kth_neighbour_radius(k_dm, eps)
    c_counts, c_idxsum = csr_rows(c_dm, c_radius)
    k_counts, k_idxsum = csr_rows(k_dm, k_radius)
    step = oracle.two_way_contraction_step(tcp, kpc)
    # bins = the generator's clusters; centroids as in findMergeGroups (medians)
    lab = np.random.default_rng(1).integers(0, max(8, 10000 // 2000), size=10000)
    rng = np.random.default_rng(77)
    lab64 = rng.integers(0, 64, size=10000)  # 64 pseudo-bins for the KDTree shape
    cov_med = np.array([np.median(tcp[lab64 == b], axis=0) for b in range(64)])
    k_med = np.array([np.median(kpc[lab64 == b], axis=0) for b in range(64)])
    cov_w = oracle.whiten(cov_med)
    kd_c_d, kd_c_i = oracle.kdtree_knn(cov_w, 6)
    kd_k_d, kd_k_i = oracle.kdtree_knn(k_med, 6)
    bins = [np.flatnonzero(lab == b) for b in range(lab.max() + 1)]
    kdist = np.array([oracle.k_dist(kpc[r]) for r in bins])
    pairs = [(0, 1), (1, 2), (2, 3), (0, 4), (3, 4), (2, 2)]
    kbetween = np.array([oracle.k_dist_between(kpc[bins[a]], kpc[bins[b]]) for a, b in pairs])
    np.savez_compressed(
        OUT / "cfg1_core.npz", versions=VERS, tcp=tcp, kpc=kpc, tcp_sha=digest(tcp), kpc_sha=digest(kpc), eps=eps,
        c_radius=c_radius, k_radius=k_radius, c_kth=c_kth, k_kth=k_kth, c_counts=c_counts,
        c_idxsum=c_idxsum, k_counts=k_counts, k_idxsum=k_idxsum,
        new_k_head=step["k_dat"][:256], new_c_head=step["c_dat"][:256],
        new_k_sum=step["k_dat"].sum(axis=0), new_c_sum=step["c_dat"].sum(axis=0),
        k_noise=np.array(sorted(step["k_noise"]), dtype=np.int32),
        c_noise=np.array(sorted(step["c_noise"]), dtype=np.int32),
        kd_c_i=kd_c_i.astype(np.int32), kd_c_d=kd_c_d, kd_k_i=kd_k_i.astype(np.int32), kd_k_d=kd_k_d,
        kdist=kdist, kbetween=kbetween, pairs=np.array(pairs), lab64=lab64.astype(np.int8))
    print("cfg1", c_radius, k_radius, kdist[:3], kbetween[:3])


def knn_fixture(name, x, k, rows, metrics=("euclidean",)):
    out = dict(versions=VERS, x_sha=digest(x), rows=rows.astype(np.int32), k=k)
    for m in metrics:
        d, i = oracle.knn(x[rows], x, k, m)
        out[f"idx_{m}"] = i.astype(np.int32)
        out[f"dist_{m}"] = d
    np.savez_compressed(OUT / name, **out)
    print(name, out[f"dist_{metrics[0]}"][0, :3])


def cfg2():
    x = synth.cfg_knn_rows(20000, 10, 2)
    rows = np.random.default_rng(22).choice(20000, 1024, replace=False)
    knn_fixture("cfg2_knn_20k_d10_k16.npz", x, 16, np.sort(rows), ("euclidean", "cityblock"))


def cfg3():
    block, nodes = synth.cfg_bmu(20000, 64, 3)
    flat = nodes.reshape(-1, 4)
    loc = oracle.bmu_batch(flat, block)
    binass = (np.arange(64 * 64).reshape(64, 64) % 37).astype(np.int32)
    bids = oracle.classify_contigs(flat, block, binass, 64, 64)
    np.savez_compressed(OUT / "cfg3_bmu_20k_64x64.npz", versions=VERS, block_sha=digest(block),
                        nodes_sha=digest(nodes), loc=loc.astype(np.int32), bids=bids.astype(np.int32))
    print("cfg3", loc[:5])


def cfg4():
    sigs = synth.quantise(synth.kmer_sigs(4000, 4), 34)
    rows = np.sort(np.random.default_rng(44).choice(4000, 512, replace=False))
    knn_fixture("cfg4_knn_4k_d136_k16.npz", sigs, 16, rows, ("euclidean", "cityblock"))
    lab = np.random.default_rng(4).integers(0, 8, size=4000)
    pairs = [(0, 1), (2, 5), (7, 7)]
    med = np.array([oracle.k_dist_between(sigs[lab == a], sigs[lab == b]) for a, b in pairs])
    np.savez_compressed(OUT / "cfg4_median_d136.npz", versions=VERS, x_sha=digest(sigs), pairs=np.array(pairs),
                        med=med, lab=lab.astype(np.int8))
    print("cfg4", med)


def cfg5():
    x = synth.cfg_knn_rows(30000, 30, 5)
    rows = np.sort(np.random.default_rng(55).choice(30000, 512, replace=False))
    knn_fixture("cfg5_knn_30k_d30_k32.npz", x, 32, rows)


if __name__ == "__main__":
    cfg1(); cfg2(); cfg3(); cfg4(); cfg5()
