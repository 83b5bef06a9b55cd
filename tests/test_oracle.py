"""CPU: the oracle against its committed golden vectors and analytic known answers."""
import hashlib

import numpy as np
import pytest
from scipy.spatial.distance import cdist, pdist, squareform

import oracle
from groopm_b200 import synth


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_inputs_are_reproducible(golden):
    g = golden("cfg2_knn_20k_d10_k16.npz")
    assert sha(synth.cfg_knn_rows(20000, 10, 2)) == str(g["x_sha"])
    g3 = golden("cfg3_bmu_20k_64x64.npz")
    block, nodes = synth.cfg_bmu(20000, 64, 3)
    assert sha(block) == str(g3["block_sha"]) and sha(nodes) == str(g3["nodes_sha"])


def test_knn_matches_golden(golden):
    g = golden("cfg2_knn_20k_d10_k16.npz")
    x = synth.cfg_knn_rows(20000, 10, 2)
    rows = g["rows"][:128]
    for m in ("euclidean", "cityblock"):
        d, i = oracle.knn(x[rows], x, 16, m)
        assert np.array_equal(i, g[f"idx_{m}"][:128])
        assert np.array_equal(d, g[f"dist_{m}"][:128])


def test_bmu_matches_golden_and_literal_loop(golden):
    g = golden("cfg3_bmu_20k_64x64.npz")
    block, nodes = synth.cfg_bmu(20000, 64, 3)
    flat = nodes.reshape(-1, 4)
    loc = oracle.bmu_batch(flat, block[:3000])
    assert np.array_equal(loc, g["loc"][:3000])
    # the batched restatement equals the reference's literal per-vector call
    for i in range(0, 3000, 150):
        r, c = oracle.best_match(flat, block[i], 64, 64)
        assert r * 64 + c == loc[i]  # square mesh: col = loc - row*rows is the true column


def test_kth_radius_matches_golden(golden):
    g = golden("cfg1_core.npz")
    sub = np.arange(0, 10000, 5)  # 2000 rows: the same call shape, seconds on CPU
    cw = oracle.whiten(g["tcp"])[sub]
    dm = oracle.allpairs_cityblock(cw)
    radius, kth = oracle.kth_neighbour_radius(dm, 10)
    assert kth.shape == (2000,) and radius == np.median(kth)
    assert np.all(kth >= g["c_kth"][sub])  # fewer candidates can only push the k-th distance out
    assert sha(g["tcp"]) == str(g["tcp_sha"])


# ---------------------------------------------------------------- analytic known answers

def test_lattice_distances_are_exact():
    g = np.stack(np.meshgrid(np.arange(6.0), np.arange(6.0), np.arange(6.0), indexing="ij"), -1).reshape(-1, 3)
    d, i = oracle.knn(g, g, 7, "cityblock")
    assert np.array_equal(i[:, 0], np.arange(len(g))) and np.all(d[:, 0] == 0)
    inner = np.all((g > 0) & (g < 5), axis=1)
    assert np.all(d[inner, 1:7] == 1.0)
    d2, _ = oracle.knn(g, g, 7, "euclidean")
    assert np.all(d2[inner, 1:7] == 1.0)


def test_duplicates_pin_lowest_index_rule():
    rng = np.random.default_rng(0)
    nodes = rng.random((50, 4))
    nodes[30] = nodes[7]
    nodes[41] = nodes[7]
    x = nodes[[7, 30, 41]] + 0.0
    assert np.array_equal(oracle.bmu_batch(nodes, x), [7, 7, 7])
    d, i = oracle.knn(x, nodes, 3)
    assert np.array_equal(i, [[7, 30, 41]] * 3) and np.all(d == 0)


def test_squareform_index_identity():
    x = np.random.default_rng(1).random((9, 3))
    p = pdist(x, "cityblock")
    sq = squareform(p)
    for idx in range(len(p)):
        a, b = oracle.small2indices(idx, 8)  # mstore.py:1880 is called with side = n - 1
        assert sq[a, b] == p[idx]


def test_bmu_on_mesh_whose_nodes_are_the_queries():
    nodes = synth.som_nodes(12, 4, 9).reshape(-1, 4)
    assert np.array_equal(oracle.bmu_batch(nodes, nodes), np.arange(len(nodes)))


def test_radius_quirk_drops_lowest_index_neighbour():
    row = np.array([0.5, 0.0, 0.2, 3.0, 0.1])
    assert np.array_equal(oracle.radius_neighbours(row, 0.5, 10), [1, 2, 4])  # drops index 0, not self (1)
    assert oracle.radius_neighbours(np.array([0.0, 9.0]), 0.5, 10) is None


def test_kdtree_equals_bruteforce():
    x = np.random.default_rng(2).random((300, 3))
    d, i = oracle.kdtree_knn(x, 6)
    bd, bi = oracle.knn(x, x, 6)
    assert oracle.knn_tie_aware_equal(x, x, i, bi, bd) == 0
    assert np.allclose(d, bd, rtol=0, atol=1e-15)


def test_tie_aware_compare_flags_real_errors():
    x = np.random.default_rng(3).random((100, 5))
    d, i = oracle.knn(x, x, 5)
    wrong = i.copy()
    wrong[3, 4] = (wrong[3, 4] + 1) % 100 if (wrong[3, 4] + 1) % 100 not in wrong[3] else 99
    assert oracle.knn_tie_aware_equal(x, x, i, i, d) == 0
    assert oracle.knn_tie_aware_equal(x, x, wrong, i, d) == 1


def test_medians():
    rng = np.random.default_rng(4)
    a, b = rng.random((31, 6)), rng.random((18, 6))
    assert oracle.k_dist_between(a, b) == np.median(cdist(a, b, "cityblock"))
    assert oracle.k_dist(a) == np.median(pdist(a, "cityblock"))
    assert oracle.k_dist(a[:1]) is None


def test_cfg1_full_matches_golden(golden):
    """The reference's literal call shape at cfg1 size: two 10k x 10k float64 matrices."""
    g = golden("cfg1_core.npz")
    assert sha(g["kpc"]) == str(g["kpc_sha"])
    assert sha(synth.cfg_core(10000, 3, 1)["kmerPCs"]) == str(g["kpc_sha"])  # reproducible in this image
    dm = oracle.allpairs_cityblock(g["kpc"])
    radius, kth = oracle.kth_neighbour_radius(dm, int(g["eps"]))
    assert radius == float(g["k_radius"]) and np.array_equal(kth, g["k_kth"])
    counts = (dm <= radius).sum(axis=1)
    assert np.array_equal(counts, g["k_counts"])
