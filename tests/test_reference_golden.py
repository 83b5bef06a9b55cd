"""The oracle (CPU) and the CUDA path (GPU) against outputs of the REFERENCE ITSELF.

``tests/golden/ref_*.npz`` were produced by ``tests/golden/make_reference_golden.py``,
which executes the reference's own ``TorusMesh.bestMatch`` (torusMesh.py:123-130),
``SOM.classifyContig`` (som.py:572-575), ``ClusterEngine.twoWayContraction``
(cluster.py:411-600) and ``RefineEngine.kDist`` / ``kDistBetweenBins`` / ``kDistMergedBins``
(refine.py:1490-1503) and ``getEvenlySpacedPtsZ`` / ``cDist`` / ``cDistBetweenBins`` / ``cDistMergedBins``
(refine.py:1505-1583) from /root/reference under Python 3.  Inputs travel inside the
fixtures, so nothing here depends on regenerating them or on /root/reference."""
import numpy as np
import pytest

import oracle


# ------------------------------------------------------------------ CPU: oracle == reference

def test_oracle_bestmatch_equals_reference(golden):
    g = golden("ref_bestmatch.npz")
    nodes, x, rc = g["nodes"], g["x"], g["rc"]
    flat = nodes.reshape(-1, 4)
    loc = oracle.bmu_batch(flat, x)
    assert np.array_equal(loc, rc[:, 0] * 24 + rc[:, 1])     # square mesh: [row, col] <-> flat index
    for i in range(0, len(x), 97):
        assert oracle.best_match(flat, x[i], 24, 24) == rc[i].tolist()
    # duplicates and masked (-1) nodes: first minimum wins, masked nodes never do
    assert rc[50].tolist() == [2, 9] and rc[51].tolist() == [0, 0]
    assert not np.any((rc[52:, 0] == 3) & (rc[52:, 1] == 5))  # rows 52.. are data in [0, 1]
    # non-square mesh: the reference's col = loc - row*rows (torusMesh.py:128-129)
    flat2 = g["nodes2"].reshape(-1, 3)
    got = np.array([oracle.best_match(flat2, v, 12, 20) for v in g["x2"]])
    assert np.array_equal(got, g["rc2"])
    assert np.array_equal(oracle.classify_contigs(flat, x, g["bin_assignments"], 24, 24), g["bids"])


def test_oracle_contraction_equals_reference(golden):
    g = golden("ref_contraction.npz")
    for tag in "abc":
        got = oracle.two_way_contraction_step(g[f"{tag}_transformedCP"], g[f"{tag}_kmerPCs"])
        assert got["eps"] == int(g[f"{tag}_eps_neighbours"])
        assert got["c_radius"] == float(g[f"{tag}_c_radius"]) and got["k_radius"] == float(g[f"{tag}_k_radius"])
        assert np.array_equal(got["k_dat"], g[f"{tag}_k_dat"]) and np.array_equal(got["c_dat"], g[f"{tag}_c_dat"])
        assert sorted(got["k_noise"]) == g[f"{tag}_k_noise"].tolist()
        assert sorted(got["c_noise"]) == g[f"{tag}_c_noise"].tolist()


def _kdist_cases(g):
    rows = [g[f"rows{i}"] for i in range(int(g["n_sets"]))]
    return g["kmerPCs"], rows


def test_oracle_kdist_equals_reference(golden):
    g = golden("ref_kdist.npz")
    pcs, rows = _kdist_cases(g)
    for r, want in zip(rows, g["k_dist"]):
        got = oracle.k_dist(pcs[r])
        assert (got is None and np.isnan(want)) or got == want
    for (a, b), want in zip(g["pairs"], g["k_dist_between"]):
        assert oracle.k_dist_between(pcs[rows[a]], pcs[rows[b]]) == want
    for (a, b), want in zip(g["pairs"], g["k_dist_merged"]):
        assert oracle.k_dist(pcs[np.concatenate((rows[a], rows[b]))]) == want


# ------------------------------------------------------------------ GPU: engine == reference

@pytest.mark.gpu
def test_engine_bestmatch_equals_reference(cuda, golden):
    from groopm_b200 import dropin, engine

    g = golden("ref_bestmatch.npz")
    nodes, x, rc = g["nodes"], g["x"], g["rc"]
    flat = nodes.reshape(-1, 4)
    row, col = dropin.best_match_many(flat, x, 24, 24)
    assert np.array_equal(np.stack([row, col], 1), rc)
    for i in (0, 50, 51, 1234):
        assert dropin.best_match(flat, x[i], 24, 24) == rc[i].tolist()
    flat2 = g["nodes2"].reshape(-1, 3)
    row2, col2 = dropin.best_match_many(flat2, g["x2"], 12, 20)
    assert np.array_equal(np.stack([row2, col2], 1), g["rc2"])
    assert np.array_equal(dropin.classify_contigs(flat, x, g["bin_assignments"], 24, 24), g["bids"])
    assert np.array_equal(engine.bmu(flat, x), rc[:, 0] * 24 + rc[:, 1])


@pytest.mark.gpu
def test_engine_contraction_equals_reference(cuda, golden):
    from groopm_b200 import dropin

    g = golden("ref_contraction.npz")
    for tag in "abc":
        got = dropin.two_way_contraction_step(g[f"{tag}_transformedCP"], g[f"{tag}_kmerPCs"])
        # radii, neighbour lists and noise sets are distance work: exact
        assert got["c_radius"] == float(g[f"{tag}_c_radius"]) and got["k_radius"] == float(g[f"{tag}_k_radius"])
        assert sorted(got["k_noise"]) == g[f"{tag}_k_noise"].tolist()
        assert sorted(got["c_noise"]) == g[f"{tag}_c_noise"].tolist()
        # the weighted moves are host numpy sums in a different association order
        np.testing.assert_allclose(got["k_dat"], g[f"{tag}_k_dat"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(got["c_dat"], g[f"{tag}_c_dat"], rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
def test_engine_kdist_equals_reference(cuda, golden):
    from groopm_b200 import dropin

    g = golden("ref_kdist.npz")
    pcs, rows = _kdist_cases(g)
    for r, want in zip(rows, g["k_dist"]):
        got = dropin.k_dist(pcs[r])
        assert (got is None and np.isnan(want)) or got == want
    for (a, b), want in zip(g["pairs"], g["k_dist_between"]):
        assert dropin.k_dist_between(pcs[rows[a]], pcs[rows[b]]) == want
    for (a, b), want in zip(g["pairs"], g["k_dist_merged"]):
        assert dropin.k_dist(pcs[np.concatenate((rows[a], rows[b]))]) == want


# ------------------------------------------------------------------ SOM training (row f1)

def _som_cases(g):
    full = dict(nodes=g["w0"], train=g["train"], weights=None, iterations=int(g["iterations"]),
                vector_subset=int(g["vector_subset"]), influence_rate=0.4, mask=None, radius=0.0,
                seed=int(g["shuffle_seed"]), want=g["weights"])
    sub = dict(nodes=None, train=g["sub_block"], weights=g["sub_w0"], iterations=int(g["sub_iterations"]),
               vector_subset=1000, influence_rate=float(g["sub_influence"]), mask=g["sub_mask"],
               radius=float(g["sub_radius"]), seed=9, want=g["sub_weights"])
    return full, sub


def test_oracle_som_train_equals_reference(golden):
    """SOM.train run by the reference itself (som.py:214-392): the whole-mesh call of
    refine.py:1066-1071 (subset shuffles from numpy's global generator) and the masked
    sub-mesh call of refine.py:1168-1173 (caller-owned non-square weights, mask, radius)."""
    for c in _som_cases(golden("ref_som_train.npz")):
        np.random.seed(c["seed"])
        w = oracle.som_train(c["nodes"], c["train"], weights=c["weights"], iterations=c["iterations"],
                             vector_subset=c["vector_subset"], influence_rate=c["influence_rate"], mask=c["mask"],
                             radius=c["radius"])
        assert np.array_equal(w, c["want"])
    assert (c["want"] != c["weights"]).any() and np.array_equal(c["want"][c["mask"] == 0], c["weights"][c["mask"] == 0])


def test_host_stamp_equals_oracle():
    from groopm_b200 import engine

    for rd in np.linspace(2.0, 40.0, 191):
        for ir in (0.4, 0.15, 0.1):
            assert np.array_equal(engine.som_stamp(rd, ir), oracle.som_stamp(rd, ir))


@pytest.mark.gpu
def test_engine_som_train_equals_reference(cuda, golden):
    from groopm_b200 import dropin

    for c in _som_cases(golden("ref_som_train.npz")):
        np.random.seed(c["seed"])
        mask = c["mask"]
        if mask is not None:   # the reference passes a dict of points (refine.py:1160-1172)
            mask = {(int(r), int(col)): None for r, col in np.argwhere(mask != 0)}
        w = dropin.som_train(c["nodes"], c["train"], weights=c["weights"], iterations=c["iterations"],
                             vectorSubSet=c["vector_subset"], influenceRate=c["influence_rate"], mask=mask,
                             radius=c["radius"])
        assert np.array_equal(w, c["want"])            # bit-identical float64 weights


@pytest.mark.gpu
def test_engine_som_train_equals_oracle_larger(cuda):
    from groopm_b200 import engine, synth

    rng = np.random.default_rng(41)
    nodes = synth.som_nodes(28, 4, 5)
    block, _ = synth.cfg_bmu(5000, 8, 6)
    for kw in (dict(iterations=25, vector_subset=600, influence_rate=0.15),
               dict(iterations=12, vector_subset=10000, influence_rate=0.4, radius=9.0)):
        np.random.seed(3)
        want = oracle.som_train(nodes, block, **kw)
        np.random.seed(3)
        got = engine.som_train(nodes, block, **kw)
        assert np.array_equal(got, want)
    # caller-owned weights without a mask: the BMU search keeps the weights it was handed
    w0 = rng.random((9, 13, 4))
    np.random.seed(4)
    want = oracle.som_train(None, block[:700], weights=w0, iterations=8, radius=5.0, influence_rate=0.2)
    np.random.seed(4)
    got = engine.som_train(None, block[:700], weights=w0, iterations=8, radius=5.0, influence_rate=0.2)
    assert np.array_equal(got, want)


# ---------------------------------------------------------------- f4: cDist family (refine.py:1505-1583)

def _cdist_sets(g):
    return [g[f"rows{i}"] for i in range(int(g["n_sets"]))], [g[f"pick{i}"] for i in range(int(g["n_sets"]))]


def _same_or_both_nan(a, b, atol=0.0):
    return (np.isnan(a) and np.isnan(b)) or abs(a - b) <= atol


def test_oracle_cdist_equals_reference(golden):
    """getEvenlySpacedPtsZ, cDist, cDistBetweenBins and cDistMergedBins of the reference, run unmodified
    on bins below and above the 100-row sampling limit, with a zero-coverage row (its pairs raise
    FloatingPointError in the reference and are dropped) and a bin whose every pair is dropped."""
    g = golden("ref_cdist.npz")
    cov, tcp, norms = g["cov"], g["tcp"], g["norms"]
    rows, picks = _cdist_sets(g)
    for r, pick, want in zip(rows, picks, g["c_dist"]):
        if len(r) > 100:
            assert np.array_equal(oracle.evenly_spaced_pts_z(tcp, r, 100), pick)
        assert _same_or_both_nan(oracle.c_dist_rows(cov, tcp, r, norms), float(want))
    for (a, b), want in zip(g["pairs"], g["c_dist_between"]):
        assert _same_or_both_nan(oracle.c_dist_between_rows(cov, tcp, rows[a], rows[b], norms), float(want))
    for (a, b), want in zip(g["pairs"][:3], g["c_dist_merged"]):
        assert _same_or_both_nan(oracle.c_dist_merged_rows(cov, tcp, rows[a], rows[b], norms), float(want))


def test_host_row_selection_equals_reference(golden):
    from groopm_b200 import dropin
    g = golden("ref_cdist.npz")
    rows, picks = _cdist_sets(g)
    for r, pick in zip(rows, picks):
        if len(r) > 100:
            assert np.array_equal(dropin.evenly_spaced_pts_z(g["tcp"], r, 100), pick)
    rng = np.random.default_rng(5)
    z = rng.random((5000, 3))
    for m, size in ((101, 100), (333, 100), (4999, 100), (250, 7), (1000, 999)):
        r = np.sort(rng.choice(5000, m, replace=False))
        assert np.array_equal(dropin.evenly_spaced_pts_z(z, r, size), oracle.evenly_spaced_pts_z(z, r, size))


# The one implementation-defined corner of refine.py:1545-1546: a pair whose cosine argument
# dot/(n1*n2) lands within a few ulp of +-1 (identical or collinear profiles).  Whether it rounds to
# <= 1 (angle ~0, kept) or just above (arccos raises, pair dropped) is decided by the summation order
# of BLAS' ddot on the host the reference runs on.  Such pairs form an explicit TIE CLASS: the engine
# may keep them (angle <= TIE_ANGLE) or drop them; every other pair must agree to ANGLE_ATOL.
TIE_ULPS = 8
TIE_ANGLE = 1e-6        # arccos(1 - 8 ulp) ~ 6e-8 rad
ANGLE_ATOL = 1e-12      # radians: the remaining difference is arccos' own rounding


def _tie_class(arg):
    return np.abs(np.abs(arg) - 1.0) <= TIE_ULPS * np.finfo(np.float64).eps


def _check_block(got, want_ang, want_arg, upper_only=False):
    """got: engine angles; want_*: oracle.angle_pairs.  Returns the tie-class mask."""
    tie = _tie_class(want_arg)
    if upper_only:
        keep = np.triu(np.ones_like(tie, dtype=bool), 1)
        got, want_ang, tie = got[keep], want_ang[keep], tie[keep]
    plain = ~tie
    assert np.array_equal(np.isnan(got[plain]), np.isnan(want_ang[plain]))
    ok = plain & ~np.isnan(want_ang)
    assert np.all(np.abs(got[ok] - want_ang[ok]) <= ANGLE_ATOL)
    g = got[tie]
    # a cosine argument near +1 gives an angle near 0, near -1 an angle near pi
    assert np.all(np.isnan(g) | (g <= TIE_ANGLE) | (g >= np.pi - TIE_ANGLE))
    return tie, got, want_ang


def _median_bounds(want_ang, tie):
    """The medians the reference's arithmetic can produce when the tie class is dropped or kept."""
    dropped = want_ang[~tie & ~np.isnan(want_ang)]
    kept = np.concatenate([dropped, np.where(np.isnan(want_ang[tie]), 0.0, want_ang[tie])])
    med = lambda v: float(np.median(v)) if len(v) else float("nan")   # noqa: E731
    return med(kept), med(dropped)


@pytest.mark.gpu
def test_engine_cdist_equals_reference(cuda, golden):
    """cDist / cDistBetweenBins / cDistMergedBins of the reference (run unmodified, refine.py:1528-1583)
    against the drop-ins, with the reference's own ``normCoverages`` as the divisor.  Pair by pair:
    identical keep/drop decisions and angles outside the tie class; the medians equal the reference's
    whenever a set holds no tie-class pair, and lie between the two admissible medians otherwise."""
    from groopm_b200 import dropin, engine
    g = golden("ref_cdist.npz")
    cov, tcp, norms = g["cov"], g["tcp"], g["norms"]
    rows, picks = _cdist_sets(g)
    n_exact = 0

    def check(got_med, want_med, p1, p2):
        nonlocal n_exact
        same = p2 is None
        a, b = cov[p1], cov[p1 if same else p2]
        want_ang, want_arg = oracle.angle_pairs(a, b, norms[p1], norms[p1 if same else p2])
        got = engine.angle_block(a, None if same else b, norms[p1], None if same else norms[p2])
        tie, got_v, want_v = _check_block(got, want_ang, want_arg, upper_only=same)
        if not tie.any():
            assert _same_or_both_nan(got_med, float(want_med), ANGLE_ATOL)
            n_exact += 1
        else:
            lo, hi = _median_bounds(want_v, tie)
            assert (np.isnan(got_med) and np.isnan(hi)) or lo - ANGLE_ATOL <= got_med <= hi + ANGLE_ATOL

    for r, pick, want in zip(rows, picks, g["c_dist"]):
        check(dropin.c_dist_rows(cov, tcp, r, norms), want, pick, None)
    for (a, b), want in zip(g["pairs"], g["c_dist_between"]):
        check(dropin.c_dist_between_rows(cov, tcp, rows[a], rows[b], norms), want, picks[a], picks[b])
    for (a, b), want in zip(g["pairs"][:3], g["c_dist_merged"]):
        merged = np.concatenate((rows[a], rows[b]))
        pick = oracle.evenly_spaced_pts_z(tcp, merged, 100) if len(merged) > 100 else merged
        check(dropin.c_dist_merged_rows(cov, tcp, rows[a], rows[b], norms), want, pick, None)
    assert n_exact >= 8          # most sets hold no tie-class pair and must match the reference outright
