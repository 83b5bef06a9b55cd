"""SURVEY.md section 8 "next" rows f3 (nearest-centroid assign / k-means step,
binManager.py:551-566) and f4 (angle between coverage profiles, refine.py:1528-1583)."""
import numpy as np
import pytest

import oracle
from groopm_b200 import synth

ANGLE_ATOL = 1e-12   # radians; np.dot goes through BLAS, whose summation order is not the kernel's


def _cov(n, s, seed):
    return synth.coverage_profiles(n, s, seed)


def test_oracle_angle_cases():
    a = np.array([[1.0, 0.0, 0.0], [0.0, 2.0, 0.0], [3.0, 3.0, 0.0], [0.0, 0.0, 0.0]])
    # orthogonal pair = pi/2, 45 degree pairs, zero-norm row dropped (FloatingPointError in the reference)
    assert oracle.c_dist(a[:2]) == pytest.approx(np.pi / 2, abs=1e-15)
    assert oracle.c_dist_between(a[:1], a[2:3]) == pytest.approx(np.pi / 4, abs=1e-15)
    assert np.isnan(oracle.c_dist(a[[0, 3]]))
    assert oracle.c_dist(a) == pytest.approx(np.median([np.pi / 2, np.pi / 4, np.pi / 4]), abs=1e-15)


def test_oracle_vq_first_minimum():
    code, dist = oracle.vq(np.array([0.0, 1.0, 0.5]), np.array([0.0, 1.0, 0.0]))
    assert code.tolist() == [0, 1, 0] and np.allclose(dist, [0.0, 0.0, 0.5])


@pytest.mark.gpu
def test_engine_angles(cuda):
    from groopm_b200 import dropin, engine

    for s, seed in ((3, 1), (10, 2), (30, 3)):
        c = _cov(260, s, seed)
        a, b = c[:100], c[100:187]
        assert dropin.c_dist(a) == pytest.approx(oracle.c_dist(a), abs=ANGLE_ATOL)
        assert dropin.c_dist_between(a, b) == pytest.approx(oracle.c_dist_between(a, b), abs=ANGLE_ATOL)
    z = np.array([[1.0, 0.0], [0.0, 0.0], [2.0, 2.0], [1.0, 0.0]])
    blk = engine.angle_block(z)
    assert np.isnan(blk[1]).all() and np.isnan(blk[:, 1]).all()          # zero norm: dropped
    assert blk[0, 2] == pytest.approx(np.pi / 4, abs=1e-15) and blk[0, 3] == 0.0
    assert np.isnan(dropin.c_dist(z[:2])) and dropin.c_dist(z) == pytest.approx(oracle.c_dist(z), abs=ANGLE_ATOL)


@pytest.mark.gpu
def test_engine_vq_and_kmeans_step(cuda):
    from scipy.cluster.vq import vq as sp_vq

    from groopm_b200 import dropin

    rng = np.random.default_rng(5)
    # the reference's three observation kinds (binManager.py:553-558): kmerNormPC1 and
    # contigLengths are single-feature, covProfiles has one feature per sample
    pc1 = rng.random(5000)
    code, dist = dropin.vq(pc1, np.array([0.2, 0.5, 0.5, 0.9]))
    wc, wd = sp_vq(pc1, np.array([0.2, 0.5, 0.5, 0.9]))
    assert np.array_equal(code, wc) and np.array_equal(dist, wd) and not np.any(code == 2)   # first minimum
    cov = _cov(4000, 3, 7)
    book = cov[rng.choice(4000, 5, replace=False)]
    code, dist = dropin.vq(cov, book)
    wc, wd = sp_vq(cov, book)
    assert np.array_equal(code, wc) and np.array_equal(dist, wd)                              # < 5 features: bit-exact
    cov10 = _cov(3000, 10, 8)
    book10 = cov10[rng.choice(3000, 7, replace=False)]
    code, dist = dropin.vq(cov10, book10)
    wc, wd = sp_vq(cov10, book10)            # >= 5 features: scipy expands the square through BLAS
    assert np.array_equal(code, wc)
    # ... and loses the small distances to cancellation (an observation that IS a centroid
    # comes back as 3.8e-6 instead of 0); the engine returns the direct-difference value
    np.testing.assert_allclose(dist, wd, rtol=1e-9, atol=1e-5)
    from scipy.spatial.distance import cdist as sp_cdist
    assert np.array_equal(dist, sp_cdist(cov10, book10).min(axis=1))
    new, codes, distortion = dropin.kmeans_step(cov, book)
    want = np.array([cov[wc_ == k].mean(axis=0) for k, wc_ in ((k, sp_vq(cov, book)[0]) for k in range(5))])
    np.testing.assert_allclose(new, want, rtol=1e-13)
    assert distortion == pytest.approx(sp_vq(cov, book)[1].mean(), rel=1e-14)
    with pytest.raises(ValueError):
        dropin.vq(cov, book[:, :2])
