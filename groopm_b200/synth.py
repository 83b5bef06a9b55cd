"""Seeded synthetic contig profiles shaped like GroopM's ProfileManager arrays.

These are the *inputs* to the distance / neighbour-search path, not part of it.
The reference builds them on the host from BAM/FASTA files
(profileManager.py:124-139, 201-226); here the same shapes are drawn from a
seeded generator so that the oracle and the CUDA engine
see identical float64 arrays:

    covProfiles    N x S   lognormal per "genome" cluster
    transformedCP  N x 3   restated coverage transform (mstore.py:1810-1854)
    kmerSigs       N x 136 Dirichlet rows, each sums to 1 (mstore.py:1701-1711)
    kmerPCs        N x P   centre/scale + SVD, PCs to 80 % variance
                           (mstore.py:1605-1618, PCA.py:64-95)
    kmerNormPC1    N       PC1 min-max scaled (profileManager.py:221-223)

Everything is float64, C-contiguous.
"""
from __future__ import annotations

import numpy as np

KMER_DIM = 136  # canonical tetranucleotides (mstore.py:1640-1666)


def quantise(x: np.ndarray, bits: int) -> np.ndarray:
    """Snap to a 2^-bits grid.  The generators go through exp/log/SVD whose last bits vary
    between CPUs and BLAS builds; on a grid the arrays reproduce bit for bit (the golden
    fixtures record a SHA-256 of every input they were made from)."""
    s = float(2 ** bits)
    return np.round(x * s) / s


def _labels(n: int, rng: np.random.Generator) -> tuple[np.ndarray, int]:
    g = max(8, n // 2000)
    return rng.integers(0, g, size=n), g


def add_duplicates(x: np.ndarray, rng: np.random.Generator, frac: float = 0.001) -> np.ndarray:
    """Overwrite ``frac`` of the rows with exact copies of other rows (tie cases)."""
    n = x.shape[0]
    m = int(n * frac)
    if m > 0:
        dst = rng.choice(n, size=m, replace=False)
        src = rng.integers(0, n, size=m)
        x[dst] = x[src]
    return x


def coverage_profiles(n: int, s: int, seed: int) -> np.ndarray:
    """covProfiles N x S: lognormal(mu_g[s], 0.25), mu_g ~ U(0, 5)."""
    rng = np.random.default_rng(seed)
    lab, g = _labels(n, rng)
    mu = rng.uniform(0.0, 5.0, size=(g, s))
    cov = np.exp(mu[lab] + 0.25 * rng.standard_normal((n, s)))
    return np.ascontiguousarray(cov)


def transform_cp(cov: np.ndarray, scale_factor: float = 1000.0) -> np.ndarray:
    """Vectorised restatement of CoverageTransformer.transformCP (mstore.py:1810-1854).

    Unit-circle barycentric XY, log10 shrink towards the centre, Z = log10 of
    the coverage norm, then each axis rescaled to [0, scale_factor - 1].
    """
    n, s = cov.shape
    ang = np.arange(s) * 2.0 * np.pi / s
    flat = cov / cov.sum(axis=1, keepdims=True)
    xy = np.stack([flat @ np.cos(ang), flat @ np.sin(ang)], axis=1)
    sv = np.linalg.norm(xy * scale_factor, axis=1)
    shrink = np.where(sv > 1.0, np.log10(np.maximum(sv, 1.0)), 1.0)
    shrink = np.where(shrink > 0.0, shrink, 1.0)
    xy = xy / shrink[:, None]
    z = np.log10(np.linalg.norm(cov, axis=1))
    tcp = np.concatenate([xy, z[:, None]], axis=1)
    tcp -= tcp.min(axis=0)
    mx = tcp.max(axis=0) / (scale_factor - 1.0)
    tcp /= mx
    return np.ascontiguousarray(tcp)


def whitened_log_coverage(cov: np.ndarray) -> np.ndarray:
    """log-coverage rows whitened per column the way cluster.py:488-491 whitens."""
    x = np.log(cov)
    mean = x.mean(axis=0)
    std = x.std(axis=0)
    std = std + np.where(std == 0, 1, 0)
    return np.ascontiguousarray((x - mean) / std)


def kmer_sigs(n: int, seed: int) -> np.ndarray:
    """kmerSigs N x 136: Dirichlet(alpha_g * 50) per cluster; rows sum to 1."""
    rng = np.random.default_rng(seed)
    lab, g = _labels(n, rng)
    alpha = rng.dirichlet(np.ones(KMER_DIM), size=g) * 50.0 * KMER_DIM / 8.0 + 1e-3
    gam = rng.standard_gamma(alpha[lab])
    gam /= gam.sum(axis=1, keepdims=True)
    return np.ascontiguousarray(gam)


def kmer_pcs(sigs: np.ndarray, fraction: float = 0.80) -> np.ndarray:
    """Centre/scale then SVD; keep PCs up to ``fraction`` variance, at least 2.

    Restates KmerSigEngine.PCAKSigs (mstore.py:1605-1618) and PCA.__init__/pc
    (PCA.py:64-95) including the sign fix on Vt[0, 0].
    """
    a = np.array(sigs, dtype=np.float64, copy=True)
    a -= a.mean(axis=0)
    std = a.std(axis=0)
    a /= np.where(std, std, 1.0)
    u, d, vt = np.linalg.svd(a, full_matrices=False)
    if vt[0, 0] < 0:
        u *= -1.0
    sumvar = np.cumsum(d ** 2)
    sumvar /= sumvar[-1]
    npc = int(np.searchsorted(sumvar, fraction)) + 1
    while npc == 1:
        fraction *= 1.1
        npc = int(np.searchsorted(sumvar, fraction)) + 1
    return np.ascontiguousarray(u[:, :npc] * d[:npc])


def kmer_norm_pc1(kpcs: np.ndarray) -> np.ndarray:
    """profileManager.py:221-223."""
    v = np.array(kpcs[:, 0], copy=True)
    v -= v.min()
    v /= v.max()
    return v


def som_block(tcp: np.ndarray, kpc1: np.ndarray) -> np.ndarray:
    """block N x 4 = [transformedCP | kmerNormPC1], global min-max normalised
    (refine.py:1048-1058, 1211-1217)."""
    block = np.zeros((tcp.shape[0], 4))
    block[:, :-1] = tcp
    block[:, -1] = kpc1
    minz = block.min(axis=0)
    maxz = block.max(axis=0) - minz
    block -= minz
    block /= maxz
    return block


def som_nodes(side: int, dim: int, seed: int) -> np.ndarray:
    """Random torus mesh weights U(0,1), rows x cols x dim (torusMesh.py:97-98)."""
    rng = np.random.default_rng(seed)
    return rng.random(side * side * dim).reshape((side, side, dim))


# ---------------------------------------------------------------------------
# the five BASELINE.json configs


def cfg_knn_rows(n: int, d: int, seed: int, dups: bool = True) -> np.ndarray:
    """Whitened log-coverage rows N x D (cfg2: D=10 seed 2, cfg5: D=30 seed 5)."""
    x = quantise(whitened_log_coverage(coverage_profiles(n, d, seed)), 20)
    if dups:
        add_duplicates(x, np.random.default_rng(seed + 1000))
    return x


def cfg_bmu(n: int, side: int, seed: int = 3) -> tuple[np.ndarray, np.ndarray]:
    """cfg3: block N x 4 and side x side x 4 torus nodes."""
    cov = coverage_profiles(n, 3, seed)
    tcp = transform_cp(cov)
    rng = np.random.default_rng(seed + 1)
    lab, g = _labels(n, np.random.default_rng(seed))
    centre = rng.uniform(0.0, 1.0, size=g)
    kpc1 = np.clip(centre[lab] + 0.03 * rng.standard_normal(n), 0.0, None)
    kpc1 -= kpc1.min()
    kpc1 /= kpc1.max()
    block = quantise(som_block(tcp, kpc1), 24)
    add_duplicates(block, np.random.default_rng(seed + 1000))
    return block, quantise(som_nodes(side, 4, seed + 2), 24)


def cfg_core(n: int = 10000, s: int = 3, seed: int = 1) -> dict:
    """cfg1: transformedCP (D=3) and kmerPCs (D=P) for the core-clustering calls."""
    cov = coverage_profiles(n, s, seed)
    tcp = quantise(transform_cp(cov), 16)
    sigs = quantise(kmer_sigs(n, seed), 34)
    kpcs = quantise(kmer_pcs(sigs), 20)
    add_duplicates(tcp, np.random.default_rng(seed + 1000))
    return {"cov": cov, "transformedCP": tcp, "kmerSigs": sigs, "kmerPCs": kpcs,
            "kmerNormPC1": kmer_norm_pc1(kpcs)}
