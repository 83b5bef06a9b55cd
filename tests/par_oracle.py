"""The oracle's kNN on every host core (test infrastructure): the query rows are cut into slices and each
forked worker runs ``oracle.knn`` -- the literal scipy expressions -- on its slice against all candidates."""
import multiprocessing as mp
import os

import numpy as np

import oracle

_Q = _C = None


def _work(args):
    lo, hi, k, metric = args
    return oracle.knn(_Q[lo:hi], _C, k, metric, chunk=8)


def knn(q: np.ndarray, c: np.ndarray, k: int, metric: str = "euclidean", procs: int | None = None):
    """-> (dist [nq, k] float64, idx [nq, k]) exactly as ``oracle.knn``."""
    global _Q, _C
    procs = procs or len(os.sched_getaffinity(0))
    procs = max(1, min(procs, len(q) // 8 or 1))
    if procs == 1:
        return oracle.knn(q, c, k, metric, chunk=8)
    _Q, _C = q, c
    bounds = np.linspace(0, len(q), procs + 1).astype(int)
    jobs = [(int(a), int(b), k, metric) for a, b in zip(bounds, bounds[1:]) if b > a]
    with mp.get_context("fork").Pool(procs) as pool:
        parts = pool.map(_work, jobs)
    _Q = _C = None
    return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
