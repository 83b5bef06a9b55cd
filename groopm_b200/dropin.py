"""Drop-in mirrors of the reference's call sites on the distance path.

Same names, argument meaning, return layouts and error behaviour as the scipy /
numpy / GroopM calls they replace; every distance,
argmin, top-k and radius selection runs in ``libgmdist.so``.  Host-side glue that
is not distance work (whitening, the inverse-distance weighted moves) stays
numpy, as in the reference.
"""
from __future__ import annotations

import numpy as np

from . import engine
from .engine import cdist, pdist, squareform  # noqa: F401  (scipy.spatial.distance names)


# ------------------------------------------------------------------ torusMesh.py / som.py

def best_match(flat_nodes, target_vector, rows: int, columns: int):
    """TorusMesh.bestMatch (torusMesh.py:123-130) -> [row, col] Python ints,
    including the reference's ``col = loc - row*rows`` arithmetic."""
    loc = int(engine.bmu(flat_nodes, np.asarray(target_vector, dtype=np.float64)[None, :])[0])
    row = int(loc / columns)
    col = loc - (row * rows)
    return [row, col]


class BestMatcher:
    """TorusMesh.bestMatch (torusMesh.py:123-130) for a mesh whose ``flatNodes`` stay put between calls -- what
    ``fixFlatNodes`` (torusMesh.py:103) leaves behind: the nodes are uploaded ONCE, and every call costs a
    32-byte copy, one kernel launch (``gmd_bmu_small_f64``) and a 4-byte read.  ``matcher(v) -> [row, col]``
    with the reference's ``col = loc - row*rows`` arithmetic; ``matcher.many(vectors)`` for a batch."""

    def __init__(self, flat_nodes, rows: int, columns: int, device=None):
        import torch

        fn = np.ascontiguousarray(flat_nodes, dtype=np.float64)
        if fn.ndim != 2:
            raise ValueError("XA must be a 2-dimensional array.")
        self.rows, self.columns = int(rows), int(columns)
        self._dev = engine._device(device)
        self._nodes = torch.from_numpy(fn).to(self._dev)
        self._v_host = torch.empty((1, fn.shape[1]), dtype=torch.float64, pin_memory=True)
        self._v = torch.empty((1, fn.shape[1]), dtype=torch.float64, device=self._dev)
        self._out = torch.empty(2, dtype=torch.int32, device=self._dev)      # [loc, non-finite flag]
        self._out_host = torch.empty(2, dtype=torch.int32, pin_memory=True)
        self._lib = engine._lib.lib()

    def __call__(self, target_vector):
        import torch

        v = np.asarray(target_vector, dtype=np.float64).reshape(1, -1)
        if v.shape[1] != self._nodes.shape[1]:
            raise ValueError("XA and XB must have the same number of columns "
                             "(i.e. feature dimension.)")
        self._v_host.numpy()[:] = v
        with torch.cuda.device(self._dev):
            st = engine._stream_ptr()
            self._v.copy_(self._v_host, non_blocking=True)
            engine._lib.check(self._lib.gmd_bmu_small_f64(self._nodes.data_ptr(), self._nodes.shape[0], self._v.data_ptr(),
                                                          1, v.shape[1], self._out.data_ptr(), 0,
                                                          self._out.data_ptr() + 4, st))
            self._out_host.copy_(self._out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        loc, bad = int(self._out_host[0]), int(self._out_host[1])
        if bad:
            raise FloatingPointError("non-finite value in an input profile")
        row = int(loc / self.columns)
        return [row, loc - (row * self.rows)]

    def many(self, vectors):
        return best_match_many(self._nodes, vectors, self.rows, self.columns)


def best_match_many(flat_nodes, vectors, rows: int, columns: int):
    """bestMatch for every row of ``vectors`` in one launch -> (row[], col[]) int64."""
    loc = engine.bmu(flat_nodes, vectors)
    loc = (loc.cpu().numpy() if hasattr(loc, "cpu") else np.asarray(loc)).astype(np.int64)
    row = (loc / columns).astype(np.int64)
    col = loc - row * rows
    return row, col


def classify_contigs(flat_nodes, block, bin_assignments, rows: int, columns: int):
    """The per-contig ``SS.classifyContig(block[i])`` loops (refine.py:1219-1222,
    2137-2138; som.py:572-575) as one call -> int64 bin id per contig."""
    row, col = best_match_many(flat_nodes, block, rows, columns)
    return np.asarray(bin_assignments)[row, col].astype(np.int64)


def som_train_bmus(flat_nodes, train_vectors, index_array):
    """The inner ``argmin(cdist(flat_nodes, [trainVector[j]]))`` of SOM.train
    (som.py:348-353) for all j of one iteration (nodes are frozen inside it)
    -> flat locations ``loc`` [len(index_array)] int64; the caller's ``row = int(loc/cols)``,
    ``col = loc - row*cols`` (som.py:352-353) follow from them."""
    tv = np.asarray(train_vectors, dtype=np.float64)[np.asarray(index_array)]
    return np.asarray(engine.bmu(flat_nodes, tv)).astype(np.int64)


def som_train(nodes, trainVector, weights=None, iterations=1000, vectorSubSet=1000, weightImgFileNamePrefix="",
              epsilom=0.0001, influenceRate=0.4, mask=None, radius=0., silent=True):
    """SOM.train (som.py:214-392) with the reference's argument names; ``nodes`` stands for
    ``self.weights.nodes`` and ``mask`` may be the reference's dict of (row, col) points or a
    [rows, cols] array.  Image dumps (``weightImgFileNamePrefix``) and progress printing stay
    with the reference; ``epsilom`` is unused there too.  Returns the trained weights; a caller
    that passed ``weights=None`` stores them with ``self.weights.fixFlatNodes(weights=...)``."""
    base = nodes if weights is None else weights
    if isinstance(mask, dict):
        rows, cols = np.shape(base)[0], np.shape(base)[1]
        m = np.zeros((rows, cols), dtype=np.uint8)
        for (r, c) in mask.keys():
            m[r, c] = 1
        mask = m
    return engine.som_train(nodes, trainVector, weights=weights, iterations=iterations, vector_subset=vectorSubSet,
                            influence_rate=influenceRate, mask=mask, radius=radius)


# ------------------------------------------------------------------ refine.py

class KDTree:
    """``scipy.spatial.KDTree(data).query(x, k)`` as used by findMergeGroups
    (refine.py:653-665): Euclidean, ascending distance, returns (d, i)."""

    def __init__(self, data):
        self.data = np.ascontiguousarray(data, dtype=np.float64)
        if self.data.ndim != 2:
            raise ValueError("data must be 2 dimensions")
        self.n, self.m = self.data.shape
        self._dev_data = None                               # uploaded on the first query, then resident

    def query(self, x, k=1, p=2):
        if p != 2:
            raise ValueError("only the Euclidean norm (p=2) is used by the reference")
        xa = np.asarray(x, dtype=np.float64)
        single = xa.ndim == 1
        xq = xa[None, :] if single else xa
        if xq.shape[-1] != self.m:
            raise ValueError(f"x must consist of vectors of length {self.m} but has shape {xa.shape}")
        kk = min(int(k), self.n)
        import torch

        if self._dev_data is None:
            self._dev_data = torch.from_numpy(self.data).to(engine._device())
        if single and self.n <= engine._lib.lib().gmd_knn_small_max_candidates():
            # one vector against a small tree (refine.py:664-665 calls this once per bin): staging buffers are kept,
            # the call is a 24-byte copy, one kernel launch (gmd_knn_small_f64) and one read-back
            return self._query_one(xq, int(k), kk)
        # the tree's points stay on the device; only the query rows travel (refine.py:664-665 calls this once
        # per bin in a Python loop)
        xd = torch.from_numpy(np.ascontiguousarray(xq)).to(self._dev_data.device)
        idx_d, dist_d = engine.knn(xd, self._dev_data, k=kk, metric="euclidean")
        idx, dist = idx_d.cpu().numpy().astype(np.intp), dist_d.cpu().numpy()
        if kk < k:  # scipy pads missing neighbours with (inf, n)
            pad = k - kk
            dist = np.concatenate([dist, np.full((len(xq), pad), np.inf)], axis=1)
            idx = np.concatenate([idx, np.full((len(xq), pad), self.n, dtype=np.intp)], axis=1)
        if k == 1:
            dist, idx = dist[:, 0], idx[:, 0]
        if single:
            return dist[0], idx[0]
        return dist, idx


def _kdtree_query_one(self, xq, k, kk):
    import torch

    dev = self._dev_data.device
    st = getattr(self, "_one", None)
    if st is None or st[0] != kk:
        st = self._one = (kk, torch.empty((1, self.m), dtype=torch.float64, pin_memory=True),
                          torch.empty((1, self.m), dtype=torch.float64, device=dev),
                          torch.empty(3 * kk + 2, dtype=torch.float64, device=dev),      # dist [kk] | idx, flag (as int32)
                          torch.empty(3 * kk + 2, dtype=torch.float64, pin_memory=True))
    _, v_host, v_dev, out_dev, out_host = st
    v_host.numpy()[:] = xq
    with torch.cuda.device(dev):
        v_dev.copy_(v_host, non_blocking=True)
        base = out_dev.data_ptr()
        engine._lib.check(engine._lib.lib().gmd_knn_small_f64(
            v_dev.data_ptr(), 1, self._dev_data.data_ptr(), self.n, self.m, kk, engine._lib.L2,
            base + 8 * kk, base, base + 8 * kk + 4 * kk, engine._stream_ptr()))
        out_host.copy_(out_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    raw = out_host.numpy()
    dist = raw[:kk].copy()
    ints = raw[kk:].view(np.int32)
    idx = ints[:kk].astype(np.intp)
    if int(ints[kk]):
        raise FloatingPointError("non-finite value in an input profile")
    if kk < k:  # scipy pads missing neighbours with (inf, n)
        dist = np.concatenate([dist, np.full(k - kk, np.inf)])
        idx = np.concatenate([idx, np.full(k - kk, self.n, dtype=np.intp)])
    if k == 1:
        return dist[0], idx[0]
    return dist, idx


KDTree._query_one = _kdtree_query_one


def k_dist(bin_k_vals):
    """RefineEngine.kDist (refine.py:1490-1496): median cityblock pdist of a bin, or None."""
    return engine.median_dist(bin_k_vals, None, "cityblock")


def k_dist_between(k1, k2):
    """RefineEngine.kDistBetweenBins (refine.py:1502-1503)."""
    return engine.median_dist(k1, k2, "cityblock")


def k_dist_many(kmer_pcs, bins_row_indices):
    """kDist for every bin in one launch -- the loop of RefineEngine.getKCut
    (refine.py:1479-1488) over ``self.BM.getBin(bid).rowIndices``: list of medians (None for a
    bin of fewer than two contigs)."""
    return engine.median_dist_batch(kmer_pcs, [(r, None) for r in bins_row_indices], "cityblock")


def k_dist_between_many(kmer_pcs, bin_pairs):
    """kDistBetweenBins for many (rowIndices1, rowIndices2) pairs in one launch -- the
    neighbour tests of findMergeGroups / combineMergers (refine.py:718, 876)."""
    return engine.median_dist_batch(kmer_pcs, list(bin_pairs), "cityblock")


def c_dist(cov_rows, norms=None):
    """RefineEngine.cDist (refine.py:1528-1551) on the already selected rows (<= 100 per
    bin, refine.py:1529-1536): median angle over the pairs i < j, pairs that raise a
    FloatingPointError in the reference left out.  ``norms`` = ``self.PM.normCoverages`` of
    those rows (refine.py:1546 divides by the stored norms); None = norms of the rows.
    nan when no pair is left (np.median([]))."""
    ang = engine.angle_block(cov_rows, None, norms)
    iu = np.triu_indices(ang.shape[0], 1)
    v = ang[iu]
    v = v[~np.isnan(v)]
    return float(np.median(v)) if len(v) else float("nan")


def c_dist_between(cov1, cov2, norms1=None, norms2=None):
    """RefineEngine.cDistBetweenBins (refine.py:1557-1583) on the selected rows."""
    v = engine.angle_block(cov1, cov2, norms1, norms2).ravel()
    v = v[~np.isnan(v)]
    return float(np.median(v)) if len(v) else float("nan")


def evenly_spaced_pts_z(transformed_cp, row_indices, sample_size: int):
    """RefineEngine.getEvenlySpacedPtsZ (refine.py:1505-1516): ``sample_size`` of the rows, spread
    evenly along the Z axis (last column) of the transformed coverage space.  Host logic; the
    positions are the truncations of the reference's RUNNING float sum (0, s, s+s, ...), which
    is not always ``int(i * s)``."""
    row_indices = np.asarray(row_indices)
    by_z = np.argsort(np.asarray(transformed_cp)[row_indices, -1])
    step = float(len(row_indices)) / sample_size
    pos, acc = [], 0.0
    for _ in range(sample_size):
        pos.append(int(acc))
        acc += step
    return row_indices[by_z[np.array(pos, dtype=np.int64)]]


def _c_pick(transformed_cp, row_indices, max_in_bin: int = 100):
    row_indices = np.asarray(row_indices)
    if len(row_indices) > max_in_bin:
        return evenly_spaced_pts_z(transformed_cp, row_indices, max_in_bin)
    return row_indices


def _norms_of(norm_coverages, pick):
    return None if norm_coverages is None else np.asarray(norm_coverages, dtype=np.float64)[pick]


def c_dist_rows(cov_profiles, transformed_cp, row_indices, norm_coverages=None):
    """RefineEngine.cDist (refine.py:1528-1551) from a bin's row indices: at most 100 rows, picked
    evenly along Z when the bin is larger, then the median angle over the pairs.
    ``norm_coverages`` = ``self.PM.normCoverages`` (all contigs)."""
    pick = _c_pick(transformed_cp, row_indices)
    return c_dist(np.asarray(cov_profiles)[pick], _norms_of(norm_coverages, pick))


def c_dist_between_rows(cov_profiles, transformed_cp, rows1, rows2, norm_coverages=None):
    """RefineEngine.cDistBetweenBins (refine.py:1557-1583) from two bins' row indices."""
    cov = np.asarray(cov_profiles)
    p1, p2 = _c_pick(transformed_cp, rows1), _c_pick(transformed_cp, rows2)
    return c_dist_between(cov[p1], cov[p2], _norms_of(norm_coverages, p1), _norms_of(norm_coverages, p2))


def c_dist_merged_rows(cov_profiles, transformed_cp, rows1, rows2, norm_coverages=None):
    """RefineEngine.cDistMergedBins (refine.py:1553-1555)."""
    return c_dist_rows(cov_profiles, transformed_cp, np.concatenate((np.asarray(rows1), np.asarray(rows2))),
                       norm_coverages)


# ------------------------------------------------------------------ binManager.py

def vq(obs, code_book):
    """``scipy.cluster.vq.vq(obs, code_book)`` as used by BinManager.getSplitties
    (binManager.py:560-566): nearest centroid of every observation (first minimum) and the
    Euclidean distance to it -> (code, dist).  1-D inputs are single-feature observations."""
    o = np.asarray(obs, dtype=np.float64)
    cb = np.asarray(code_book, dtype=np.float64)
    if o.ndim != cb.ndim:
        raise ValueError("Observation and code_book should have the same rank")
    if o.ndim == 1:
        o, cb = o[:, None], cb[:, None]
    elif o.ndim != 2:
        raise ValueError("ndim different than 1 or 2 are not supported")
    if o.shape[1] != cb.shape[1]:
        raise ValueError("Observation and code_book should have the same number of features")
    code, dist = engine.bmu(cb, o, return_dist=True)
    return np.asarray(code).astype(np.int32), np.asarray(dist)


def kmeans_step(obs, code_book):
    """One iteration of scipy's k-means loop (``kmeans`` at binManager.py:560): assign with
    ``vq``, then move every centroid to the mean of its members; centroids without members
    are dropped.  -> (new_code_book, codes, mean distortion)."""
    o = np.asarray(obs, dtype=np.float64)
    o2 = o[:, None] if o.ndim == 1 else o
    cb = np.asarray(code_book, dtype=np.float64)
    code, dist = vq(o, cb)
    nc = len(cb)
    sums = np.zeros((nc, o2.shape[1]))
    np.add.at(sums, code, o2)              # observation order, like _vq.update_cluster_means
    counts = np.bincount(code, minlength=nc)
    has = counts > 0
    new = sums[has] / counts[has][:, None]
    return (new[:, 0] if o.ndim == 1 else new), code, float(dist.mean())


# ------------------------------------------------------------------ cluster.py

def whiten(dat):
    """cluster.py:488-491 / refine.py:647-651 (host numpy, as in the reference)."""
    dat = np.asarray(dat, dtype=np.float64)
    mean = np.mean(dat, axis=0)
    std = np.std(dat, axis=0)
    std += np.where(std == 0, 1, 0)
    return (dat - mean) / std


def kth_neighbour_radius(x, eps_neighbours: int, metric="cityblock"):
    """``np.median(np.sort(squareform(pdist(X, metric)))[:, eps])`` (cluster.py:495-498)
    without the n x n matrix -> (radius, per-row k-th distances)."""
    kth = engine.kth_neighbour_dist(x, eps_neighbours, metric)
    return float(np.median(kth)), kth


def two_way_contraction_step(c_dat, k_dat, move: float = 0.1):
    """The distance-driven part of one twoWayContraction iteration (cluster.py:430,
    488-498, 530-590): both radii, the radius-neighbour lists with the reference's
    drop-the-first-neighbour quirk, cross-metric inverse-distance weights and the
    contraction moves.  The n x n matrices are never built."""
    c_dat = np.ascontiguousarray(c_dat, dtype=np.float64)
    k_dat = np.ascontiguousarray(k_dat, dtype=np.float64)
    n = len(c_dat)
    eps = int(np.min([10, int(n / 2)]))
    c_w = whiten(c_dat)
    c_radius, _ = kth_neighbour_radius(c_w, eps, "cityblock")
    k_radius, _ = kth_neighbour_radius(k_dat, eps, "cityblock")
    min_nb = np.max([1, 0.1 * eps])

    def contract(space_x, radius, other_x, other_radius, moved):
        indptr, indices = engine.radius_neighbours(space_x, radius, "cityblock")
        counts = np.diff(indptr)
        keep_row = counts > min_nb
        # drop the first (lowest-index) neighbour of every kept row: cluster.py:535-536
        drop = np.zeros(len(indices), dtype=bool)
        drop[indptr[:-1][counts > 0]] = True
        row_of = np.repeat(np.arange(n), counts)
        sel = (~drop) & keep_row[row_of]
        other = engine.gather_dist(other_x, indptr, indices, "cityblock")
        inv = 1.0 / (other[sel] + 0.1 * other_radius)
        rows = row_of[sel]
        nb = indices[sel].astype(np.int64)
        sums = np.zeros(n)
        np.add.at(sums, rows, inv)
        outs = []
        for dat in moved:
            acc = np.zeros_like(dat)
            np.add.at(acc, rows, dat[nb] * (inv / sums[rows])[:, None])
            out = np.where(keep_row[:, None], (1 - move) * dat + move * acc, dat)
            outs.append(out)
        noise = set(np.flatnonzero(~keep_row).tolist())
        return outs, noise

    (new_k,), k_noise = contract(c_w, c_radius, k_dat, k_radius, (k_dat,))
    (new_c,), c_noise = contract(k_dat, k_radius, c_w, c_radius, (c_dat,))
    return {"k_dat": new_k, "c_dat": new_c, "c_radius": c_radius, "k_radius": k_radius,
            "k_noise": k_noise, "c_noise": c_noise, "eps": eps}
