"""Literal float64 restatement of the reference's hot-path call sites.

Each function cites the ``/root/reference/groopm`` file:line it follows and
uses the same scipy/numpy calls on float64 inputs.  Test infrastructure only
(see ``oracle/__init__.py``).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import KDTree
from scipy.spatial.distance import cdist, pdist, squareform

__all__ = [
    "best_match", "bmu_batch", "classify_contigs", "som_train_bmus",
    "whiten", "allpairs_cityblock", "kth_neighbour_radius", "radius_neighbours",
    "contraction_weights", "two_way_contraction_step", "kdtree_knn", "knn",
    "k_dist", "k_dist_between", "small2indices", "knn_tie_aware_equal", "som_stamp", "som_train", "c_dist", "c_dist_between", "angle_pairs", "evenly_spaced_pts_z", "c_dist_rows", "c_dist_between_rows",
    "c_dist_merged_rows", "vq",
]


# ---------------------------------------------------------------- a1 / a2 / a3

def best_match(flat_nodes: np.ndarray, target: np.ndarray, rows: int, columns: int):
    """TorusMesh.bestMatch, torusMesh.py:123-130 (including ``col = loc - row*rows``)."""
    loc = int(np.argmin(cdist(flat_nodes, [target])))
    row = int(loc / columns)
    col = loc - (row * rows)
    return [row, col]


def bmu_batch(flat_nodes: np.ndarray, x: np.ndarray, chunk: int = 8192) -> np.ndarray:
    """``argmin(cdist(flat_nodes, [x_i]))`` for every row of ``x`` (torusMesh.py:127,
    som.py:351), batched: cdist(X_chunk, nodes) is the same float64 direct-difference
    arithmetic per pair, argmin keeps the first (lowest-index) minimum."""
    out = np.empty(x.shape[0], dtype=np.int64)
    for s in range(0, x.shape[0], chunk):
        out[s:s + chunk] = np.argmin(cdist(x[s:s + chunk], flat_nodes), axis=1)
    return out


def classify_contigs(flat_nodes, block, bin_assignments, rows, columns, chunk: int = 8192):
    """The per-contig loop refine.py:1219-1222 / 2137-2138 over SOM.classifyContig
    (som.py:572-575): ``int(binAssignments[r, c])`` with [r, c] from bestMatch."""
    loc = bmu_batch(flat_nodes, block, chunk)
    row = (loc / columns).astype(np.int64)
    col = loc - row * rows
    return np.asarray(bin_assignments)[row, col].astype(np.int64)


def som_train_bmus(flat_nodes: np.ndarray, train_vectors: np.ndarray, index_array) -> np.ndarray:
    """Inner BMU of SOM.train (som.py:348-353): nodes are frozen for the whole
    ``for j in index_array`` loop, so the T searches are one batch."""
    tv = np.asarray(train_vectors)[np.asarray(index_array)]
    return bmu_batch(flat_nodes, tv)


# ---------------------------------------------------------------- a4 / a5 / a6

def whiten(dat: np.ndarray) -> np.ndarray:
    """cluster.py:488-491 (same at refine.py:647-651)."""
    mean = np.mean(dat, axis=0)
    std = np.std(dat, axis=0)
    std += np.where(std == 0, 1, 0)
    return (dat - mean) / std


def allpairs_cityblock(x: np.ndarray) -> np.ndarray:
    """cluster.py:495 / :497 -- ``squareform(pdist(X, 'cityblock'))``."""
    return squareform(pdist(x, 'cityblock'))


def kth_neighbour_radius(dist_matrix: np.ndarray, eps_neighbours: int):
    """cluster.py:496 / :498 -- ``median(sort(D)[:, eps])``; column 0 is self.
    Returns (radius, per-row k-th distances)."""
    kth = np.sort(dist_matrix)[:, eps_neighbours]
    return float(np.median(kth)), kth


def radius_neighbours(dist_row: np.ndarray, radius: float, eps_neighbours: int):
    """cluster.py:534-541 / 571-578: ascending ``where(row <= r)[0]``; when there
    are more than max(1, 0.1*eps) of them the FIRST (lowest-index) one is dropped
    as the presumed self match; otherwise the row is putative noise (None)."""
    nb = np.where(dist_row <= radius)[0]
    if len(nb) > np.max([1, 0.1 * eps_neighbours]):
        return nb[1:]
    return None


def contraction_weights(other_dist_row: np.ndarray, neighbours: np.ndarray, other_radius: float):
    """cluster.py:547-557 / 583-588: inverse-distance weights from the OTHER
    metric's matrix, floor of 0.1 * other radius."""
    nd = other_dist_row[neighbours] + 0.1 * other_radius
    inv = 1.0 / nd
    return inv / np.sum(inv)


def two_way_contraction_step(c_dat: np.ndarray, k_dat: np.ndarray, move: float = 0.1):
    """One iteration of twoWayContraction's distance-driven part, cluster.py:430,
    488-498, 530-590 (max_iterations = 1, :437).  Returns the new k_dat, c_dat,
    both radii and the putative-noise index sets."""
    n = len(c_dat)
    eps = int(np.min([10, int(n / 2)]))
    c_w = whiten(c_dat)
    c_dm = allpairs_cityblock(c_w)
    c_radius, _ = kth_neighbour_radius(c_dm, eps)
    k_dm = allpairs_cityblock(k_dat)
    k_radius, _ = kth_neighbour_radius(k_dm, eps)

    new_k = np.zeros(k_dat.shape)
    k_noise = set()
    for i, row in enumerate(c_dm):
        nb = radius_neighbours(row, c_radius, eps)
        if nb is None:
            k_noise.add(i)
            new_k[i] = k_dat[i]
            continue
        w = contraction_weights(k_dm[i], nb, k_radius)
        new_k[i] = (1 - move) * k_dat[i] + move * np.sum((k_dat[nb].T * w).T, axis=0)

    # NB the reference has already replaced k_dat by new_k_dat here (cluster.py:563)
    # but k_dist_matrix is the one computed before the move.
    new_c = np.zeros(c_dat.shape)
    c_noise = set()
    for i, row in enumerate(k_dm):
        nb = radius_neighbours(row, k_radius, eps)
        if nb is None:
            c_noise.add(i)
            new_c[i] = c_dat[i]
            continue
        w = contraction_weights(c_dm[i], nb, c_radius)
        new_c[i] = (1 - move) * c_dat[i] + move * np.sum((c_dat[nb].T * w).T, axis=0)
    return {"k_dat": new_k, "c_dat": new_c, "c_radius": c_radius, "k_radius": k_radius,
            "k_noise": k_noise, "c_noise": c_noise, "eps": eps}


# ---------------------------------------------------------------- a7

def kdtree_knn(data: np.ndarray, k: int):
    """findMergeGroups, refine.py:653-665: ``KDTree(data).query(data[i], k=K)``
    for every row, Euclidean, self included, ascending distance.  K is capped at
    the number of rows (refine.py:590-594)."""
    k = min(k, len(data))
    tree = KDTree(data)
    d, i = tree.query(data, k=k)
    if k == 1:
        d = d[:, None]
        i = i[:, None]
    return d, i


def knn(q: np.ndarray, c: np.ndarray, k: int, metric: str = "euclidean", chunk: int = 1024):
    """Generalised a4+a5+a7 (BASELINE configs 2 and 5): the k smallest entries of
    each row of ``cdist(Q, C, metric)`` ascending, ties broken by lowest index.
    Same float64 per-pair arithmetic as the reference's cdist/pdist calls; the
    full-row ``sort`` of cluster.py:496 is replaced by argpartition + sort of the
    survivors, which returns the same k values."""
    nq = q.shape[0]
    k = min(k, c.shape[0])
    out_i = np.empty((nq, k), dtype=np.int64)
    out_d = np.empty((nq, k), dtype=np.float64)
    for s in range(0, nq, chunk):
        dm = cdist(q[s:s + chunk], c, metric)
        if k < c.shape[0]:
            part = np.argpartition(dm, k - 1, axis=1)[:, :k]
            kth = np.take_along_axis(dm, part, axis=1).max(axis=1)
        else:
            kth = dm.max(axis=1)
        for r in range(dm.shape[0]):
            row = dm[r]
            cand = np.flatnonzero(row <= kth[r])
            order = np.lexsort((cand, row[cand]))[:k]
            out_i[s + r] = cand[order]
            out_d[s + r] = row[cand[order]]
    return out_d, out_i


# ---------------------------------------------------------------- a8 / a9

def k_dist(bin_k_vals: np.ndarray):
    """RefineEngine.kDist, refine.py:1490-1496."""
    kd = pdist(bin_k_vals, 'cityblock')
    if len(kd) > 0:
        return float(np.median(kd))
    return None


def k_dist_between(k1: np.ndarray, k2: np.ndarray) -> float:
    """RefineEngine.kDistBetweenBins, refine.py:1502-1503."""
    return float(np.median(cdist(k1, k2, 'cityblock')))


def small2indices(index: int, side: int):
    """mstore.py:1880-1888: condensed pdist index -> (i, j)."""
    step = 0
    while index >= (side - step):
        index = index - side + step
        step += 1
    return (step, step + index + 1)


# ---------------------------------------------------------------- parity helper

def knn_tie_aware_equal(q, c, idx_engine, idx_oracle, dist_oracle, metric="euclidean",
                        tie: float = 1e-6):
    """Index parity rule of BASELINE.json's north_star: indices are identical to
    the oracle's except where the competing oracle distances are within ``tie``.

    A rank where the engine's index differs is excused only if the float64
    distance from the query to the index the engine chose equals the oracle's
    distance at that rank within ``tie`` (i.e. it is a genuine tie), and no row
    may repeat an index.  Returns the number of violating rows (0 == parity)."""
    idx_engine = np.asarray(idx_engine)
    idx_oracle = np.asarray(idx_oracle)
    bad = 0
    for r in np.flatnonzero((idx_engine != idx_oracle).any(axis=1)):
        cols = np.flatnonzero(idx_engine[r] != idx_oracle[r])
        if len(set(idx_engine[r].tolist())) != idx_engine.shape[1]:
            bad += 1
            continue
        d = cdist(q[r:r + 1], c[idx_engine[r, cols]], metric)[0]
        if np.any(np.abs(d - dist_oracle[r, cols]) > tie):
            bad += 1
    return bad


# ---------------------------------------------------------------- f1: SOM training

def som_stamp(radius_decaying: float, influence_rate: float) -> np.ndarray:
    """The square "stamp" of neighbourhood influences of one training iteration,
    som.py:282-318: linear decay -influence/radius * dist + influence inside the rounded
    radius, axis entries halved before the four quadrants are overlaid, all-zero borders
    culled."""
    grad = -1 * influence_rate / radius_decaying
    max_radius = int(radius_decaying)
    q = np.zeros((max_radius + 1, max_radius + 1))
    for row in range(0, max_radius + 1):
        for col in range(0, max_radius + 1):
            true_dist = np.sqrt(row ** 2 + col ** 2)
            check_dist = np.round(true_dist + 0.00001)
            if check_dist <= radius_decaying:
                influence = true_dist * grad + influence_rate
                q[row, col] = influence
                q[col, row] = influence
    q[:, 0] /= 2
    q[0, :] /= 2
    stamp = np.zeros((2 * max_radius + 1, 2 * max_radius + 1))
    stamp[max_radius:, max_radius:] += q
    stamp[:max_radius + 1, max_radius:] += np.rot90(q, 1)
    stamp[:max_radius + 1, :max_radius + 1] += np.rot90(q, 2)
    stamp[max_radius:, :max_radius + 1] += np.rot90(q, 3)
    stamp[max_radius, max_radius] = influence_rate
    max_vals = np.max(stamp, axis=0)
    k = 0
    while k < len(max_vals):
        if max_vals[k] > 0.0:
            break
        k += 1
    if k < len(max_vals):
        stamp = stamp[k:2 * max_radius + 1 - k, k:2 * max_radius + 1 - k]
    return stamp


def som_train(nodes: np.ndarray, train_vectors: np.ndarray, weights=None, iterations: int = 1000,
              vector_subset: int = 1000, influence_rate: float = 0.4, mask=None, radius: float = 0.0):
    """SOM.train, som.py:214-392, as a function of the mesh ``nodes`` [rows, cols, D]
    (``self.weights.nodes`` when ``weights`` is None, else the caller's ``weights`` patch).

    Kept from the reference: the BMU search of every iteration runs on weights frozen for
    that iteration (:348-353); deltas are accumulated on a 3 x 3 tiling of the mesh and folded
    columns first, then rows (:366-370); clip to [0, 1] (:374); the subset shuffle draws from
    numpy's GLOBAL generator and shuffles the same index array again every iteration
    (:324-326); with caller-owned ``weights`` and no mask the BMU search keeps using the
    weights it was handed (``flat_nodes`` is only refreshed at :380 / :384); training stops
    when the decayed radius falls under 2 (:279-280).  ``mask`` is a uint8/bool [rows, cols]
    array of the nodes allowed to move (the reference passes a dict of those points).
    Returns the trained weights (a new array; the inputs are not modified)."""
    replace = weights is None
    w = np.array(nodes if replace else weights, dtype=np.float64, copy=True)
    rows, cols, dim = w.shape
    if radius == 0.0:
        radius = float(rows) / 2
    flat_nodes = w.reshape((rows * cols, dim))
    tv = np.asarray(train_vectors, dtype=np.float64)
    time_constant = iterations / np.log(radius)
    if len(tv) <= vector_subset:
        index_array = np.arange(len(tv))
        cut_off = len(tv)
    else:
        rand_index_array = np.arange(len(tv))
        cut_off = vector_subset
    mask_pts = None if mask is None else np.argwhere(np.asarray(mask) != 0)
    for i in range(1, iterations + 1):
        radius_decaying = radius * np.exp(-1.0 * i / time_constant)
        if radius_decaying < 2:
            return w
        stamp = som_stamp(radius_decaying, influence_rate)
        sr = int((len(stamp) - 1) / 2)
        if len(tv) > vector_subset:
            np.random.shuffle(rand_index_array)
            index_array = rand_index_array[:cut_off]
        sheet = np.tile(w, (3, 3, 1))
        delta = np.zeros_like(sheet)
        for j in index_array:
            loc = int(np.argmin(cdist(flat_nodes, [tv[j]])))
            row = int(loc / cols)
            col = loc - (row * cols)
            r0, r1 = rows + row - sr, rows + row + sr + 1
            c0, c1 = cols + col - sr, cols + col + sr + 1
            patch = -1 * (sheet[r0:r1, c0:c1] - tv[j])
            # som.py:360-363 scales dimensions 0..3 by name: a mesh of fewer than four dimensions raises
            # IndexError there, dimensions past the fourth keep the whole difference
            for dd in range(4):
                patch[:, :, dd] *= stamp
            delta[r0:r1, c0:c1] += patch
        delta[:, cols:2 * cols] += delta[:, 0:cols]
        delta[:, cols:2 * cols] += delta[:, 2 * cols:3 * cols]
        delta[rows:2 * rows, cols:2 * cols] += delta[0:rows, cols:2 * cols]
        delta[rows:2 * rows, cols:2 * cols] += delta[2 * rows:3 * rows, cols:2 * cols]
        fold = delta[rows:2 * rows, cols:2 * cols]
        if mask_pts is None:
            w = np.clip(w + fold, 0, 1)
            if replace:
                flat_nodes = w.reshape((rows * cols, dim))
        else:
            for r, c in mask_pts:
                w[r, c] = np.clip(w[r, c] + fold[r, c], 0, 1)
            if len(mask_pts):
                flat_nodes = w.reshape((rows * cols, dim))
    return w


# ---------------------------------------------------------------- f3 / f4

def _angle(a, b, na, nb):
    """One pair of refine.py:1545-1546 / 1577-1578 under np.seterr(all='raise')."""
    with np.errstate(all='raise'):
        try:
            return np.arccos(np.dot(a, b) / (na * nb))
        except FloatingPointError:
            return None


def _row_norms(cov_rows, norms):
    """``self.PM.normCoverages`` of the rows: an INPUT of the reference (loaded from its database,
    profileManager.py:205).  None: the per-row Euclidean norm, which is how the fixtures make it."""
    if norms is None:
        return [np.linalg.norm(r) for r in cov_rows]
    return list(np.asarray(norms, dtype=np.float64))


def angle_pairs(cov1, cov2, norms1=None, norms2=None):
    """Every pair of refine.py:1545-1546 / 1577-1578 -> (angle [m1, m2] with NaN where the reference
    drops the pair, cosine argument [m1, m2] as the reference's arithmetic produced it)."""
    n1, n2 = _row_norms(cov1, norms1), _row_norms(cov2, norms2)
    ang = np.full((len(cov1), len(cov2)), np.nan)
    arg = np.full((len(cov1), len(cov2)), np.nan)
    for i in range(len(cov1)):
        for j in range(len(cov2)):
            with np.errstate(all='ignore'):
                arg[i, j] = np.dot(cov1[i], cov2[j]) / (n1[i] * n2[j])
            a = _angle(cov1[i], cov2[j], n1[i], n2[j])
            if a is not None:
                ang[i, j] = a
    return ang, arg


def c_dist(cov_rows: np.ndarray, norms=None):
    """RefineEngine.cDist, refine.py:1528-1551, on the selected rows; ``norms`` =
    ``normCoverages`` of those rows."""
    norms = _row_norms(cov_rows, norms)
    angles = []
    for i in range(len(cov_rows)):
        for j in range(i + 1, len(cov_rows)):
            a = _angle(cov_rows[i], cov_rows[j], norms[i], norms[j])
            if a is not None:
                angles.append(a)
    return float(np.median(angles)) if angles else float("nan")


def c_dist_between(cov1: np.ndarray, cov2: np.ndarray, norms1=None, norms2=None):
    """RefineEngine.cDistBetweenBins, refine.py:1557-1583, on the selected rows."""
    n1 = _row_norms(cov1, norms1)
    n2 = _row_norms(cov2, norms2)
    angles = []
    for i in range(len(cov1)):
        for j in range(len(cov2)):
            a = _angle(cov1[i], cov2[j], n1[i], n2[j])
            if a is not None:
                angles.append(a)
    return float(np.median(angles)) if angles else float("nan")


def evenly_spaced_pts_z(transformed_cp: np.ndarray, row_indices, sample_size: int):
    """RefineEngine.getEvenlySpacedPtsZ, refine.py:1505-1516: ``sample_size`` rows spread evenly
    along the Z axis (last column) of the transformed coverage space.  The positions are the
    truncations of a RUNNING float sum of len/sample_size, as in the reference."""
    rows = np.asarray(row_indices)
    z_rank = np.argsort(transformed_cp[rows, -1])
    stride = float(len(rows)) / sample_size
    chosen, cursor = [], 0.0
    for _ in range(sample_size):
        chosen.append(rows[z_rank[int(cursor)]])
        cursor += stride
    return np.array(chosen)


def _c_pick(transformed_cp, row_indices, max_in_bin=100):
    """The row selection at the head of cDist / cDistBetweenBins (refine.py:1529-1536, 1558-1568)."""
    row_indices = np.asarray(row_indices)
    if len(row_indices) > max_in_bin:
        return evenly_spaced_pts_z(transformed_cp, row_indices, max_in_bin)
    return row_indices


def _pick_norms(norm_coverages, pick):
    return None if norm_coverages is None else np.asarray(norm_coverages, dtype=np.float64)[pick]


def c_dist_rows(cov_profiles, transformed_cp, row_indices, norm_coverages=None):
    """RefineEngine.cDist (refine.py:1528-1551) from the row indices of a bin."""
    pick = _c_pick(transformed_cp, row_indices)
    return c_dist(np.asarray(cov_profiles)[pick], _pick_norms(norm_coverages, pick))


def c_dist_between_rows(cov_profiles, transformed_cp, rows1, rows2, norm_coverages=None):
    """RefineEngine.cDistBetweenBins (refine.py:1557-1583) from the row indices of two bins."""
    cov = np.asarray(cov_profiles)
    p1, p2 = _c_pick(transformed_cp, rows1), _c_pick(transformed_cp, rows2)
    return c_dist_between(cov[p1], cov[p2], _pick_norms(norm_coverages, p1), _pick_norms(norm_coverages, p2))


def c_dist_merged_rows(cov_profiles, transformed_cp, rows1, rows2, norm_coverages=None):
    """RefineEngine.cDistMergedBins (refine.py:1553-1555)."""
    return c_dist_rows(cov_profiles, transformed_cp, np.concatenate((rows1, rows2)), norm_coverages)


def vq(obs: np.ndarray, code_book: np.ndarray):
    """``scipy.cluster.vq.vq`` as called at binManager.py:564."""
    from scipy.cluster.vq import vq as sp_vq

    return sp_vq(obs, code_book)
