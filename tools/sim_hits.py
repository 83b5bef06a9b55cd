"""CPU study for the round-2 plan of the tensor-core scan (no GPU needed).

For one CTA's worth of queries (128 rows) of the cfg5 synthetic profiles it replays the scan's
streaming top-kk selection over candidate tiles of 128 and counts, per candidate ORDER:
  * hits per query (values that beat the query's running threshold when their tile arrives),
  * tiles in which any lane of a 32-query warp has a hit (the selection warps' slow path),
  * tiles a triangle-inequality bound on (query tile, candidate tile) would skip outright.
Orders: the rows as given (what the kernel does today), and rows sorted by a k-means cluster key
with the sweep starting at the query tile's own position.

    python tools/sim_hits.py [N] [D] [k] [n_clusters]
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from groopm_b200 import synth  # noqa: E402


def kmeans_key(x, nc, iters=6, seed=0):
    rng = np.random.default_rng(seed)
    samp = x[rng.choice(len(x), min(len(x), 50 * nc), replace=False)]
    cent = samp[rng.choice(len(samp), nc, replace=False)].copy()
    for _ in range(iters):
        lab = assign(samp, cent)
        for c in range(nc):
            m = samp[lab == c]
            if len(m):
                cent[c] = m.mean(axis=0)
    return assign(x, cent), cent


def assign(x, cent, chunk=20000):
    out = np.empty(len(x), dtype=np.int64)
    c2 = (cent * cent).sum(1)
    for a in range(0, len(x), chunk):
        xs = x[a:a + chunk]
        d = c2[None, :] - 2.0 * xs @ cent.T
        out[a:a + chunk] = d.argmin(1)
    return out


def replay(q, c, kk, start_tile=0, prune=True, check=False):
    """q [128, D], c [Nc, D] in the sweep's order -> hits per query, firing tiles per warp, skipped tiles."""
    tm = 128
    nt = (len(c) + tm - 1) // tm
    best = np.full((len(q), kk), np.inf)            # running top-kk squared distances
    thr = np.full(len(q), np.inf)
    hits = np.zeros(len(q), dtype=np.int64)
    fire = np.zeros(len(q) // 32, dtype=np.int64)
    skipped = 0
    qc = q.mean(0)
    qr = np.sqrt(((q - qc) ** 2).sum(1)).max()
    q2 = (q * q).sum(1)
    for i in range(nt):
        t = (start_tile + i) % nt
        ct = c[t * tm:(t + 1) * tm]
        if prune and np.isfinite(thr).all():
            cc = ct.mean(0)
            cr = np.sqrt(((ct - cc) ** 2).sum(1)).max()
            lb = np.sqrt(((qc - cc) ** 2).sum()) - qr - cr
            if lb > 0 and lb * lb > thr.max():
                skipped += 1
                continue
        d = q2[:, None] - 2.0 * q @ ct.T + (ct * ct).sum(1)[None, :]
        h = d < thr[:, None]
        hits += h.sum(1)
        fire += h.reshape(len(q) // 32, 32, -1).any(axis=(1, 2))
        if h.any():
            rows = np.flatnonzero(h.any(1))
            merged = np.concatenate([best[rows], np.where(h[rows], d[rows], np.inf)], axis=1)
            merged.partition(kk - 1, axis=1)
            best[rows] = merged[:, :kk]
            thr[rows] = best[rows].max(1)
    if check:
        # the bound is only sound if the kept top-kk is the brute-force one
        d_all = q2[:, None] - 2.0 * q @ c.T + (c * c).sum(1)[None, :]
        want = np.sort(np.partition(d_all, kk - 1, axis=1)[:, :kk], axis=1)
        assert np.allclose(np.sort(best, axis=1), want, rtol=0, atol=1e-9 * max(1.0, np.abs(want).max())), "pruning lost a neighbour"
    return hits, fire, skipped, nt


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    nc = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
    kk = k + 8
    t0 = time.time()
    x = synth.cfg_knn_rows(n, d, 5).astype(np.float64)
    x -= x.mean(0)
    print(f"cfg5 rows {n} x {d}, k = {k} (kk = {kk}); built in {time.time() - t0:.1f} s", flush=True)
    rng = np.random.default_rng(1)
    for qt in rng.choice(n // 128, 3, replace=False):
        q = x[qt * 128:(qt + 1) * 128]
        hits, fire, skipped, nt = replay(q, x, kk, 0, prune=False)
        print(f"as given, query tile {qt}: hits per query mean {hits.mean():.0f} max {hits.max()}, "
              f"tiles with a hit per warp {fire.mean():.0f} of {nt}", flush=True)
    t0 = time.time()
    lab, cent = kmeans_key(x, nc)
    order = np.argsort(lab, kind="stable")
    xs = x[order]
    print(f"k-means key ({nc} clusters, 6 Lloyd steps on a sample): {time.time() - t0:.1f} s", flush=True)
    for qt in rng.choice(n // 128, 6, replace=False):
        q = xs[qt * 128:(qt + 1) * 128]
        for prune in (False, True):
            hits, fire, skipped, nt = replay(q, xs, kk, qt, prune=prune, check=prune)
            print(f"cluster order, sweep from own tile {qt}, prune={prune}: hits per query mean {hits.mean():.0f} max {hits.max()}, "
                  f"tiles with a hit per warp {fire.mean():.0f}, tiles skipped by the bound {skipped} of {nt}", flush=True)


if __name__ == "__main__":
    main()
