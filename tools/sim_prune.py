"""CPU study of tile pruning for the ordered tensor-core scan (no GPU needed).

Models what scan_tc_kernel does with pruning on:
  staging order  = counting sort by key = nearest pivot; pivots = `ncl` strided sample rows, plus
                   (two-round) `ncl2` more drawn from the rows whose distance to their pivot
                   exceeds `far` x the median distance; every key group is padded to whole
                   128-row tiles, so a tile holds rows of ONE pivot;
  ball per tile  = centroid + largest member distance;
  rounds         = the CTA of a query tile first visits a window of W tiles around its own position,
                   then, repeatedly, the unvisited tiles whose bound |cq - cc| - rq - rc is at most
                   tau, tau = 0, T/2, T (T = the tile's largest running threshold, re-read after
                   every round); it stops when nothing unvisited is within T.
Reports visited tiles per query tile and checks that no neighbour was lost.

    python tools/sim_prune.py [N] [D] [k] [ncl] [ncl2] [W]
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from groopm_b200 import synth  # noqa: E402

TM = 128


def assign(x, cent, chunk=20000):
    key = np.empty(len(x), dtype=np.int64)
    dist = np.empty(len(x))
    c2 = (cent * cent).sum(1)
    for a in range(0, len(x), chunk):
        xs = x[a:a + chunk]
        d = c2[None, :] - 2.0 * xs @ cent.T
        key[a:a + chunk] = d.argmin(1)
        dist[a:a + chunk] = np.sqrt(np.maximum(d.min(1) + (xs * xs).sum(1), 0.0))
    return key, dist


def keys_two_round(x, ncl, ncl2, far=2.0):
    n = len(x)
    key, dist = assign(x, x[::n // ncl][:ncl])
    if ncl2 > 0:
        med = np.median(dist)
        unc = np.flatnonzero(dist > far * med)
        if len(unc) >= ncl2:
            piv2 = x[unc[::len(unc) // ncl2][:ncl2]]
            k2, d2 = assign(x[unc], piv2)
            better = d2 < dist[unc]
            key[unc[better]] = ncl + k2[better]
            dist[unc[better]] = d2[better]
            print(f"  round 2: {len(unc)} rows beyond {far} x median {med:.3f}; {better.sum()} re-keyed", flush=True)
    return key, dist


def padded_order(key, nkeys):
    """perm with -1 holes: every key group starts on a tile boundary"""
    cnt = np.bincount(key, minlength=nkeys)
    size = (cnt + TM - 1) // TM * TM
    start = np.concatenate([[0], np.cumsum(size)])
    perm = np.full(start[-1], -1, dtype=np.int64)
    order = np.argsort(key, kind="stable")
    pos_in = np.arange(len(key)) - np.repeat(np.concatenate([[0], np.cumsum(cnt)])[:-1], cnt)
    perm[start[key[order]] + pos_in] = order
    return perm


def main():
    a = sys.argv[1:]
    n = int(a[0]) if len(a) > 0 else 1_000_000
    d = int(a[1]) if len(a) > 1 else 30
    k = int(a[2]) if len(a) > 2 else 32
    ncl = int(a[3]) if len(a) > 3 else 1024
    ncl2 = int(a[4]) if len(a) > 4 else 512
    w = int(a[5]) if len(a) > 5 else 8
    seed = 5 if d == 30 else (2 if d == 10 else 4)
    kk = k + min(8, max(3, k // 4))
    x = synth.cfg_knn_rows(n, d, seed) if d != 136 else synth.kmer_sigs(n, seed)
    x = x - x.mean(0)
    t0 = time.time()
    key, dist = keys_two_round(x, ncl, ncl2)
    perm = padded_order(key, ncl + ncl2)
    nt = len(perm) // TM
    real = perm >= 0
    xs = np.where(real[:, None], x[np.maximum(perm, 0)], np.nan)
    print(f"{n} x {d} k={k} kk={kk} pivots {ncl}+{ncl2} W={w}: {nt} tiles ({nt * TM / n - 1:.1%} padding); keys in {time.time() - t0:.1f} s", flush=True)
    tiles = xs.reshape(nt, TM, d)
    cen = np.nan_to_num(np.nanmean(tiles, axis=1))
    rad = np.nan_to_num(np.sqrt(np.nanmax(((tiles - cen[:, None, :]) ** 2).sum(2), axis=1)))
    print("tile radius percentiles 10/50/90/99:", np.percentile(rad, [10, 50, 90, 99]).round(2), flush=True)
    rng = np.random.default_rng(1)
    visited_n, rounds_n = [], []
    x2 = (x * x).sum(1)
    for qt in rng.choice(nt, 40, replace=False):
        q = tiles[qt][real[qt * TM:(qt + 1) * TM]]
        if len(q) == 0:
            continue
        lb = np.sqrt(((cen - cen[qt]) ** 2).sum(1)) - rad - rad[qt]
        visited = np.zeros(nt, dtype=bool)
        best = np.full((len(q), kk), np.inf)

        def visit(mask):
            nonlocal best
            rows = perm.reshape(nt, TM)[mask].ravel()
            rows = rows[rows >= 0]
            if len(rows) == 0:
                return
            c = x[rows]
            d2 = (q * q).sum(1)[:, None] - 2.0 * q @ c.T + x2[rows][None, :]
            merged = np.concatenate([best, d2], axis=1)
            merged.partition(kk - 1, axis=1)
            best = merged[:, :kk]
            visited[mask] = True

        win = np.zeros(nt, dtype=bool)
        win[[(qt + i - w // 2) % nt for i in range(w)]] = True
        visit(win)
        r = 1
        for frac in (0.0, 0.5, 1.0, 1.0, 1.0):
            t_max = np.sqrt(max(best.max(), 0.0))
            m = (~visited) & (lb <= frac * t_max)
            if frac == 1.0 and not m.any():
                break
            if m.any():
                visit(m)
                r += 1
        # nothing unvisited may hold a neighbour
        d_all = (q * q).sum(1)[:, None] - 2.0 * q @ x.T + x2[None, :]
        want = np.sort(np.partition(d_all, kk - 1, axis=1)[:, :kk], axis=1)
        assert np.allclose(np.sort(best, axis=1), want, rtol=0, atol=1e-9 * max(1.0, want.max())), "pruning lost a neighbour"
        visited_n.append(visited.sum())
        rounds_n.append(r)
    v = np.array(visited_n)
    print(f"tiles visited per query tile: mean {v.mean():.0f} median {np.median(v):.0f} max {v.max()} of {nt} "
          f"-> {v.mean() / nt:.2%} of the tile pairs; rounds mean {np.mean(rounds_n):.1f}; no neighbour lost", flush=True)
    print("per sampled query tile:", np.sort(v).tolist())


if __name__ == "__main__":
    main()
