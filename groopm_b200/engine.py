"""Host side of the distance / neighbour-search engine.

numpy or torch float64 arrays in, numpy or torch arrays out, in the layouts the
reference's call sites consume.  PyTorch is used only for
device memory, streams and (in ``dist.py``) torch.distributed; every kernel is
reached through the C ABI of ``libgmdist.so`` with raw device pointers.
"""
from __future__ import annotations

import os

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib

METRICS = {"euclidean": _lib.L2, "cityblock": _lib.L1, "l2": _lib.L2, "l1": _lib.L1,
           _lib.L2: _lib.L2, _lib.L1: _lib.L1}


def _metric(metric) -> int:
    try:
        return METRICS[metric]
    except KeyError:
        raise ValueError(f"Unknown Distance Metric: {metric}") from None


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.GmdistError("no CUDA device: the gmdist engine has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _as_rows(x, name: str) -> tuple[torch.Tensor, bool]:
    """-> (2-D float64 C-contiguous torch tensor on its current device, was_numpy)."""
    was_numpy = not isinstance(x, torch.Tensor)
    if was_numpy:
        a = np.asarray(x, dtype=np.float64)
        if a.ndim != 2:
            raise ValueError(f"{name} must be a 2-dimensional array.")
        t = torch.from_numpy(np.ascontiguousarray(a))
    else:
        t = x
        if t.ndim != 2:
            raise ValueError(f"{name} must be a 2-dimensional array.")
        if t.dtype != torch.float64:
            t = t.to(torch.float64)
        t = t.contiguous()
    return t, was_numpy


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


class KnnPipeline:
    """The staged device pipeline for one (Nq, Nc, D, k, metric) shape.

    Owns every device buffer (the library never allocates on this path):

        center [D] f64 | stats [4] f32 | flags [4] i32
        ctiles / qtiles  FP32 profile tiles  [ntiles][DP+1][64]
        cand_idx / cand_val                  [Nq][nsplit][kk]   proposals of the FP32 scan
        cand_tau                             [Nq][nsplit]       bound on everything not proposed
        out_idx [Nq, k] i32, out_dist [Nq, k] f64

    ``run`` = col_mean -> prep -> scan -> refine -> exact_rows, all stream-ordered.
    """

    def __init__(self, nq: int, nc: int, d: int, k: int, metric="euclidean", form: Optional[int] = None,
                 same: bool = False, device=None, order: Optional[bool] = None, prune: Optional[bool] = None):
        self.lib = _lib.lib()
        self.dev = _device(device)
        self.nq, self.nc, self.d, self.k = int(nq), int(nc), int(d), int(k)
        self.metric = _metric(metric)
        if self.nq <= 0 or self.nc <= 0 or self.d <= 0:
            raise ValueError("empty profile set")
        if self.d > _lib.MAX_D:
            raise ValueError(f"D={self.d} exceeds {_lib.MAX_D}")
        if not 1 <= self.k <= _lib.MAX_K:
            raise ValueError(f"k={self.k} outside [1, {_lib.MAX_K}]")
        if self.k > self.nc:
            raise ValueError(f"k={self.k} exceeds the {self.nc} candidates")
        if form is None:
            # large Euclidean problems go to the tensor-core scan, the rest to the FP32 forms
            form = (_lib.FORM_TC if self.lib.gmd_use_tc_scan(self.nq, self.nc, self.k, self.metric, self.d)
                    else self.lib.gmd_pick_form(self.metric, self.d))
            auto = True
        else:
            auto = False
        self.form = int(form)
        self.tc = self.form == _lib.FORM_TC
        if self.tc and not self.lib.gmd_tc_supported(self.k, self.metric, self.d):
            raise ValueError("the tensor-core scan needs the Euclidean metric, 2 <= k <= 64 and 4 <= D <= 136")
        self.same = bool(same)
        # k = 1 on short profiles (BMU search) has its own heap-free scan
        self.bmu = bool(auto and self.lib.gmd_use_bmu_scan(self.k, self.metric, self.d))
        if self.tc:
            self.bmu = False
        with torch.cuda.device(self.dev):
            if self.bmu:
                self.kk, self.nsplit = self.lib.gmd_bmu_kk(), 1
            elif self.tc:
                self.kk, self.nsplit = self.lib.gmd_pick_kk(self.k), 1
            else:
                self.kk = self.lib.gmd_pick_kk(self.k)
                self.nsplit = self.lib.gmd_pick_nsplit(self.nq, self.nc, self.d, self.kk)
        e = dict(device=self.dev)
        self.center = torch.zeros(_lib.MAX_D, dtype=torch.float64, **e)
        self.stats = torch.zeros(4, dtype=torch.float32, **e)
        self.flags = torch.zeros(4, dtype=torch.int32, **e)
        # Pruned tensor-core scan (the default wherever gmd_knn_f64 uses it: both sets large): profiles
        # staged in the grouped order (leaders from a sample, nearest-leader keys, groups padded to whole
        # tiles), a ball per tile and group, and every query tile loads only the candidate tiles its
        # running thresholds cannot rule out.  prune=False: the full sweep (plain or `order`ed).
        if prune is None:
            prune = bool(self.tc and auto and self.lib.gmd_use_tc_prune(self.nq, self.nc, self.k, self.metric, self.d))
        self.prune = bool(prune and self.tc)
        if self.prune and not self.lib.gmd_tc_prune_max_tiles(self.d, self.kk):
            raise ValueError("the pruned tensor-core scan does not fit this (D, k)")
        if self.prune:
            order = False
        tiles_bytes = ((lambda n: self.lib.gmd_tc_grouped_tiles_bytes(n, self.nc, self.d)) if self.prune
                       else (lambda n: self.lib.gmd_tc_tiles_bytes(n, self.d)) if self.tc
                       else (lambda n: self.lib.gmd_tiles_bytes(n, self.d)))
        self.ctiles = torch.empty(tiles_bytes(self.nc) // 4, dtype=torch.float32, **e)
        # tensor-core tiles differ by role (column D), so queries are always staged separately there
        self.qtiles = self.ctiles if (same and not self.tc) else torch.empty(
            tiles_bytes(self.nq) // 4, dtype=torch.float32, **e)
        m = self.nq * self.nsplit * self.kk
        self.cand_idx = torch.empty(m, dtype=torch.int32, **e)
        self.cand_val = torch.empty(m, dtype=torch.float32, **e)
        self.cand_tau = torch.empty(self.nq * self.nsplit, dtype=torch.float32, **e)
        self.fail_rows = torch.empty(self.nq, dtype=torch.int32, **e)
        self.exact_ws = torch.empty(self.lib.gmd_exact_workspace_bytes(self.nq) // 4, dtype=torch.int32, **e)
        # Locality-ordered sweep of the tensor-core scan (all-pairs only; opt-in: order=True or
        # GMD_TC_ORDER=1).  Profiles are grouped by their nearest of ~N/1024 sampled profiles (an
        # inner k = 1 pipeline of this engine), staged in that order, and every query tile starts
        # its sweep at its own position; proposals come back under the original row numbers.
        if order is None:
            order = os.environ.get("GMD_TC_ORDER", "0") == "1"
        self.order = bool(order and self.tc and self.same and self.nc >= 4096)
        if self.order:
            ncl = 64
            while ncl < 4096 and ncl * 1024 < self.nc:
                ncl *= 2
            self.order_ncl = ncl
            self.key_pipe = KnnPipeline(self.nc, ncl, self.d, 1, "euclidean", None, same=False, device=self.dev,
                                        order=False)
            self.perm = torch.empty(self.nc, dtype=torch.int32, **e)
            self.order_ws = torch.empty(self.lib.gmd_order_workspace_bytes(ncl) // 4, dtype=torch.int32, **e)
        if self.prune:
            self.grouped_ws = torch.empty(self.lib.gmd_tc_grouped_workspace_bytes(self.nq, self.nc, self.d),
                                          dtype=torch.uint8, **e)
            self.grouped = _lib.Grouped()
        # row-sharded callers (dist.py) hand in a function that finds the nearest leader of all rows with the work
        # split over the ranks: lead_rows [nlead, D] -> (index int32 [nc], distance float64 [nc])
        self.shard_assign = None
        self.scan_events = None     # a list: run() appends a (start, end) CUDA event pair around every scan launch
        self.out_idx = torch.empty((self.nq, self.k), dtype=torch.int32, **e)
        self.out_dist = torch.empty((self.nq, self.k), dtype=torch.float64, **e)

    # -- stages (each a single C-ABI call on the current stream)
    def reset(self):
        _lib.check(self.lib.gmd_zero_async(_ptr(self.stats), 16, _stream_ptr()))
        _lib.check(self.lib.gmd_zero_async(_ptr(self.flags), 16, _stream_ptr()))

    def col_mean(self, c: torch.Tensor):
        _lib.check(self.lib.gmd_col_mean_f64(_ptr(c), self.nc, self.d, _ptr(self.center), _stream_ptr()))

    def prep_candidates(self, c: torch.Tensor):
        if self.tc:
            # pass 1 of the tensor-core staging: the largest centred row norm fixes the power-of-two scale
            _lib.check(self.lib.gmd_tc_norms_f64(_ptr(c), self.nc, self.d, _ptr(self.center), _ptr(self.stats),
                                                 _ptr(self.flags), _stream_ptr()))
            self._tc_c = c
            return
        _lib.check(self.lib.gmd_prep_f64(_ptr(c), self.nc, self.d, _ptr(self.center), self.form,
                                         _ptr(self.ctiles), _ptr(self.stats), _ptr(self.flags), _stream_ptr()))

    def prep_queries(self, q: torch.Tensor):
        if self.tc:
            sq = _ptr(self.stats) + 8
            if not self.same:
                _lib.check(self.lib.gmd_tc_norms_f64(_ptr(q), self.nq, self.d, _ptr(self.center), sq, _ptr(self.flags),
                                                     _stream_ptr()))
            # pass 2 for both roles, now that both sets' norms are known
            c = self._tc_c
            if self.prune and self.shard_assign is not None:
                # the same staging in two halves: leaders here, nearest leaders by the caller (every rank its
                # own rows, all-gathered), the rest here
                nlead = C.c_int32(0)
                lead_ptr = C.c_void_p(0)
                _lib.check(self.lib.gmd_tc_grouped_leaders_f64(
                    _ptr(c), self.nq, self.nc, self.d, _ptr(self.center), _ptr(self.grouped_ws),
                    self.grouped_ws.numel(), C.byref(nlead), C.byref(lead_ptr), _stream_ptr()))
                off = lead_ptr.value - self.grouped_ws.data_ptr()
                lead = self.grouped_ws[off:off + nlead.value * self.d * 8].view(torch.float64).view(nlead.value, self.d)
                near_idx, near_dist = self.shard_assign(lead)
                _lib.check(self.lib.gmd_tc_stage_grouped_ext_f64(
                    _ptr(q), self.nq, _ptr(c), self.nc, self.d, _ptr(self.center), _ptr(self.stats), sq,
                    _ptr(self.qtiles), _ptr(self.ctiles), _ptr(self.grouped_ws), self.grouped_ws.numel(),
                    nlead.value, _ptr(near_idx), _ptr(near_dist), C.byref(self.grouped), _stream_ptr()))
                self._tc_c = None
                return
            if self.prune:
                # leaders, nearest-leader keys, padded counting sort, tiles + balls: one C call (it
                # synchronises the stream once, to learn the number of leaders)
                _lib.check(self.lib.gmd_tc_stage_grouped_f64(
                    _ptr(q), self.nq, _ptr(c), self.nc, self.d, _ptr(self.center), _ptr(self.stats), sq,
                    _ptr(self.qtiles), _ptr(self.ctiles), _ptr(self.grouped_ws), self.grouped_ws.numel(),
                    C.byref(self.grouped), _stream_ptr()))
                self._tc_c = None
                return
            perm = 0
            if self.order:
                # staging order: nearest sampled profile as the key, counting sort by key
                stride = self.nc // self.order_ncl
                cent = c[::stride][:self.order_ncl].contiguous()
                key, _ = self.key_pipe.run(c, cent)
                _lib.check(self.lib.gmd_order_by_key(_ptr(key), self.nc, self.order_ncl, _ptr(self.perm),
                                                     _ptr(self.order_ws), self.order_ws.numel() * 4, _stream_ptr()))
                perm = _ptr(self.perm)
            _lib.check(self.lib.gmd_prep_tc_perm_f64(_ptr(c), self.nc, self.d, _ptr(self.center), 1, _ptr(self.stats),
                                                     sq, perm, _ptr(self.ctiles), _stream_ptr()))
            _lib.check(self.lib.gmd_prep_tc_perm_f64(_ptr(q), self.nq, self.d, _ptr(self.center), 0, _ptr(self.stats),
                                                     sq, perm, _ptr(self.qtiles), _stream_ptr()))
            self._tc_c = None
            return
        if self.same:
            return
        _lib.check(self.lib.gmd_prep_f64(_ptr(q), self.nq, self.d, _ptr(self.center), self.form,
                                         _ptr(self.qtiles), _ptr(self.stats) + 8, _ptr(self.flags), _stream_ptr()))

    def scan(self):
        if self.prune:
            _lib.check(self.lib.gmd_scan_tc_grouped_f32(
                _ptr(self.qtiles), _ptr(self.ctiles), self.d, self.kk, 0, _ptr(self.stats), _ptr(self.stats) + 8,
                C.byref(self.grouped), _ptr(self.cand_idx), _ptr(self.cand_val), _ptr(self.cand_tau), _stream_ptr()))
            return
        if self.tc:
            _lib.check(self.lib.gmd_scan_tc_sweep_f32(
                _ptr(self.qtiles), self.nq, _ptr(self.ctiles), self.nc, self.d, self.kk, 0, _ptr(self.stats),
                _ptr(self.stats) + 8, _ptr(self.perm) if self.order else 0, 1 if self.order else 0,
                _ptr(self.cand_idx), _ptr(self.cand_val), _ptr(self.cand_tau), _stream_ptr()))
            return
        if self.bmu:
            _lib.check(self.lib.gmd_scan_bmu_f32(_ptr(self.qtiles), self.nq, _ptr(self.ctiles), self.nc, self.d, 0,
                                                 _ptr(self.cand_idx), _ptr(self.cand_val), _ptr(self.cand_tau),
                                                 _stream_ptr()))
            return
        _lib.check(self.lib.gmd_scan_topk_f32(_ptr(self.qtiles), self.nq, _ptr(self.ctiles), self.nc, self.d,
                                              self.form, self.kk, self.nsplit, 0, _ptr(self.cand_idx),
                                              _ptr(self.cand_val), _ptr(self.cand_tau), _stream_ptr()))

    def refine(self, q: torch.Tensor, c: torch.Tensor, out=None):
        # after a pruned scan the queries are re-ranked in their staging order (candidate rows stay in L2)
        qorder, nqpos = (self.grouped.perm_q, self.grouped.npos_q) if self.prune else (0, 0)
        out_idx, out_dist = out or (self.out_idx, self.out_dist)
        _lib.check(self.lib.gmd_refine_topk_order_f64(
            _ptr(q), self.nq, _ptr(c), self.nc, self.d, _ptr(self.center), self.k, self.kk, self.nsplit,
            self.metric, self.form, _ptr(self.cand_idx), _ptr(self.cand_val), _ptr(self.cand_tau), _ptr(self.stats),
            _ptr(out_idx), _ptr(out_dist), _ptr(self.flags) + 4, _ptr(self.fail_rows), qorder, nqpos,
            _stream_ptr()))

    def exact_rows(self, q: torch.Tensor, c: torch.Tensor, out=None):
        out_idx, out_dist = out or (self.out_idx, self.out_dist)
        _lib.check(self.lib.gmd_exact_rows_f64(
            _ptr(q), self.nq, _ptr(c), self.nc, self.d, self.k, self.metric, _ptr(self.flags) + 4,
            _ptr(self.fail_rows), _ptr(out_idx), _ptr(out_dist), _ptr(self.exact_ws),
            self.exact_ws.numel() * 4, _stream_ptr()))

    def run_next(self, q: torch.Tensor, c: torch.Tensor, out=None):
        """The pipeline again for ANOTHER run of ``c``'s rows of the same length (``q = c[r0:r0 + nq]``) after a
        ``run`` whose queries were such a run: the candidates' staging is kept, only the query side is staged
        (gmd_tc_restage_queries_f64), then scan -> refine -> exact rows into ``out`` (default: the pipeline's)."""
        if not self.prune or self.same:
            raise ValueError("run_next needs the pruned scan with queries that are a run of the candidates' rows")
        if tuple(q.shape) != (self.nq, self.d) or q.dtype != torch.float64 or not q.is_contiguous():
            raise ValueError(f"q: expected contiguous float64 [{self.nq}, {self.d}]")
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.gmd_zero_async(_ptr(self.flags) + 4, 4, _stream_ptr()))   # the rejected-row count
            _lib.check(self.lib.gmd_tc_restage_queries_f64(
                _ptr(q), self.nq, _ptr(c), self.nc, self.d, _ptr(self.center), _ptr(self.stats), _ptr(self.stats) + 8,
                _ptr(self.qtiles), _ptr(self.grouped_ws), self.grouped_ws.numel(), C.byref(self.grouped),
                _stream_ptr()))
            self.scan()
            self.refine(q, c, out)
            self.exact_rows(q, c, out)
        return out or (self.out_idx, self.out_dist)

    def run(self, q: torch.Tensor, c: torch.Tensor, out=None):
        """q [Nq, D], c [Nc, D] float64 on this pipeline's device -> (out_idx, out_dist) device tensors."""
        for name, t, n in (("q", q, self.nq), ("c", c, self.nc)):
            if t.device != self.dev or t.dtype != torch.float64 or not t.is_contiguous() or tuple(t.shape) != (n, self.d):
                raise ValueError(f"{name}: expected contiguous float64 [{n}, {self.d}] on {self.dev}")
        with torch.cuda.device(self.dev):
            self.reset()
            self.col_mean(c)
            self.prep_candidates(c)
            self.prep_queries(q)
            if self.scan_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                self.scan()
                e1.record()
                self.scan_events.append((e0, e1))
            else:
                self.scan()
            self.refine(q, c, out)
            self.exact_rows(q, c, out)
        return out or (self.out_idx, self.out_dist)

    # -- diagnostics (synchronise)
    def exact_row_count(self) -> int:
        return int(self.flags[1].item())

    def saw_nonfinite(self) -> bool:
        return bool(self.flags[0].item())

    def prune_stats(self) -> dict:
        """Counters of the last pruned scan: candidate tiles loaded, over all query tiles, against the
        full sweep's (synchronises)."""
        if not self.prune:
            return {}
        torch.cuda.current_stream().synchronize()
        g = self.grouped
        off = g.counters - self.grouped_ws.data_ptr()
        c = [int(v) for v in self.grouped_ws[off:off + 128].view(torch.int64).tolist()]
        visited, ctas, groups = c[0], c[1], c[2]
        ntc = -(-g.npos_c // 128)
        return {"tiles_visited": visited, "query_tiles": ctas, "groups": g.ngroups,
                "candidate_tiles": ntc, "groups_surviving_per_query_tile": groups / max(ctas, 1),
                "visited_fraction": visited / max(1, ctas * ntc),
                # only a -DGMD_TC_PROBES build fills these: thread 0's SM cycles per CTA and phase
                "probes": {n: c[4 + i] / max(ctas, 1) for i, n in enumerate(
                    ("wait", "load", "select", "park", "drain", "drains", "entries", "longest"))}}

    # launches of the engine's own kernels in one run() (col_mean, prep x1/2, scan, refine, exact)
    @property
    def launches_per_run(self) -> int:
        # the exact stage is three kernels: batched pass, its finish, one-row-per-block remainder
        if self.prune:
            # col_mean, norms x1/2; leaders (samples, nearest sample, radius, 2 x (bit matrix, greedy pass), cover,
            # uncovered, leader rows: 10); assignment (col_mean, prep x2, BMU scan, pick: 5); grouped sort (keys,
            # histogram, padded scan, scatter, tile groups: 5) once or twice; prep x2, group balls; scan, refine,
            # exact x3.  (Queries that are no rows of the candidates add a second assignment, not counted.)
            return (2 if self.same else 3) + 10 + 5 + 5 + 2 + 1 + 5 + (0 if self.same else 5)
        if self.tc:
            # col_mean, norms x1/2, prep x2, scan, refine, exact x3 (+ the key pipeline and 3 sort kernels)
            extra = (self.key_pipe.launches_per_run + 3) if self.order else 0
            return (9 if self.same else 10) + extra
        return 7 if self.same else 8


# Pipelines own every device buffer of a shape (hundreds of MB at 1 M rows): the reference calls these entry
# points in loops over bins and iterations with one shape, so the last few are kept.
_PIPES: "dict[tuple, KnnPipeline]" = {}
_PIPES_MAX = 4


def _pipeline(nq, nc, d, k, metric, form, same, dev, order, prune) -> KnnPipeline:
    key = (int(nq), int(nc), int(d), int(k), _metric(metric), form, bool(same), str(dev), order, prune)
    pipe = _PIPES.pop(key, None)
    if pipe is None:
        pipe = KnnPipeline(nq, nc, d, k, metric, form, same=same, device=dev, order=order, prune=prune)
        while len(_PIPES) >= _PIPES_MAX:
            _PIPES.pop(next(iter(_PIPES)))
    _PIPES[key] = pipe                                      # most recently used last
    return pipe


def clear_cache():
    """Drop the cached pipelines (and their device buffers)."""
    _PIPES.clear()


def _host_copy(t: torch.Tensor) -> torch.Tensor:
    """Device tensor -> pinned host tensor, asynchronously on the current stream (torch keeps the pinned
    blocks it has handed out, so repeated calls of one shape do not page-lock memory again)."""
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    return h


def knn(q, c=None, k: int = 1, metric="euclidean", form: Optional[int] = None, device=None,
        return_info: bool = False, order: Optional[bool] = None, prune: Optional[bool] = None):
    """k nearest rows of ``c`` for every row of ``q``; ascending (distance, index).

    ``c=None`` means all-pairs within ``q`` (self included at rank 0, as in
    ``sort(squareform(pdist(X)))`` cluster.py:495-496 and ``KDTree.query``
    refine.py:664-665).  numpy in -> numpy out (int32 indices, float64 distances);
    CUDA tensors in -> CUDA tensors out.
    """
    qt, q_np = _as_rows(q, "XA")
    same = c is None or c is q
    ct, _ = (qt, q_np) if same else _as_rows(c, "XB")
    if qt.shape[1] != ct.shape[1]:
        raise ValueError("XA and XB must have the same number of columns "
                         "(i.e. feature dimension.)")
    if qt.shape[0] == 0 or ct.shape[0] == 0 or qt.shape[1] == 0:
        raise ValueError("empty profile set")
    dev = qt.device if qt.is_cuda else (ct.device if ct.is_cuda else _device(device))
    with torch.cuda.device(dev):
        qd = qt.to(dev, non_blocking=True)
        cd = qd if same else ct.to(dev, non_blocking=True)
        if (form is None and not return_info and qd.shape[0] <= SMALL_KNN_ROWS and 1 <= k <= cd.shape[0]
                and cd.shape[0] <= _lib.lib().gmd_knn_small_max_candidates()):
            # a few query rows against a small candidate set (a KDTree.query call per bin, refine.py:664-665):
            # one float64 kernel, no staging
            n = qd.shape[0]
            idx = torch.empty((n, k), dtype=torch.int32, device=dev)
            dist = torch.empty((n, k), dtype=torch.float64, device=dev)
            flag = torch.empty(1, dtype=torch.int32, device=dev)
            _lib.check(_lib.lib().gmd_knn_small_f64(_ptr(qd), n, _ptr(cd), cd.shape[0], qd.shape[1], int(k),
                                                    _metric(metric), _ptr(idx), _ptr(dist), _ptr(flag), _stream_ptr()))
            if q_np:
                idx_h, dist_h, flag_h = _host_copy(idx), _host_copy(dist), _host_copy(flag)
                torch.cuda.current_stream().synchronize()
                bad, idx, dist = int(flag_h[0]), idx_h.numpy(), dist_h.numpy()
            else:
                bad = int(flag.item())
            if bad:
                raise FloatingPointError("non-finite value in an input profile")
            return idx, dist
        if q_np and same and form is None and not return_info and order is None and prune is None:
            streamed = _knn_streamed(qd, k, metric, dev)
            if streamed is not None:
                return streamed
        pipe = _pipeline(qd.shape[0], cd.shape[0], qd.shape[1], k, metric, form, same, dev, order, prune)
        idx, dist = pipe.run(qd, cd)
        if q_np:
            idx_h, dist_h, flags_h = _host_copy(idx), _host_copy(dist), _host_copy(pipe.flags)
            torch.cuda.current_stream().synchronize()
            if int(flags_h[0]):
                raise FloatingPointError("non-finite value in an input profile")
            out = (idx_h.numpy(), dist_h.numpy())
        else:
            if pipe.saw_nonfinite():
                raise FloatingPointError("non-finite value in an input profile")
            out = (idx, dist)
        if return_info:
            return out + ({"exact_rows": pipe.exact_row_count(), "form": pipe.form, "kk": pipe.kk,
                           "nsplit": pipe.nsplit, "ordered": pipe.order, "pruned": pipe.prune,
                           **({"prune": pipe.prune_stats()} if pipe.prune else {})},)
        return out


STREAM_MIN_ROWS = 262144  # all-pairs problems on host arrays from this size on return their results in runs
STREAM_RUNS = 4


def _knn_streamed(xd: torch.Tensor, k: int, metric, dev, runs: int = STREAM_RUNS, min_rows: int = STREAM_MIN_ROWS):
    """All-pairs k-NN of device rows ``xd`` with the results going to pinned host arrays while the scan is still
    running: the rows are walked in ``runs`` runs (queries = a run of the candidates' rows, the candidates staged
    once), and the device-to-host copy of a run's indices and distances (12 k bytes per row: 384 MB at 1 M rows,
    k = 32 -- as long as the whole scan) overlaps the next run's scan and re-rank on a second stream.  Same values
    as one pass (every row's result is the exact one).  None if the shape is not one for the pruned scan."""
    n, d = xd.shape
    runs = int(runs)
    if runs < 2 or n < max(int(min_rows), 2 * runs):
        return None
    lib = _lib.lib()
    m = -(-n // runs)
    if not (lib.gmd_use_tc_scan(m, n, k, _metric(metric), d) and lib.gmd_use_tc_prune(m, n, k, _metric(metric), d)):
        return None
    pipe = _pipeline(m, n, d, k, metric, None, False, dev, None, None)
    if not pipe.prune:
        return None
    if getattr(pipe, "_stream_out", None) is None:
        pipe._stream_out = [(torch.empty_like(pipe.out_idx), torch.empty_like(pipe.out_dist)) for _ in range(2)]
        pipe._stream_copy = torch.cuda.Stream(device=dev)
    idx_h = torch.empty((n, k), dtype=torch.int32, pin_memory=True)
    dist_h = torch.empty((n, k), dtype=torch.float64, pin_memory=True)
    main = torch.cuda.current_stream()
    copied = [None, None]
    for i in range(runs):
        r0 = min(i * m, n - m)                 # the last run overlaps its predecessor instead of being shorter
        out = pipe._stream_out[i & 1]
        if copied[i & 1] is not None:
            main.wait_event(copied[i & 1])     # the buffers of run i - 2 must have left the device
        q = xd[r0:r0 + m]
        (pipe.run if i == 0 else pipe.run_next)(q, xd, out)
        done = torch.cuda.Event()
        done.record(main)
        with torch.cuda.stream(pipe._stream_copy):
            pipe._stream_copy.wait_event(done)
            idx_h[r0:r0 + m].copy_(out[0], non_blocking=True)
            dist_h[r0:r0 + m].copy_(out[1], non_blocking=True)
            copied[i & 1] = torch.cuda.Event()
            copied[i & 1].record(pipe._stream_copy)
    flags_h = _host_copy(pipe.flags)
    main.synchronize()
    pipe._stream_copy.synchronize()
    if int(flags_h[0]):
        raise FloatingPointError("non-finite value in an input profile")
    return idx_h.numpy(), dist_h.numpy()


SMALL_BMU_ROWS = 256     # up to this many vectors go to the one-launch float64 kernel
SMALL_KNN_ROWS = 64      # up to this many query rows (against a small candidate set) likewise


def bmu(nodes, x, device=None, return_dist: bool = False):
    """Flat best-matching-unit index of every row of ``x`` on ``nodes`` [M, D]:
    ``argmin(cdist(nodes, [x_i]))`` (torusMesh.py:127, som.py:351), first minimum.
    A handful of vectors (the per-contig bestMatch call) takes one float64 kernel launch
    (``gmd_bmu_small_f64``); batches go through the staged k = 1 pipeline."""
    xt, x_np = _as_rows(x, "XA")
    nt, _ = _as_rows(nodes, "XB")
    if xt.shape[1] != nt.shape[1]:
        raise ValueError("XA and XB must have the same number of columns "
                         "(i.e. feature dimension.)")
    if 0 < xt.shape[0] <= SMALL_BMU_ROWS and nt.shape[0] > 0:
        dev = xt.device if xt.is_cuda else (nt.device if nt.is_cuda else _device(device))
        with torch.cuda.device(dev):
            xd, nd = _to_dev(xt, dev), _to_dev(nt, dev)
            n = xd.shape[0]
            loc = torch.empty(n, dtype=torch.int32, device=dev)
            dist = torch.empty(n, dtype=torch.float64, device=dev)
            flag = torch.empty(1, dtype=torch.int32, device=dev)
            _lib.check(_lib.lib().gmd_bmu_small_f64(_ptr(nd), nd.shape[0], _ptr(xd), n, xd.shape[1], _ptr(loc),
                                                    _ptr(dist), _ptr(flag), _stream_ptr()))
            if x_np:
                loc_h, dist_h, flag_h = _host_copy(loc), _host_copy(dist), _host_copy(flag)
                torch.cuda.current_stream().synchronize()
                bad, loc, dist = int(flag_h[0]), loc_h.numpy(), dist_h.numpy()
            else:
                bad = int(flag.item())
            if bad:
                raise FloatingPointError("non-finite value in an input profile")
            return (loc, dist) if return_dist else loc
    idx, dist = knn(x, nodes, k=1, metric="euclidean", device=device)
    if return_dist:
        return idx[:, 0], dist[:, 0]
    return idx[:, 0]


def knn_host(q: np.ndarray, c: Optional[np.ndarray], k: int, metric="euclidean", out=None):
    """Same result through the pure C-ABI host entry point (no torch involved):
    host float64 pointers in, host arrays out -- what a ctypes-only caller binds.
    ``out``: (int32 [nq, k], float64 [nq, k]) C-contiguous arrays to fill instead of new ones."""
    lib = _lib.lib()
    qa = np.ascontiguousarray(q, dtype=np.float64)
    ca = qa if c is None else np.ascontiguousarray(c, dtype=np.float64)
    if qa.ndim != 2 or ca.ndim != 2 or qa.shape[1] != ca.shape[1]:
        raise ValueError("XA and XB must be 2-dimensional with the same number of columns")
    if out is None:
        out_i = np.empty((qa.shape[0], k), dtype=np.int32)
        out_d = np.empty((qa.shape[0], k), dtype=np.float64)
    else:
        out_i, out_d = out
        if (out_i.dtype != np.int32 or out_d.dtype != np.float64 or out_i.shape != (qa.shape[0], k)
                or out_d.shape != (qa.shape[0], k) or not out_i.flags.c_contiguous or not out_d.flags.c_contiguous):
            raise ValueError("out: expected C-contiguous int32 and float64 arrays of shape (nq, k)")
    _lib.check(lib.gmd_knn_host_f64(qa.ctypes.data, qa.shape[0], ca.ctypes.data, ca.shape[0], qa.shape[1], int(k),
                                    _metric(metric), out_i.ctypes.data, out_d.ctypes.data))
    return out_i, out_d


def fp32_peak_tflops(iters: int = 4096, repeats: int = 3, form: str = "uniform") -> float:
    """Measured FP32 FMA peak of the current device (TFLOP/s), best of ``repeats``: a
    register-only 8x8 outer-product stream of packed FMAs (FFMA2), ~15 ms per launch.

    ``form="uniform"`` (default, the roofline denominator): one multiplicand in uniform
    registers -- the highest rate the FMA pipe reaches.  ``form="vector"``: three vector-register
    operands, the operand mix of the scan's inner loop (register-file bandwidth caps it at
    ~85 % of the uniform form on B200).  ``form="scalar"``: unpacked FFMA."""
    import os

    lib = _lib.lib()
    dev = _device()
    sink = torch.zeros(4, dtype=torch.float32, device=dev)
    flops = C.c_double(0.0)
    best = 0.0
    old = os.environ.get("GMD_PROBE")
    if form == "uniform":
        os.environ.pop("GMD_PROBE", None)
    else:
        os.environ["GMD_PROBE"] = form
    try:
        for r in range(repeats + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.gmd_fp32_peak_probe(iters, sink.data_ptr(), C.byref(flops), _stream_ptr()))
            e1.record()
            e1.synchronize()
            if r:
                best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    finally:
        if old is None:
            os.environ.pop("GMD_PROBE", None)
        else:
            os.environ["GMD_PROBE"] = old
    return best


# ---------------------------------------------------------------------------
# float64 pairwise stages (small / medium call sites of the reference)

def _to_dev(t: torch.Tensor, dev) -> torch.Tensor:
    return t if t.is_cuda else t.to(dev, non_blocking=True)


def _out(t: torch.Tensor, as_numpy: bool):
    return t.cpu().numpy() if as_numpy else t


def _check_finite(*tensors):
    """np.seterr(all='raise') semantics of the reference (torusMesh.py:57): a NaN / Inf profile is a
    FloatingPointError.  One flag kernel per array and one 4-byte read."""
    lib = _lib.lib()
    flag = torch.empty(1, dtype=torch.int32, device=tensors[0].device)
    _lib.check(lib.gmd_zero_async(_ptr(flag), 4, _stream_ptr()))
    for t in tensors:
        _lib.check(lib.gmd_nonfinite_f64(_ptr(t), t.numel(), _ptr(flag), _stream_ptr()))
    if int(flag.item()):
        raise FloatingPointError("non-finite value in an input profile")


def cdist(xa, xb, metric="euclidean", device=None):
    """scipy.spatial.distance.cdist(XA, XB, metric) -> [mA, mB] float64, bit-identical
    values (torusMesh.py:127; som.py:351; refine.py:927, 945, 1006, 1503)."""
    at, a_np = _as_rows(xa, "XA")
    bt, _ = _as_rows(xb, "XB")
    if at.shape[1] != bt.shape[1]:
        raise ValueError("XA and XB must have the same number of columns "
                         "(i.e. feature dimension.)")
    m = _metric(metric)
    dev = at.device if at.is_cuda else _device(device)
    with torch.cuda.device(dev):
        if at.shape[0] == 0 or bt.shape[0] == 0:
            return _out(torch.empty((at.shape[0], bt.shape[0]), dtype=torch.float64, device=dev), a_np)
        ad, bd = _to_dev(at, dev), _to_dev(bt, dev)
        _check_finite(ad, bd)
        out = torch.empty((ad.shape[0], bd.shape[0]), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().gmd_cdist_f64(_ptr(ad), ad.shape[0], _ptr(bd), bd.shape[0], ad.shape[1], m,
                                            _ptr(out), _stream_ptr()))
        return _out(out, a_np)


def pdist(x, metric="euclidean", device=None):
    """scipy.spatial.distance.pdist(X, metric) -> condensed [n(n-1)/2] float64
    (cluster.py:495, 497; refine.py:1492)."""
    xt, x_np = _as_rows(x, "X")
    m = _metric(metric)
    dev = xt.device if xt.is_cuda else _device(device)
    n = xt.shape[0]
    with torch.cuda.device(dev):
        if n < 2:
            return _out(torch.empty(0, dtype=torch.float64, device=dev), x_np)
        xd = _to_dev(xt, dev)
        _check_finite(xd)
        out = torch.empty(n * (n - 1) // 2, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().gmd_pdist_f64(_ptr(xd), n, xd.shape[1], m, _ptr(out), _stream_ptr()))
        return _out(out, x_np)


def angle_block(xa, xb=None, norm_a=None, norm_b=None, device=None):
    """``arccos(dot(a, b) / (norm_a[i] * norm_b[j]))`` for every pair of rows of two small profile
    sets (refine.py:1545-1546, 1577-1578) -> [mA, mB] float64; pairs the reference drops with a
    FloatingPointError are NaN.  ``norm_a`` / ``norm_b`` are the reference's ``normCoverages`` of
    the rows (it divides by the stored norms, profileManager.py:205); None = the rows' own
    Euclidean norms.  ``xb=None``: pairs within ``xa`` (``norm_b`` then defaults to ``norm_a``)."""
    at, a_np = _as_rows(xa, "XA")
    bt = at if xb is None else _as_rows(xb, "XB")[0]
    if at.shape[1] != bt.shape[1]:
        raise ValueError("XA and XB must have the same number of columns "
                         "(i.e. feature dimension.)")
    dev = at.device if at.is_cuda else _device(device)

    def norms(v, m):
        if v is None:
            return None
        t = torch.as_tensor(np.ascontiguousarray(v, dtype=np.float64) if not isinstance(v, torch.Tensor) else v,
                            dtype=torch.float64).reshape(-1)
        if t.shape[0] != m:
            raise ValueError("one norm per profile row is required")
        return t.contiguous().to(dev)

    with torch.cuda.device(dev):
        if at.shape[0] == 0 or bt.shape[0] == 0:
            return _out(torch.empty((at.shape[0], bt.shape[0]), dtype=torch.float64, device=dev), a_np)
        ad = _to_dev(at, dev)
        bd = ad if xb is None else _to_dev(bt, dev)
        na = norms(norm_a, ad.shape[0])
        nb = na if (xb is None and norm_b is None) else norms(norm_b, bd.shape[0])
        out = torch.empty((ad.shape[0], bd.shape[0]), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().gmd_angle_block_f64(_ptr(ad), ad.shape[0], _ptr(bd), bd.shape[0], ad.shape[1],
                                                  _ptr(na), _ptr(nb), _ptr(out), _stream_ptr()))
        # a non-finite profile or norm gives a NaN pair, which is what the reference's
        # FloatingPointError handler does with it too (refine.py:1548-1549)
        return _out(out, a_np)


def squareform(v):
    """scipy.spatial.distance.squareform for the two directions the reference uses
    (condensed -> square at cluster.py:495; square -> condensed in small2indices math,
    mstore.py:1880-1888).  Pure index movement, done with numpy/torch indexing."""
    is_np = not isinstance(v, torch.Tensor)
    t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)) if is_np else v
    if t.ndim == 1:
        m = t.shape[0]
        n = int(round((1 + (1 + 8 * m) ** 0.5) / 2))
        if n * (n - 1) // 2 != m:
            raise ValueError("Incompatible vector size. It must be a binomial coefficient n choose 2 "
                             "for some integer n >= 2.")
        out = torch.zeros((n, n), dtype=t.dtype, device=t.device)
        iu = torch.triu_indices(n, n, 1, device=t.device)
        out[iu[0], iu[1]] = t
        out[iu[1], iu[0]] = t
    elif t.ndim == 2 and t.shape[0] == t.shape[1]:
        n = t.shape[0]
        iu = torch.triu_indices(n, n, 1, device=t.device)
        out = t[iu[0], iu[1]].contiguous()
    else:
        raise ValueError("The matrix argument must be square.")
    return out.numpy() if is_np else out


def radius_neighbours(x, radius: float, metric="cityblock", return_dist: bool = False, device=None):
    """CSR of ``where(D[i, :] <= radius)[0]`` for every row i (ascending indices) without
    the n x n matrix (cluster.py:533-536, 570-573).  -> (indptr int64 [n+1], indices int32
    [, dists float64])."""
    xt, x_np = _as_rows(x, "X")
    m = _metric(metric)
    dev = xt.device if xt.is_cuda else _device(device)
    n, d = xt.shape
    lib = _lib.lib()
    with torch.cuda.device(dev):
        xd = _to_dev(xt, dev)
        _check_finite(xd)
        if radius != radius:
            raise FloatingPointError("radius is NaN")
        counts = torch.empty(n, dtype=torch.int64, device=dev)
        _lib.check(lib.gmd_radius_count_f64(_ptr(xd), n, d, m, float(radius), _ptr(counts), _stream_ptr()))
        indptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
        _lib.check(lib.gmd_indptr_from_counts_i64(_ptr(counts), n, _ptr(indptr), _stream_ptr()))
        nnz = int(indptr[-1].item())
        indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        dists = torch.empty(max(nnz, 1), dtype=torch.float64, device=dev) if return_dist else None
        _lib.check(lib.gmd_radius_fill_f64(_ptr(xd), n, d, m, float(radius), _ptr(indptr), _ptr(indices),
                                           _ptr(dists), _stream_ptr()))
        res = (_out(indptr, x_np), _out(indices[:nnz], x_np))
        if return_dist:
            res += (_out(dists[:nnz], x_np),)
        return res


def gather_dist(x, indptr, indices, metric="cityblock", device=None):
    """Distances of the CSR pairs (i, indices[p]) in the profile space ``x`` -- the
    ``other_dist_matrix[index][neighbours]`` gathers of cluster.py:547, 583."""
    xt, x_np = _as_rows(x, "X")
    m = _metric(metric)
    dev = xt.device if xt.is_cuda else _device(device)
    with torch.cuda.device(dev):
        xd = _to_dev(xt, dev)
        ip = torch.as_tensor(indptr, dtype=torch.int64).to(dev)
        ix = torch.as_tensor(indices, dtype=torch.int32).to(dev)
        out = torch.empty(max(ix.shape[0], 1), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().gmd_gather_dist_f64(_ptr(xd), xd.shape[0], xd.shape[1], m, _ptr(ip), _ptr(ix),
                                                  _ptr(out), _stream_ptr()))
        return _out(out[:ix.shape[0]], x_np)


def median_dist(a, b=None, metric="cityblock", device=None):
    """``np.median(cdist(a, b, metric))`` (b given; kDistBetweenBins refine.py:1502-1503)
    or ``np.median(pdist(a, metric))`` (b None; kDist refine.py:1490-1496) without
    materialising the block.  Returns a Python float, or None when there are no pairs
    (refine.py:1493-1496)."""
    at, _ = _as_rows(a, "XA")
    m = _metric(metric)
    dev = at.device if at.is_cuda else _device(device)
    lib = _lib.lib()
    with torch.cuda.device(dev):
        ad = _to_dev(at, dev)
        bd = None
        if b is not None:
            bt, _ = _as_rows(b, "XB")
            if bt.shape[1] != at.shape[1]:
                raise ValueError("XA and XB must have the same number of columns "
                                 "(i.e. feature dimension.)")
            bd = _to_dev(bt, dev)
            if ad.shape[0] == 0 or bd.shape[0] == 0:
                return float("nan")  # np.median of an empty block
            _check_finite(ad, bd)
        else:
            if ad.shape[0] < 2:
                return None
            _check_finite(ad)
        ws = torch.empty(lib.gmd_median_workspace_bytes(), dtype=torch.uint8, device=dev)
        out = torch.empty(1, dtype=torch.float64, device=dev)
        _lib.check(lib.gmd_median_dist_f64(_ptr(ad), ad.shape[0], _ptr(bd), 0 if bd is None else bd.shape[0],
                                           ad.shape[1], m, _ptr(out), _ptr(ws), _stream_ptr()))
        return float(out.item())


BATCH_MEDIAN_MAX_PAIRS = 1 << 21   # beyond this a job gets the multi-block select of median_dist


def median_dist_batch(x, jobs, metric="cityblock", device=None):
    """Medians of many never-materialised distance blocks of one profile array in a single
    launch.  ``jobs`` is a sequence of ``(rows_a, rows_b)`` row-index arrays into ``x``;
    ``rows_b=None`` asks for ``median(pdist(x[rows_a]))`` (kDist, refine.py:1490-1496) and
    otherwise ``median(cdist(x[rows_a], x[rows_b]))`` (kDistBetweenBins, refine.py:1502-1503).
    -> list of floats; ``None`` where a pdist job has fewer than two rows (the reference's
    ``return None``), nan where a cdist job is empty."""
    xt, _ = _as_rows(x, "X")
    m = _metric(metric)
    dev = xt.device if xt.is_cuda else _device(device)
    lib = _lib.lib()
    n_rows = xt.shape[0]
    rows, quads, big = [], [], {}
    off = 0
    for j, (ra, rb) in enumerate(jobs):
        ra = np.ascontiguousarray(ra, dtype=np.int32).ravel()
        rbv = None if rb is None else np.ascontiguousarray(rb, dtype=np.int32).ravel()
        for r in (ra,) if rbv is None else (ra, rbv):
            if len(r) and (r.min() < 0 or r.max() >= n_rows):
                raise IndexError("row index out of range")
        pairs = len(ra) * (len(ra) - 1) // 2 if rbv is None else len(ra) * len(rbv)
        if pairs > BATCH_MEDIAN_MAX_PAIRS:
            big[j] = (ra, rbv)
            quads.append((0, 0, 0, 0))
            continue
        a_off = off
        rows.append(ra)
        off += len(ra)
        if rbv is None:
            quads.append((a_off, len(ra), 0, -1))
        else:
            quads.append((a_off, len(ra), off, len(rbv)))
            rows.append(rbv)
            off += len(rbv)
    with torch.cuda.device(dev):
        xd = _to_dev(xt, dev)
        _check_finite(xd)
        out = torch.empty(max(len(quads), 1), dtype=torch.float64, device=dev)
        if quads:
            rows_d = torch.from_numpy(np.concatenate(rows) if rows else np.zeros(1, np.int32)).to(dev)
            if rows_d.numel() == 0:
                rows_d = torch.zeros(1, dtype=torch.int32, device=dev)
            jobs_d = torch.from_numpy(np.asarray(quads, dtype=np.int32).reshape(-1)).to(dev)
            _lib.check(lib.gmd_median_batch_f64(_ptr(xd), n_rows, xd.shape[1], m, _ptr(rows_d), _ptr(jobs_d),
                                                len(quads), _ptr(out), _stream_ptr()))
        vals = out.cpu().numpy()[:len(quads)].tolist()
    res = []
    for j, (q, v) in enumerate(zip(quads, vals)):
        if j in big:
            ra, rbv = big[j]
            res.append(median_dist(xt[torch.from_numpy(ra).long()],
                                   None if rbv is None else xt[torch.from_numpy(rbv).long()], metric, device=dev))
        elif q[3] < 0 and q[1] < 2:
            res.append(None)
        else:
            res.append(float(v))
    return res


def kth_neighbour_dist(x, k: int, metric="cityblock", device=None):
    """``np.sort(squareform(pdist(X, metric)))[:, k]`` -- the k-th smallest entry of every
    row of the all-pairs matrix, self (0.0) at column 0 (cluster.py:496, 498)."""
    idx, dist = knn(x, None, k=k + 1, metric=metric, device=device)
    return dist[:, k]


# ---------------------------------------------------------------------------
# SOM training with the mesh resident on the device

def som_stamp(radius_decaying: float, influence_rate: float) -> np.ndarray:
    """The neighbourhood "stamp" of one SOM.train iteration (som.py:282-318), vectorised:
    linear decay inside the rounded radius, axis entries halved, four quadrants overlaid,
    centre = influence rate, all-zero borders culled.  Host-side and tiny."""
    grad = -1 * influence_rate / radius_decaying
    mr = int(radius_decaying)
    r = np.arange(mr + 1)
    dist = np.sqrt((r[:, None] ** 2 + r[None, :] ** 2).astype(np.float64))
    q = np.where(np.round(dist + 0.00001) <= radius_decaying, dist * grad + influence_rate, 0.0)
    q[:, 0] /= 2
    q[0, :] /= 2
    stamp = np.zeros((2 * mr + 1, 2 * mr + 1))
    stamp[mr:, mr:] += q
    stamp[:mr + 1, mr:] += np.rot90(q, 1)
    stamp[:mr + 1, :mr + 1] += np.rot90(q, 2)
    stamp[mr:, :mr + 1] += np.rot90(q, 3)
    stamp[mr, mr] = influence_rate
    nz = np.flatnonzero(np.max(stamp, axis=0) > 0.0)
    if len(nz):
        k = int(nz[0])
        stamp = stamp[k:2 * mr + 1 - k, k:2 * mr + 1 - k]
    return np.ascontiguousarray(stamp)


def som_train(nodes, train_vectors, weights=None, iterations: int = 1000, vector_subset: int = 1000,
              influence_rate: float = 0.4, mask=None, radius: float = 0.0, device=None) -> np.ndarray:
    """SOM.train (som.py:214-392) with weights, training vectors and deltas on the device.

    Per iteration: one BMU search of the iteration's vectors on the frozen weights (the
    k = 1 scan + float64 re-rank, replacing the ``argmin(cdist(flat_nodes, [v]))`` loop of
    som.py:348-353) and one gather-update kernel (``gmd_som_update_f64``).  The host keeps
    what is not distance work: the stamp of the iteration and the subset shuffle, which draws
    from numpy's GLOBAL generator exactly as the reference does (som.py:324-326), so a caller
    that seeds ``np.random`` gets the reference's subsets.  All of the reference's semantics
    are kept (see ``oracle.som_train`` for the list); the result is bit-identical.
    ``nodes`` is the mesh [rows, cols, D] used when ``weights`` is None (``self.weights.nodes``);
    ``mask`` a bool/uint8 [rows, cols] array of the nodes allowed to move.  Returns new weights."""
    lib = _lib.lib()
    replace = weights is None
    w_host = np.array(nodes if replace else weights, dtype=np.float64, copy=True)
    if w_host.ndim != 3:
        raise ValueError("weights must be [rows, cols, dimension]")
    rows, cols, dim = w_host.shape
    tv = np.ascontiguousarray(train_vectors, dtype=np.float64)
    if tv.ndim != 2 or tv.shape[1] != dim:
        raise ValueError("training vectors must be [n, dimension] with the mesh's dimension")
    if dim < 4:
        # the reference scales weights_patch[:, :, 0..3] by name (som.py:360-363)
        raise IndexError(f"index 3 is out of bounds for axis 2 with size {dim}")
    if not (np.isfinite(w_host).all() and np.isfinite(tv).all()):
        raise FloatingPointError("non-finite value in the SOM weights or training vectors")
    if radius == 0.0:
        radius = float(rows) / 2
    time_constant = iterations / np.log(radius)
    n = len(tv)
    if n <= vector_subset:
        index_array = np.arange(n)
        cut_off = n
    else:
        rand_index_array = np.arange(n)
        cut_off = vector_subset
    dev = _device(device)
    with torch.cuda.device(dev):
        w = torch.from_numpy(w_host).to(dev)                 # [rows, cols, D] float64, updated in place
        flat_search = w.view(rows * cols, dim)               # what the BMU search reads
        if not replace and mask is None:
            flat_search = w.clone().view(rows * cols, dim)   # reference quirk: never refreshed (som.py:380-384)
        x = torch.from_numpy(tv).to(dev)
        mask_d = None
        mask_any = True
        if mask is not None:
            m = np.ascontiguousarray(np.asarray(mask) != 0, dtype=np.uint8)
            if m.shape != (rows, cols):
                raise ValueError("mask must be [rows, cols]")
            mask_any = bool(m.any())
            mask_d = torch.from_numpy(m).to(dev)
        loc = torch.empty(max(cut_off, 1), dtype=torch.int32, device=dev)
        flag = torch.empty(1, dtype=torch.int32, device=dev)
        small = 0 < cut_off <= 65535                         # one float64 kernel per iteration (gmd_bmu_small_f64)
        pipe = None
        if cut_off > 0 and not small:
            pipe = KnnPipeline(cut_off, rows * cols, dim, 1, "euclidean", same=False, device=dev)
        bad = False
        for i in range(1, iterations + 1):
            radius_decaying = radius * np.exp(-1.0 * i / time_constant)
            if radius_decaying < 2:
                break
            stamp = som_stamp(radius_decaying, influence_rate)
            sr = int((len(stamp) - 1) / 2)
            if n > vector_subset:
                np.random.shuffle(rand_index_array)
                index_array = rand_index_array[:cut_off]
            if cut_off == 0:
                continue
            idx_d = torch.from_numpy(np.ascontiguousarray(index_array, dtype=np.int32)).to(dev)
            q = x.index_select(0, idx_d.long())
            if small:
                # [T] int32 BMU on the frozen weights, scipy's arithmetic, first minimum
                _lib.check(lib.gmd_bmu_small_f64(_ptr(flat_search), rows * cols, _ptr(q), cut_off, dim, _ptr(loc), 0,
                                                 _ptr(flag), _stream_ptr()))
                if i == 1:
                    bad = bool(flag.item())                  # inputs were checked on the host; once is enough
            else:
                loc, _ = pipe.run(q, flat_search)            # [T, 1] int32 BMU on the frozen weights
            stamp_d = torch.from_numpy(stamp).to(dev)
            _lib.check(lib.gmd_som_update_f64(_ptr(w), _ptr(w), rows, cols, dim, _ptr(x), n, _ptr(idx_d), _ptr(loc),
                                              cut_off, _ptr(stamp_d), sr, _ptr(mask_d), _stream_ptr()))
            if mask is not None and not mask_any:
                pass                                          # nothing moved, flat_nodes not refreshed either
        out = w.cpu().numpy()
        if bad or (pipe is not None and pipe.saw_nonfinite()):
            raise FloatingPointError("non-finite value reached the BMU search")
    return out
