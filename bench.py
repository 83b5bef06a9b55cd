#!/usr/bin/env python
"""Benchmark of the contig-profile distance / neighbour-search hot path.

    python bench.py --gpus N --steps K --warmup W            # this engine (B200)
    python bench.py --impl reference --gpus N ...            # the reference's scipy CPU path

Workload (BASELINE.json: "pair-distances/s (fused top-k); 1M-contig kNN wall time at
1/2/4/8 B200"): the exact k=32 Euclidean neighbour graph of 1 000 000 synthetic contigs x
30 samples (configs[4] at 1 M; it fits one GPU).  One step = one complete pass of the path
over the batch: [all-gather of the row shards, N>1] -> centring -> grouped staging (leaders,
nearest-leader keys, padded sort, tiles + balls) -> pruned tcgen05 distance + top-k scan ->
float64 re-rank + certificate -> exact rows.  With N GPUs the 1 M query rows are block-sharded
(strong scaling: total work fixed) and every rank runs the SAME algorithm.

Prints ONE JSON line on rank 0.  `value` = pair-distances/s with the float64 profiles
already resident in HBM; `e2e` = the same through the public plugin call with HOST buffers
(`engine.knn(numpy)` at N=1, `dist.sharded_knn` between pinned host buffers at N>1): H2D of
the profiles and D2H of the neighbour indices + distances inside the timed region.
"""
from __future__ import annotations

import argparse
import faulthandler
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "pair-distances/s (fused top-k)"
UNIT = "pairs/s"
WORKLOADS = {
    # name: (N contigs, D, k, metric, seed)
    "cfg5_1m_x30_k32": (1_000_000, 30, 32, "euclidean", 5),
    "cfg2_100k_x10_k16": (100_000, 10, 16, "euclidean", 2),
    "cfg5_4m_x30_k32": (4_000_000, 30, 32, "euclidean", 5),
    "tiny": (20_000, 30, 32, "euclidean", 5),
}
TILE_FLOP = 2.0 * 128 * 128 * 32 * 3          # executed by one (query tile, candidate tile, K block): 3 x FP16 passes


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="cfg5_1m_x30_k32", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--prune", default="auto", choices=["auto", "on", "off"],
                    help="pruned tensor-core scan (engine default for large Euclidean problems); off: the "
                         "locality-ordered full sweep of round 1")
    return ap.parse_args()


# ---------------------------------------------------------------------------- reference arm

def _ref_chunk(args):
    """scipy cdist + argpartition + sort on a slice of query rows (one host process)."""
    lo, hi, k, metric = args
    import oracle

    oracle.knn(_REF_X[_REF_ROWS[lo:hi]], _REF_X, k, metric, chunk=8)
    return hi - lo


_REF_X = None
_REF_ROWS = None
_REAL_STDOUT = None


def emit(line: dict):
    """The one JSON line, on the process's original stdout."""
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, text.encode())


class Watchdog:
    """Every leg of the single-GPU bench gets a generous deadline.  If one does not finish (a hung kernel, a copy
    that never returns), the thread stacks go to stderr and the JSON line is printed with what was measured BEFORE
    that leg -- the headline number is the first thing measured -- plus a "watchdog" key naming the leg; then the
    process leaves without tearing CUDA down.  A run that finishes normally never sees it."""

    def __init__(self):
        self.line = None
        self.name = "start-up"
        self.deadline = time.monotonic() + 600.0
        self.done = False

    def start(self):
        threading.Thread(target=self._run, daemon=True).start()

    def leg(self, name: str, seconds: float):
        self.name = name
        self.deadline = time.monotonic() + seconds

    def _run(self):
        while not self.done:
            time.sleep(0.5)
            if self.done or time.monotonic() <= self.deadline:
                continue
            try:
                faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
            except Exception:                                    # noqa: BLE001 -- nothing may stop the line
                pass
            msg = f"leg '{self.name}' did not finish in time; the line holds what was measured before it"
            sys.stderr.write("bench.py watchdog: " + msg + "\n")
            sys.stderr.flush()
            line = self.line
            if line is not None:
                line = dict(line)
                line["watchdog"] = msg
                emit(line)
                os._exit(0)
            os._exit(3)


def base_config(workload):
    n, d, k, metric, _ = WORKLOADS[workload]
    return {"workload": workload, "contigs": n, "dim": d, "k": k, "metric": metric,
            "l2": "inputs exceed L2 (float64 profiles %d MB against 126 MB)" % (n * d * 8 // 2**20)}


def reference_arm(args):
    """The reference's own CPU implementation of the path: the literal scipy expressions
    (oracle.knn = chunked cdist -> k smallest per row, float64) on every host core, each
    step a bounded sample of the workload's query rows against ALL candidates."""
    global _REF_X, _REF_ROWS
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp

    from groopm_b200 import synth

    n, d, k, metric, seed = WORKLOADS[args.workload]
    cores = len(os.sched_getaffinity(0))
    rows_per_core = max(8, int(1.5e8 // n))          # ~5 s of cdist per core and step
    sample = min(n, cores * rows_per_core)
    _REF_X = synth.cfg_knn_rows(n, d, seed)
    _REF_ROWS = np.sort(np.random.default_rng(seed).choice(n, sample, replace=False))
    bounds = np.linspace(0, sample, cores + 1).astype(int)
    jobs = [(int(a), int(b), k, metric) for a, b in zip(bounds, bounds[1:]) if b > a]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_chunk, jobs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_chunk, jobs)
        dt = time.perf_counter() - t0
    value = sample * n * args.steps / dt
    cfg = base_config(args.workload)
    cfg["path"] = "scipy cdist (float64) + argpartition/sort per query row, all candidates"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} of {n} query rows x all {n} candidates per step, "
                                   f"{cores} processes"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------- secondary workloads

def _peaks():
    try:
        mp = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        return (float(mp["bf16_tflops_sustained"]), float(mp["bf16_tflops"]), float(mp["hbm_gbs"]),
                "MEASURED_PEAKS.json (bf16_tflops_sustained; the FP16 tensor rate is the bf16 rate)")
    except (OSError, ValueError, KeyError):
        return 1590.0, 1590.0, 6500.0, "MEASURED_PEAKS.json absent: fallback 1590 TFLOP/s / 6500 GB/s (B200_PROFILING.md)"


def _time_pipeline(torch, pipe, qd, cd, steps):
    """-> (ms per step, scan ms of the last step) with everything resident."""
    for _ in range(2):
        pipe.run(qd, cd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        pipe.reset(); pipe.col_mean(cd); pipe.prep_candidates(cd); pipe.prep_queries(qd)
        if i == steps - 1:
            s0.record()
        pipe.scan()
        if i == steps - 1:
            s1.record()
        pipe.refine(qd, cd); pipe.exact_rows(qd, cd)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, s0.elapsed_time(s1)


def _scan_record(pipe, name, nq, nc, d, ms, scan_ms, tc_peak, fp32_peak, metric="euclidean"):
    pairs = float(nq) * float(nc)
    alg_tf = 2.0 * d * pairs / (scan_ms * 1e-3) / 1e12
    rec = {"workload": name, "pairs_per_s": pairs / (ms * 1e-3), "ms_per_step": ms, "scan_ms": scan_ms,
           "scan_kernel": ("pruned tensor-core 3xFP16" if pipe.prune else "tensor-core 3xFP16" +
                           (" (ordered full sweep)" if pipe.order else "")) if pipe.form == 3
           else ("k=1 BMU scan" if pipe.bmu else "FP32 scan"),
           "scan_tflops_algorithmic": alg_tf, "exact_rows": pipe.exact_row_count()}
    if pipe.form == 3:
        nkb = (d + 1 + 31) // 32
        if pipe.prune:
            st = pipe.prune_stats()
            tiles = st["tiles_visited"]
            rec["tiles_visited_fraction"] = st["visited_fraction"]
        else:
            tiles = -(-nq // 128) * -(-nc // 128)
        ex_tf = tiles * nkb * TILE_FLOP / (scan_ms * 1e-3) / 1e12
        rec["roofline"] = {"bound": "tensor", "achieved": alg_tf, "peak": tc_peak, "unit": "TFLOP/s",
                           "frac": alg_tf / tc_peak, "executed_tflops": ex_tf, "frac_executed": ex_tf / tc_peak}
    else:
        pk = fp32_peak if metric == "euclidean" else fp32_peak / 2       # L1 has no FMA form: FADD rate
        rec["roofline"] = {"bound": "fp32", "achieved": alg_tf, "peak": pk, "unit": "TFLOP/s", "frac": alg_tf / pk}
    return rec


def secondary_workloads(engine, synth, torch, tc_peak, fp32_peak, out=None, tick=None):
    """The other single-GPU configurations of BASELINE.json, device-resident, each with its own roofline
    object: configs[4] without pruning (what the tensor pipeline does when nothing can be skipped),
    configs[1] 100 k x 10 k=16 both metrics, configs[2] BMU of 1 M contigs on a 64 x 64 mesh, configs[3]
    500 k x 136 raw signatures (k = 16 graph + bin-vs-bin medians), and the latency of the small calls."""
    from groopm_b200 import _lib, dropin
    out = [] if out is None else out          # (the caller's list: finished records are visible at once)
    tick = tick or (lambda name, seconds: None)

    def run(name, q, c, k, metric, steps=10, **kw):
        tick(name, 300.0)
        same = c is None
        qd = torch.from_numpy(q).cuda()
        cd = qd if same else torch.from_numpy(c).cuda()
        pipe = engine.KnnPipeline(qd.shape[0], cd.shape[0], qd.shape[1], k, metric, same=same, **kw)
        ms, scan_ms = _time_pipeline(torch, pipe, qd, cd, steps)
        out.append(_scan_record(pipe, name, qd.shape[0], cd.shape[0], qd.shape[1], ms, scan_ms, tc_peak, fp32_peak, metric))
        del pipe
        return qd

    n, d, k, metric, seed = WORKLOADS["cfg5_1m_x30_k32"]
    x = synth.cfg_knn_rows(n, d, seed)
    run("cfg5_1m_x30_k32_full_sweep (no pruning: ordered sweep of every tile)", x, None, k, metric, steps=3,
        form=_lib.FORM_TC, order=True, prune=False)
    del x
    x = synth.cfg_knn_rows(100_000, 10, 2)
    run("cfg2_100k_x10_k16_euclidean", x, None, 16, "euclidean")
    run("cfg2_100k_x10_k16_cityblock", x, None, 16, "cityblock")
    block, nodes = synth.cfg_bmu(1_000_000, 64)
    run("cfg3_bmu_1m_x4_on_64x64", block, nodes.reshape(-1, 4), 1, "euclidean")
    # configs[3]: 500 k raw 136-d signatures: the full 2.5e11-pair k = 16 graph, then the kDistBetweenBins
    # medians of 250 bins x 2000 rows against their nearest bins (refine.py:1502-1503; 48 pairs timed)
    sigs = synth.quantise(synth.kmer_sigs(500_000, 4), 34)
    sd = run("cfg4_500k_x136_k16_euclidean", sigs, None, 16, "euclidean", steps=2)
    torch.cuda.synchronize()
    tick("cfg4 bin-vs-bin medians", 180.0)
    pairs_idx = [(np.arange(a * 2000, (a + 1) * 2000), np.arange(b * 2000, (b + 1) * 2000))
                 for a in range(8) for b in ((a + 1) % 250, (a + 7) % 250, (a + 31) % 250, (a + 64) % 250,
                                             (a + 100) % 250, (a + 200) % 250)]
    t0 = time.perf_counter()
    for ra, rb in pairs_idx:
        engine.median_dist(sd[torch.from_numpy(ra).cuda()], sd[torch.from_numpy(rb).cuda()], "cityblock")
    dt = time.perf_counter() - t0
    out.append({"workload": "cfg4_bin_vs_bin_medians_2000x2000_d136_cityblock", "medians": len(pairs_idx),
                "ms_per_median": dt / len(pairs_idx) * 1e3, "pairs_per_s": len(pairs_idx) * 4e6 / dt,
                "note": "median(cdist(bin_i, bin_j, 'cityblock')) without the block (radix select, 7 passes)"})
    del sd, sigs

    # small-call latency against the scipy calls they replace (same host, same arrays)
    tick("small-call latencies", 180.0)
    from scipy.spatial import KDTree as SpKDTree
    from scipy.spatial.distance import cdist as sp_cdist

    flat = nodes.reshape(-1, 4)
    vs = block[:2000]
    bm = dropin.BestMatcher(flat, 64, 64)
    for v in vs[:50]:
        bm(v)
    t0 = time.perf_counter()
    for v in vs:
        bm(v)
    t_gpu = (time.perf_counter() - t0) / len(vs)
    t0 = time.perf_counter()
    for v in vs:
        int(np.argmin(sp_cdist(flat, [v])))
    t_cpu = (time.perf_counter() - t0) / len(vs)
    t0 = time.perf_counter()
    bm.many(block[:200_000])
    t_many = (time.perf_counter() - t0) / 200_000
    lat = {"workload": "small_call_latency_us",
           "best_match_one_vector_4096_nodes": {"engine": t_gpu * 1e6, "scipy": t_cpu * 1e6,
                                                "engine_batched_per_vector": t_many * 1e6}}
    cent = np.random.default_rng(7).random((2000, 3))
    tree_g, tree_c = dropin.KDTree(cent), SpKDTree(cent)
    tree_g.query(cent[0], k=6)
    t0 = time.perf_counter()
    for r in cent[:500]:
        tree_g.query(r, k=6)
    tg = (time.perf_counter() - t0) / 500
    t0 = time.perf_counter()
    for r in cent[:500]:
        tree_c.query(r, k=6)
    tc_ = (time.perf_counter() - t0) / 500
    t0 = time.perf_counter()
    tree_g.query(cent, k=6)
    tb = (time.perf_counter() - t0) / len(cent)
    lat["kdtree_query_k6_2000_bins"] = {"engine_per_call": tg * 1e6, "scipy_per_call": tc_ * 1e6,
                                        "engine_all_bins_one_call_per_bin": tb * 1e6}
    a, b = x[:400], x[400:900]
    dropin.k_dist_between(a, b)
    t0 = time.perf_counter()
    for _ in range(50):
        dropin.k_dist_between(a, b)
    tg = (time.perf_counter() - t0) / 50
    t0 = time.perf_counter()
    for _ in range(50):
        float(np.median(sp_cdist(a, b, "cityblock")))
    tc_ = (time.perf_counter() - t0) / 50
    lat["k_dist_between_400x500_d10"] = {"engine": tg * 1e6, "scipy": tc_ * 1e6}
    out.append(lat)
    return out


def cpu_baseline_leg(x, k, metric, seconds_target=12.0):
    """Single-core oracle on a bounded sample of the same workload (rank 0, N=1 only), and the
    reference-literal figure BASELINE.md section 3 asks for: the per-contig ``argmin(cdist(flatNodes, [v]))``
    Python loop of TorusMesh.bestMatch (torusMesh.py:127) on a 20 k sample of the cfg3 contigs."""
    import oracle
    from groopm_b200 import synth

    n = x.shape[0]
    rows = max(8, int(seconds_target * 1.2e7 // n))
    sel = np.sort(np.random.default_rng(0).choice(n, min(rows, n), replace=False))
    t0 = time.perf_counter()
    oracle.knn(x[sel], x, k, metric, chunk=8)
    dt = time.perf_counter() - t0
    block, nodes = synth.cfg_bmu(20_000, 64)
    flat = nodes.reshape(-1, 4)
    t0 = time.perf_counter()
    for v in block:
        oracle.best_match(flat, v, 64, 64)
    lit = (time.perf_counter() - t0) / len(block)
    return {"value": len(sel) * n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{len(sel)} of {n} query rows x all {n} candidates, scipy cdist + top-k, 1 thread, {dt:.1f} s",
            "reference_literal": {"call": "TorusMesh.bestMatch: argmin(cdist(flatNodes, [v])) per contig (torusMesh.py:127)",
                                  "us_per_contig": lit * 1e6, "pairs_per_s": 4096 / lit,
                                  "sample": "20000 contigs x 4096 nodes x D=4, 1 thread",
                                  "extrapolated_s_per_1m_contigs": lit * 1e6}}


# ---------------------------------------------------------------------------- clocks

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------- engine arm

def engine_arm(args):
    import torch
    import torch.distributed as tdist

    from groopm_b200 import _lib
    from groopm_b200 import dist as gdist
    from groopm_b200 import engine, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its own lines (e.g. "NCCL version ...") on file descriptor 1: point that at
        # stderr for the whole run and keep the real stdout for the ONE JSON line
        sys.stdout.flush()
        global _REAL_STDOUT
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
        tdist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.SUM)
        return float(t.item())

    prune = {"auto": None, "on": True, "off": False}[args.prune]
    tc_peak, tc_burst, hbm_peak, peak_src = _peaks()

    def measure(workload, steps, warmup, sample_clocks, progress=None):
        """The step of `workload` on this rank's shard: device-resident and end-to-end timings.
        progress(next_leg, seconds, partial_result): called after every leg (the watchdog's view of the run)."""
        progress = progress or (lambda name, seconds, res: None)
        n, d, k, metric, seed = WORKLOADS[workload]
        x = synth.cfg_knn_rows(n, d, seed)                   # every rank draws the same array
        r0, r1 = gdist.shard_bounds(n, world, rank)
        n_local = r1 - r0
        x_host = torch.from_numpy(x[r0:r1].copy()).pin_memory()   # this rank's shard, pinned host
        x_dev = x_host.to(dev)
        kw = dict(prune=prune)
        if prune is False:
            kw.update(form=_lib.FORM_TC, order=(world == 1))
        pipe = engine.KnnPipeline(max(n_local, 1), n, d, k, metric, same=(world == 1), device=dev, **kw)
        out_idx_host = torch.empty((n_local, k), dtype=torch.int32).pin_memory()
        out_dist_host = torch.empty((n_local, k), dtype=torch.float64).pin_memory()
        scan_ev = pipe.scan_events = []

        def step(src_dev):
            """One pass of the hot path for this rank's rows; returns device (idx, dist)."""
            if world == 1:
                return pipe.run(src_dev, src_dev)
            # all-gather of the rows, leaders, nearest-leader search of this rank's rows + all-gather of the keys,
            # the rest of the staging, scan, re-rank: dist.sharded_knn
            return gdist.sharded_knn(src_dev, n, k, metric, pipeline=pipe)

        def e2e_step():
            if world == 1:
                # the plugin call: host arrays in, host arrays out (engine.knn keeps the pipeline of a shape and
                # hands back pinned arrays)
                return engine.knn(x_host.numpy(), None, k=k, metric=metric, prune=prune,
                                  **({"form": _lib.FORM_TC, "order": True} if prune is False else {}))
            xd = x_host.to(dev, non_blocking=True)           # H2D of this rank's profiles
            idx, dist_ = step(xd)
            out_idx_host.copy_(idx[:n_local], non_blocking=True)   # D2H of this rank's result
            out_dist_host.copy_(dist_[:n_local], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out_idx_host, out_dist_host

        # ---- (A) device-resident throughput
        for _ in range(warmup):
            step(x_dev)
        barrier()
        scan_ev.clear()
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record()
        for _ in range(steps):
            step(x_dev)
        t_end.record()
        barrier()
        dev_ms = max_over_ranks(t_start.elapsed_time(t_end))
        scan_ms = sum(a.elapsed_time(b) for a, b in scan_ev) / len(scan_ev)
        scan_ms_max = max_over_ranks(scan_ms)
        exact_rows = pipe.exact_row_count()
        stats = pipe.prune_stats() if pipe.prune else None
        res_idx, res_dist = pipe.out_idx[:n_local].clone(), pipe.out_dist[:n_local].clone()
        res = dict(n=n, d=d, k=k, metric=metric, n_local=n_local, pipe=pipe, dev_ms=dev_ms, scan_ms=scan_ms,
                   scan_ms_max=scan_ms_max, exact_rows=exact_rows, stats=stats, e2e_s=None, clocks=None,
                   cabi_ms=None, verified=0, ok=None, x=x, h2d=int(x_host.numel() * 8 * world),
                   d2h=int((out_idx_host.numel() * 4 + out_dist_host.numel() * 8) * world))
        progress("end to end through engine.knn(numpy)", 300.0, res)

        # ---- (B) end to end through the public API with host buffers
        for _ in range(warmup):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        clocks = sampler.stop() if sampler else None
        res.update(e2e_s=e2e_s, clocks=clocks)
        progress("C-ABI host entry point gmd_knn_host_f64", 240.0, res)

        # ---- (C) the C-ABI host entry point (gmd_knn_host_f64: library-owned buffers, pageable copies), N = 1
        cabi_ms = None
        if world == 1 and rank == 0 and prune is not False:
            # (ordinary pageable numpy arrays, the outputs allocated and touched once: a caller's own buffers)
            outs = (np.zeros((n, k), dtype=np.int32), np.zeros((n, k), dtype=np.float64))
            engine.knn_host(x, None, k, metric, out=outs)
            t0 = time.perf_counter()
            for _ in range(3):
                engine.knn_host(x, None, k, metric, out=outs)
            cabi_ms = (time.perf_counter() - t0) / 3 * 1e3

        res.update(cabi_ms=cabi_ms)
        progress("verification against the oracle", 240.0, res)
        # ---- (D) verification outside the timed regions: sampled rows of THIS rank's result against the
        # reference's scipy arithmetic (oracle.knn as the checker)
        verified = 0
        ok = True
        if not args.no_verify and n_local > 0:
            import oracle

            rows = np.sort(np.random.default_rng(1000 + rank).choice(n_local, min(64, n_local), replace=False))
            od, oi = oracle.knn(x[r0 + rows], x, k, metric, chunk=16)
            gi, gd = res_idx[rows].cpu().numpy(), res_dist[rows].cpu().numpy()
            ok = bool(np.array_equal(gi, oi) and np.array_equal(gd, od))
            verified = len(rows)
        verified = int(sum_over_ranks(float(verified)))
        all_ok = sum_over_ranks(0.0 if ok else 1.0) == 0.0
        res.update(verified=verified, ok=all_ok)
        progress("secondary workloads", 300.0, res)
        return res

    wd = Watchdog()
    if world == 1:
        wd.start()
    wd.leg("FP32 peak probe", 180.0)
    peak_probe = engine.fp32_peak_tflops()                   # measured FFMA2 peak of this GPU (uniform operand)

    def build_line(m):
        """The JSON line from a (possibly partial) result of measure(): legs not yet run read as null."""
        n, d, k, metric, pipe = m["n"], m["d"], m["k"], m["metric"], m["pipe"]
        pairs = float(n) * float(n)
        value = pairs * args.steps / (m["dev_ms"] * 1e-3)
        e2e_value = pairs * args.steps / m["e2e_s"] if m["e2e_s"] else None
        # roofline of the dominant kernel: 2*D flop per pair, one launch covers this rank's
        # n_local x n pairs
        launch_flop = 2.0 * d * float(m["n_local"]) * float(n)
        achieved = launch_flop / (m["scan_ms"] * 1e-3) / 1e12
        tc = pipe.form == 3
        traffic = None
        prof = ROOT / "profiles" / "scan_traffic.json"
        variant = "pruned" if pipe.prune else ("ordered" if pipe.order else "plain")
        if prof.exists() and world == 1:                  # the capture is of the single-GPU launch (all query tiles)
            try:
                traffic = json.loads(prof.read_text()).get(f"{args.workload}:{variant}")
            except (ValueError, OSError):
                traffic = None
        if tc:
            nkb = (d + 1 + 31) // 32
            tiles = m["stats"]["tiles_visited"] if m["stats"] else (-(-m["n_local"] // 128) * -(-n // 128))
            executed = tiles * nkb * TILE_FLOP / (m["scan_ms"] * 1e-3) / 1e12
            roofline = {
                "bound": "tensor",
                "kernel": "gmd_tc::scan_tc_prune_kernel" if pipe.prune else "gmd_tc::scan_tc_kernel",
                "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s", "frac": achieved / tc_peak,
                "peak_source": peak_src, "peak_burst": tc_burst, "frac_of_burst": achieved / tc_burst,
                "note": "achieved = ALGORITHMIC flops (2*D per pair of the whole Nq x Nc problem) / scan time; the "
                        "pruned scan loads only the candidate tiles whose ball bound its running thresholds cannot "
                        "rule out, so it runs above the dense tensor roofline while the result stays exact; "
                        "executed_tflops counts the tiles it did multiply (K padded to 32, 3 x FP16 hi/lo passes)",
                "executed_tflops": executed, "frac_executed": executed / tc_peak,
                "tiles_visited_fraction": m["stats"]["visited_fraction"] if m["stats"] else 1.0,
                "fp32_pipe_peak": peak_probe, "vs_fp32_pipe_peak": achieved / peak_probe,
                "scan_ms_per_launch": m["scan_ms_max"], "flop_per_launch": launch_flop, "traffic": traffic,
                "traffic_source": f"profiles/scan_traffic.json [{args.workload}:{variant}] (ncu --set full of this "
                                  "variant)" if traffic is not None else None}
        else:
            roofline = {"bound": "fp32", "kernel": "gmd_scan::scan_topk_kernel", "achieved": achieved,
                        "peak": peak_probe, "unit": "TFLOP/s", "frac": achieved / peak_probe,
                        "peak_source": "FFMA2 probe kernel (uniform-register multiplicand) measured on this GPU "
                                       "in this run; MEASURED_PEAKS.json has no FP32 entry",
                        "scan_ms_per_launch": m["scan_ms_max"], "flop_per_launch": launch_flop, "traffic": traffic}
        cfg = base_config(args.workload)
        cfg.update({
            "parallelism": (f"query rows block-sharded over {world} GPUs; one NCCL all-gather of the float64 rows per "
                            f"step; the same pruned scan on every rank" if world > 1 else "single GPU"),
            "arithmetic": ("tcgen05 3xFP16 hi/lo split (FP32-class) " if tc else "FP32 ") +
                          "fused distance+top-k scan proposes k+8 candidates; float64 re-rank and certificate make "
                          "the result exact" +
                          ("; profiles staged in the grouped order (leader clustering of a sample, nearest-leader "
                           "keys, groups padded to whole tiles, all inside the timed step) and candidate tiles "
                           "skipped by ball bounds" if pipe.prune else ""),
            "pruned_scan": bool(pipe.prune), "ordered_sweep": bool(pipe.order), "form": pipe.form, "kk": pipe.kk,
            "exact_rows_last_step": m["exact_rows"], "verified_rows": m["verified"], "verified": m["ok"],
            "verification": "sampled rows of every rank's result == oracle.knn (scipy float64), bit for bit, outside "
                            "the timed regions",
        })
        return {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["dev_ms"] / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16x3" if tc else "f32", "data": "synthetic",
            "config": cfg,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "ms_per_step": m["e2e_s"] / args.steps * 1e3 if m["e2e_s"] else None,
                    "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                    "call": "engine.knn(numpy float64 rows) -> numpy (indices, distances)" if world == 1 else
                            "dist.sharded_knn between pinned host buffers on every rank",
                    "c_abi_gmd_knn_host_f64_ms": m["cabi_ms"]},
            "gpu_launches": pipe.launches_per_run * args.steps,
            "clocks": m["clocks"],
            "roofline": roofline,
        }

    def progress(name, seconds, res):
        if rank == 0:
            wd.line = build_line(res)
        wd.leg(name, seconds)

    wd.leg("device-resident timing of the step", 420.0)
    m = measure(args.workload, args.steps, args.warmup, True, progress)
    extra = []
    if world == 8 and args.workload == "cfg5_1m_x30_k32" and not args.no_secondary:
        # BASELINE configs[4] at its largest size rides along on the 8-GPU line
        m4 = measure("cfg5_4m_x30_k32", 3, 2, False)
        if rank == 0:
            pairs4 = float(m4["n"]) ** 2
            extra.append({"workload": "cfg5_4m_x30_k32", "n_gpus": world, "ms_per_step": m4["dev_ms"] / 3,
                          "pairs_per_s": pairs4 * 3 / (m4["dev_ms"] * 1e-3),
                          "e2e_ms_per_step": m4["e2e_s"] / 3 * 1e3, "scan_ms": m4["scan_ms_max"],
                          "exact_rows_rank0": m4["exact_rows"], "verified_rows": m4["verified"],
                          "verified": m4["ok"],
                          "tiles_visited_fraction": m4["stats"]["visited_fraction"] if m4["stats"] else None})

    rc = 0
    if rank == 0:
        n, d, k, metric = m["n"], m["d"], m["k"], m["metric"]
        line = build_line(m)
        if m["ok"] is False:
            rc = 1
        if world == 1 and not args.no_secondary:
            del m["x"]
            line["config"]["secondary"] = extra           # filled in place: the watchdog's line sees finished records
            line["roofline"]["secondary"] = extra
            wd.line = line
            secondary_workloads(engine, synth, torch, tc_peak, peak_probe, out=extra, tick=wd.leg)
        if extra:
            line["config"]["secondary"] = extra
            line["roofline"]["secondary"] = extra
        if world == 1 and not args.no_cpu_baseline:
            wd.line = line
            wd.leg("cpu baseline (oracle, one thread)", 300.0)
            x = synth.cfg_knn_rows(n, d, WORKLOADS[args.workload][4])
            line["cpu_baseline"] = cpu_baseline_leg(x, k, metric)
        wd.done = True
        emit(line)
    wd.done = True
    if world > 1:
        tdist.destroy_process_group()
    return rc


if __name__ == "__main__":
    a = parse_args()
    sys.exit(reference_arm(a) if a.impl == "reference" else engine_arm(a))
