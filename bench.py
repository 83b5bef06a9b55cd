#!/usr/bin/env python
"""Benchmark of the contig-profile distance / neighbour-search hot path.

    python bench.py --gpus N --steps K --warmup W            # this engine (B200)
    python bench.py --impl reference --gpus N ...            # the reference's scipy CPU path

Workload (BASELINE.json: "pair-distances/s (fused top-k); 1M-contig kNN wall time at
1/2/4/8 B200"): the exact k=32 Euclidean neighbour graph of 1 000 000 synthetic contigs x
30 samples (configs[4] at 1 M; it fits one GPU).  One step = one complete pass of the path
over the batch: [all-gather of the row shards, N>1] -> centring -> FP32 tile staging ->
fused FP32 distance + top-k scan -> float64 re-rank + certificate -> exact rows.
With N GPUs the 1 M query rows are block-sharded (strong scaling: total work fixed).

Prints ONE JSON line on rank 0.  `value` = pair-distances/s with the float64 profiles
already resident in HBM; `e2e` = the same through the public API with pinned HOST buffers
(H2D of the profiles and D2H of the neighbour indices + distances inside the timed region).
"""
from __future__ import annotations

import argparse
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="cfg5_1m_x30_k32", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--order", default="auto", choices=["auto", "on", "off"],
                    help="locality-ordered sweep of the tensor-core scan (all-pairs, one GPU); auto: on for the "
                         "default workload, whose ordered results were checked identical on the GPU")
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
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "contigs": n, "dim": d, "k": k, "metric": metric,
                   "path": "scipy cdist (float64) + argpartition/sort per query row, all candidates"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} of {n} query rows x all {n} candidates per step, "
                                   f"{cores} processes"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def secondary_workloads(engine, synth, torch, peak):
    """The other single-GPU configurations of BASELINE.json, device-resident, for context
    (configs[1] 100 k x 10 k=16 both metrics; configs[2] BMU of 1 M contigs on a 64 x 64 mesh)."""
    out = []

    def run(name, q, c, k, metric, steps=10):
        same = c is None
        qd = torch.from_numpy(q).cuda()
        cd = qd if same else torch.from_numpy(c).cuda()
        pipe = engine.KnnPipeline(qd.shape[0], cd.shape[0], qd.shape[1], k, metric, same=same)
        for _ in range(3):
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
        ms = e0.elapsed_time(e1) / steps
        scan_ms = s0.elapsed_time(s1)
        pairs = float(qd.shape[0]) * float(cd.shape[0])
        tf = 2.0 * qd.shape[1] * pairs / (scan_ms * 1e-3) / 1e12
        pk = peak if metric == "euclidean" else peak / 2       # L1 has no FMA form: FADD rate
        rec = {"workload": name, "pairs_per_s": pairs / (ms * 1e-3), "ms_per_step": ms, "scan_ms": scan_ms,
               "scan_kernel": "tensor-core 3xFP16" if pipe.form == 3 else ("k=1 BMU scan" if pipe.bmu else "FP32 scan"),
               "scan_tflops_algorithmic": tf, "exact_rows": pipe.exact_row_count()}
        # fraction of the FP32 CUDA-core peak the ALGORITHMIC flops amount to (> 1 is possible on the tensor path)
        rec["vs_fp32_pipe_peak"] = tf / pk
        out.append(rec)

    x = synth.cfg_knn_rows(100_000, 10, 2)
    run("cfg2_100k_x10_k16_euclidean", x, None, 16, "euclidean")
    run("cfg2_100k_x10_k16_cityblock", x, None, 16, "cityblock")
    block, nodes = synth.cfg_bmu(1_000_000, 64)
    run("cfg3_bmu_1m_x4_on_64x64", block, nodes.reshape(-1, 4), 1, "euclidean")
    return out


def cpu_baseline_leg(x, k, metric, seconds_target=12.0):
    """Single-core oracle on a bounded sample of the same workload (rank 0, N=1 only)."""
    import oracle

    n = x.shape[0]
    rows = max(8, int(seconds_target * 1.2e7 // n))
    sel = np.sort(np.random.default_rng(0).choice(n, min(rows, n), replace=False))
    t0 = time.perf_counter()
    oracle.knn(x[sel], x, k, metric, chunk=8)
    dt = time.perf_counter() - t0
    return {"value": len(sel) * n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{len(sel)} of {n} query rows x all {n} candidates, scipy cdist + top-k, 1 thread, {dt:.1f} s"}


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
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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

    n, d, k, metric, seed = WORKLOADS[args.workload]
    x = synth.cfg_knn_rows(n, d, seed)                       # every rank draws the same array
    r0, r1 = gdist.shard_bounds(n, world, rank)
    n_local = r1 - r0
    x_host = torch.from_numpy(x[r0:r1]).pin_memory()         # this rank's shard, pinned host
    x_dev = x_host.to(dev)
    order = world == 1 and (args.order == "on" or (args.order == "auto" and args.workload == "cfg5_1m_x30_k32"))
    pipe = engine.KnnPipeline(max(n_local, 1), n, d, k, metric, same=(world == 1), device=dev, order=order)
    out_idx_host = torch.empty((n_local, k), dtype=torch.int32).pin_memory()
    out_dist_host = torch.empty((n_local, k), dtype=torch.float64).pin_memory()
    scan_ms = []

    def step(src_dev):
        """One pass of the hot path for this rank's rows; returns device (idx, dist)."""
        if world == 1:
            q = c = src_dev
        else:
            c = gdist.gather_rows(src_dev, n)
            q = c[r0:r1]
        pipe.reset()
        pipe.col_mean(c)
        pipe.prep_candidates(c)
        pipe.prep_queries(q)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.scan()
        e1.record()
        scan_ms.append((e0, e1))
        pipe.refine(q, c)
        pipe.exact_rows(q, c)
        return pipe.out_idx, pipe.out_dist

    def e2e_step():
        xd = x_host.to(dev, non_blocking=True)               # H2D of this step's profiles
        idx, dist_ = step(xd)
        out_idx_host.copy_(idx[:n_local], non_blocking=True)  # D2H of the step's result
        out_dist_host.copy_(dist_[:n_local], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    peak_probe = engine.fp32_peak_tflops()                   # measured FFMA2 peak of this GPU (uniform operand)
    peak_vector = engine.fp32_peak_tflops(form="vector")     # same stream with the scan's operand mix

    # ---- (A) device-resident throughput
    for _ in range(args.warmup):
        step(x_dev)
    barrier()
    scan_ms.clear()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        step(x_dev)
    t_end.record()
    barrier()
    dev_ms = max_over_ranks(t_start.elapsed_time(t_end))
    scan_avg_ms = sum(a.elapsed_time(b) for a, b in scan_ms) / len(scan_ms)
    scan_avg_ms_max = max_over_ranks(scan_avg_ms)
    exact_rows = pipe.exact_row_count()

    # ---- (B) end to end through the public API with host buffers
    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        pairs = float(n) * float(n)
        value = pairs * args.steps / (dev_ms * 1e-3)
        e2e_value = pairs * args.steps / e2e_s
        # roofline of the dominant kernel (the FP32 scan): 2*D flop per pair,
        # one launch covers this rank's n_local x n pairs
        launch_flop = 2.0 * d * float(n_local) * float(n)
        achieved = launch_flop / (scan_avg_ms * 1e-3) / 1e12
        nominal = None
        props = torch.cuda.get_device_properties(dev)
        if clocks and clocks.get("sm_max_mhz"):
            nominal = props.multi_processor_count * 128 * 2 * clocks["sm_max_mhz"] * 1e6 / 1e12
        traffic = None
        prof = ROOT / "profiles" / "scan_traffic.json"
        if prof.exists():
            try:
                traffic = json.loads(prof.read_text()).get(args.workload)
            except (ValueError, OSError):
                traffic = None
        tc = pipe.form == 3                                    # GMD_FORM_TC: the tensor-core scan
        if tc:
            # executed tensor work: K padded to 32, three FP16 passes (hi.hi + hi.lo + lo.hi)
            executed = 2.0 * 32 * 3 * float(n_local) * float(n) / (scan_avg_ms * 1e-3) / 1e12
            tc_peak, tc_src = 1590.0, "MEASURED_PEAKS.json absent: fallback 1590 TFLOP/s (B200_PROFILING.md)"
            try:
                mp = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
                tc_peak = float(mp["bf16_tflops"])
                tc_src = "MEASURED_PEAKS.json bf16_tflops (FP16 tensor rate = bf16 rate), burst figure"
            except (OSError, ValueError, KeyError):
                pass
            roofline = {"bound": "tensor", "kernel": "gmd_tc::scan_tc_kernel", "achieved": achieved,
                        "peak": tc_peak, "unit": "TFLOP/s", "frac": achieved / tc_peak, "peak_source": tc_src,
                        "note": "achieved = ALGORITHMIC flops (2*D per pair); the kernel executes 2*32*3 per pair "
                                "(K padded to 32, 3 x FP16 hi/lo split for FP32-class accuracy)",
                        "executed_tflops": executed, "frac_executed": executed / tc_peak,
                        "fp32_pipe_peak": peak_probe, "vs_fp32_pipe_peak": achieved / peak_probe,
                        "scan_ms_per_launch": scan_avg_ms_max, "flop_per_launch": launch_flop, "traffic": traffic}
        else:
            roofline = {"bound": "fp32", "kernel": "gmd_scan::scan_topk_kernel", "achieved": achieved,
                        "peak": peak_probe, "unit": "TFLOP/s", "frac": achieved / peak_probe,
                        "peak_source": "FFMA2 probe kernel (uniform-register multiplicand) measured on this GPU "
                                       "in this run; MEASURED_PEAKS.json has no FP32 entry",
                        "peak_vector_operands": peak_vector,
                        "peak_nominal": nominal, "frac_of_nominal": (achieved / nominal) if nominal else None,
                        "scan_ms_per_launch": scan_avg_ms_max, "flop_per_launch": launch_flop, "traffic": traffic}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16x3" if tc else "f32", "data": "synthetic",
            "config": {"workload": args.workload, "contigs": n, "dim": d, "k": k, "metric": metric,
                       "parallelism": f"query rows block-sharded over {world} GPU(s); one NCCL all-gather of the "
                                      f"float64 rows per step" if world > 1 else "single GPU",
                       "l2": "inputs exceed L2 (float64 profiles %d MB, FP32 tiles %d MB)" %
                             (n * d * 8 // 2**20, pipe.ctiles.numel() * 4 // 2**20),
                       "arithmetic": ("tcgen05 3xFP16 hi/lo split (FP32-class) " if pipe.form == 3 else "FP32 ") +
                                     "fused distance+top-k scan proposes k+8 candidates; float64 re-rank "
                                     "and certificate make the result exact" +
                                     ("; profiles staged in a locality-preserving order (nearest of %d sampled "
                                      "profiles, counting sort, inside the timed step), every query tile sweeps "
                                      "from its own position" % pipe.order_ncl if pipe.order else ""),
                       "ordered_sweep": bool(pipe.order),
                       "form": pipe.form, "kk": pipe.kk, "nsplit": pipe.nsplit, "exact_rows_last_step": exact_rows},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_s / args.steps * 1e3,
                    "h2d_bytes_per_step": int(x_host.numel() * 8 * world),
                    "d2h_bytes_per_step": int((out_idx_host.numel() * 4 + out_dist_host.numel() * 8) * world)},
            "gpu_launches": pipe.launches_per_run * args.steps,
            "clocks": clocks,
            "roofline": roofline,
        }
        if world == 1 and not args.no_secondary:
            line["secondary"] = secondary_workloads(engine, synth, torch, peak_probe)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(x, k, metric)
        emit(line)
    if world > 1:
        tdist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    sys.exit(reference_arm(a) if a.impl == "reference" else engine_arm(a))
