#!/usr/bin/env python
"""Turn gpurun_out/ artefacts into the tracked summaries under profiles/.

    python tools/make_profiles.py launches gpurun_out/launches_r1.csv profiles/launches_r1.md "<command>"
    python tools/make_profiles.py full gpurun_out/scan_bench_r1.ncu-rep profiles/scan_1m_d30_r1.md cfg5_1m_x30_k32 "<command>"

`full` also records the kernel's DRAM traffic per launch in profiles/scan_traffic.json
(bench.py copies it into roofline.traffic)."""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void ", "", name)
    return name if len(name) < 100 else name[:97] + "..."


def launches(csv_path, out_path, cmd):
    rows = [r for r in csv.DictReader(l for l in open(csv_path) if l.startswith('"'))]
    agg = OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] == "us":
            ns *= 1e3
        elif r["Metric Unit"] == "ms":
            ns *= 1e6
        elif r["Metric Unit"] == "s":
            ns *= 1e9
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    lines = [f"# Launch list ({Path(csv_path).name})", "",
             f"Command (under `ncu --metrics gpu__time_duration.sum --clock-control none`): `{cmd}`", "",
             "Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.", "",
             "| kernel | launches | total ms | share | grid | block |", "|---|---:|---:|---:|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.2f} % | {a[2]} | {a[3]} |")
    lines += ["", f"Total device time in the capture: {tot / 1e6:.1f} ms over {sum(a[0] for a in agg.values())} launches.", ""]
    Path(out_path).write_text("\n".join(lines))
    print(f"wrote {out_path}")


WANT = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory / CTA"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe instructions % of peak"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe instructions % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe instructions % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
])
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}


def full(rep, out_path, workload, cmd):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --set full: {Path(rep).name}", "", f"Command: `{cmd}`", ""]
    traffic = None
    for vals in rows[2:]:
        m = dict(zip(hdr, zip(units, vals)))
        out += [f"## `{short(m['Kernel Name'][1])}`", "", "| metric | value |", "|---|---|"]
        for key, label in WANT.items():
            if key in m:
                out.append(f"| {label} (`{key}`) | {m[key][1]} {m[key][0]} |")
        rd = float(m["dram__bytes_read.sum"][1]) * UNIT.get(m["dram__bytes_read.sum"][0], 1.0)
        wr = float(m["dram__bytes_write.sum"][1]) * UNIT.get(m["dram__bytes_write.sum"][0], 1.0)
        traffic = rd + wr
        out += ["", f"DRAM traffic per launch = {traffic / 1e9:.3f} GB.", "", "Warp stall reasons (cycles per issued instruction):", ""]
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                v = float(m[h][1])
                if v >= 0.02:
                    out.append(f"* {h.split('stalled_')[1].split('_per_issue')[0]}: {v:.3f}")
        out.append("")
        break
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 3:
        h = srows[1]
        ix = {c: i for i, c in enumerate(h)}
        data = srows[2:]

        def f(r, k):
            try:
                return float(r[ix[k]])
            except (ValueError, KeyError, IndexError):
                return 0.0

        tot = sum(f(r, "# Samples") for r in data) or 1.0
        ops = {}
        for r in data:
            op = r[ix["Source"]].split()
            op = next((t for t in op if not t.startswith("@")), "?").split(".")[0] if op else "?"
            o = ops.setdefault(op, [0.0, 0.0])
            o[0] += f(r, "Instructions Executed")
            o[1] += f(r, "# Samples")
        ti = sum(o[0] for o in ops.values()) or 1.0
        out += ["SASS mix of the launch (warp-level executed instructions, PC-sampling share):", "",
                "| opcode | executed | share of instructions | share of samples |", "|---|---:|---:|---:|"]
        for op, o in sorted(ops.items(), key=lambda kv: -kv[1][0])[:14]:
            out.append(f"| {op} | {o[0]:.3g} | {100 * o[0] / ti:.1f} % | {100 * o[1] / tot:.1f} % |")
        out.append("")
    Path(out_path).write_text("\n".join(out))
    print(f"wrote {out_path}")
    if workload and traffic is not None:
        tj = ROOT / "profiles" / "scan_traffic.json"
        d = json.loads(tj.read_text()) if tj.exists() else {}
        d[workload] = traffic
        tj.write_text(json.dumps(d, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "", sys.argv[5] if len(sys.argv) > 5 else "")
