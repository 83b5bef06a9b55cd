"""Summarise an .ncu-rep: key raw metrics + stall samples by code segment (needs ncu on PATH)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active"]
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
print("total samples", tot)
prev = None; seg = []; start = None
for r in data:
    e = int(f(r, "Instructions Executed")); a = r[ix["Address"]][-5:]
    if prev is None or abs(e - prev) > 0.02 * max(e, prev, 1):
        if start is not None: seg.append((start, a, prev, acc, cnt))
        start = a; acc = 0; cnt = 0
    acc += f(r, "# Samples"); cnt += 1; prev = e
seg.append((start, a, prev, acc, cnt))
for s in seg:
    if s[3] / tot > 0.004: print(f"seg {s[0]}-{s[1]} exec/instr {s[2]} ninstr {s[4]} samples {100*s[3]/tot:.2f}%")
cats = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for c in cats:
    v = sum(f(r, c) for r in data) / tot
    if v > 0.005: print(c, round(v, 4))
if len(sys.argv) > 2:
    top = sorted(data, key=lambda r: -f(r, "# Samples"))[: int(sys.argv[2])]
    for r in top:
        st = {c: int(f(r, c)) for c in cats if f(r, c) > 0.1 * f(r, "# Samples")}
        print(r[ix["Address"]][-5:], r[ix["Source"]][:64].ljust(64), int(f(r, "# Samples")), int(f(r, "Instructions Executed")), st)
