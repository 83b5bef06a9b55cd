"""SASS opcode counts of the tensor-core kernels in the shipped library -> profiles/sass_<tag>.md
    python tools/sass_table.py [tag]"""
import re, subprocess, sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
sass = subprocess.run(["cuobjdump", "-sass", str(ROOT / "groopm_b200" / "libgmdist.so")], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
kern, cur = {}, None
it = iter(names)
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = next(it)
        cur = cur.replace("void ", "").replace("(anonymous namespace)", "<unnamed>")
        cur = re.sub(r"\((int|bool)\)", "", cur)              # template arguments are printed as casts
        depth = 0
        for pos in range(len(cur) - 1, -1, -1):                # drop the parameter list: the last balanced (...)
            depth += cur[pos] == ")"
            depth -= cur[pos] == "("
            if depth == 0 and cur[pos] == "(":
                cur = cur[:pos]
                break
        kern[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        kern[cur][op] += 1
        kern[cur]["_n"] += 1
        if re.search(r"\b(LD|ST)\.E", line):
            kern[cur]["_generic"] += 1
        if op in ("LDS", "STS", "LDSM", "STSM"):
            kern[cur]["_shared"] += 1
cols = ["UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "SYNCS", "FFMA2"]
out = [f"# SASS of the shipped library ({tag})", "",
       "`cuobjdump -sass groopm_b200/libgmdist.so` (nvcc 12.9.86, `-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo`), instruction counts per",
       "kernel for the opcodes that prove the Blackwell path: `UTCHMMA` = tcgen05.mma kind::f16, `LDTM` = tcgen05.ld (tensor memory ->",
       "registers), `UBLKCP` = cp.async.bulk (TMA engine, no tensor map), `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier operations, `FFMA2` =",
       "packed FP32 FMA.  `LDS/STS` against `LD.E/ST.E`: shared-memory accesses that the compiler knows to be shared against generic ones",
       "(a kernel whose shared base went through an integer lost the address space: 353 generic accesses in the pruned scan before the fix).", "",
       "| kernel | SASS instructions | " + " | ".join(cols) + " | LDS/STS | LD.E/ST.E |", "|---|---:|" + "---:|" * (len(cols) + 2)]
for k in sorted(kern):
    if "scan_tc" in k:
        c = kern[k]
        out.append(f"| `{k}` | {c['_n']} | " + " | ".join(str(c[x]) for x in cols) + f" | {c['_shared']} | {c['_generic']} |")
fp = [k for k in kern if "scan_topk_kernel" in k or "scan_bmu_kernel" in k]
out += ["", f"The FP32 scans (`scan_topk_kernel<D, FORM, NW, NST>`, `scan_bmu_kernel<D, NW, NST>`: {len(fp)} instantiations) hold no tensor-core",
        f"instruction; they move their tiles with `UBLKCP` (up to {max(kern[k]['UBLKCP'] for k in fp)} per kernel) and compute with up to "
        f"{max(kern[k]['FFMA2'] for k in fp)} `FFMA2` per unrolled tile step.", "",
        "Reading the tensor-core rows: the full sweep `scan_tc_kernel<NKB, SWEEP>` issues 6 x NKB `UTCHMMA` per candidate tile (hi.hi + hi.lo + lo.hi,",
        "two K = 16 steps per 32-column block) and reads an accumulator row with ONE `LDTM.x128`; the pruned scan `scan_tc_prune_kernel<NKB, QT, TQP>`",
        "issues 6 x NKB x QT per tile (QT query tiles against one candidate stage) and reads a row as two `LDTM.x64`.  Every operand travels by `UBLKCP`.",
        "", "Other kernels of the hot path:", "", "| kernel | SASS instructions | LDG | LDS/STS | LD.E/ST.E | DADD+DMUL+DFMA |", "|---|---:|---:|---:|---:|---:|"]
for k in sorted(kern):
    if any(s in k for s in ("refine_kernel", "prep_tc_kernel", "resolve_kernel", "pick_kernel", "adjacency_kernel")):
        c = kern[k]
        out.append(f"| `{k}` | {c['_n']} | {c['LDG']} | {c['_shared']} | {c['_generic']} | {c['DADD'] + c['DMUL'] + c['DFMA']} |")
(ROOT / "profiles" / f"sass_{tag}.md").write_text("\n".join(out) + "\n")
print("wrote", f"profiles/sass_{tag}.md")
