"""Per-instruction samples of an address range of an .ncu-rep source page:
python tools/ncu_region.py rep.ncu-rep <lo_hex_suffix> <hi_hex_suffix> [min_samples]"""
import csv, io, subprocess, sys
rep, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
mins = int(sys.argv[4]) if len(sys.argv) > 4 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
cats = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(f(r, "# Samples") for r in data)
reg = 0
for r in data:
    a = int(r[ix["Address"]][-5:], 16)
    if lo <= a <= hi:
        sm = f(r, "# Samples"); reg += sm
        if sm >= mins:
            st = {c[6:]: int(f(r, c)) for c in cats if f(r, c) > 0.2 * max(sm, 1)}
            print(f"{a:05x} {r[ix['Source']][:70]:70s} {int(sm):7d} {int(f(r,'Instructions Executed')):10d} {st}")
print(f"region samples {reg:.0f} = {100*reg/tot:.2f}% of {tot:.0f}")
