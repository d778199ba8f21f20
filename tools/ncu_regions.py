"""Split an `ncu --page source --csv` dump of a kernel into the regions between its barrier / mbarrier / exit
instructions and report, per region, the warp-stall samples and executed warp instructions; plus totals per stall
reason.   python tools/ncu_regions.py dump.csv [label ...]"""
import csv, sys
path = sys.argv[1]
rows = list(csv.reader(open(path)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; col = {x: i for i, x in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
def num(r, k):
    try: return float(r[col[k]])
    except (ValueError, KeyError): return 0.0
tot = sum(num(r, "# Samples") for r in data)
reasons = [x for x in hdr if x.startswith("stall_") and "Not Issued" not in x]
print(f"kernel: {rows[0][1][:100]}")
print(f"total samples {tot:.0f}; stall reasons: " + ", ".join(f"{x[6:]} {100 * sum(num(r, x) for r in data) / tot:.1f}%" for x in
      sorted(reasons, key=lambda x: -sum(num(r, x) for r in data))[:7]))
print("| SASS index range | samples | share | warp instr executed | region ends at |\n|---|---|---|---|---|")
acc = ex = 0; start = 0
for i, r in enumerate(data):
    s = r[col["Source"]]
    acc += num(r, "# Samples"); ex += num(r, "Instructions Executed")
    if "BAR.SYNC" in s or "SYNCS.PHASECHK" in s or "EXIT" in s:
        if acc > 0.003 * tot:
            print(f"| {start}-{i} | {acc:.0f} | {100 * acc / tot:.1f}% | {ex:.3g} | `{s.strip()[:50]}` |")
        acc = ex = 0; start = i + 1
