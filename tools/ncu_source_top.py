"""Summarise an `ncu --page source --csv` dump: stall samples per source line / SASS instruction (top N), and
totals per stall reason.   python tools/ncu_source_top.py dump.csv [N]"""
import csv, sys, collections
path = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
def num(r, h):
    try: return float(r[col[h]])
    except (ValueError, KeyError): return 0.0
total = sum(num(r, "# Samples") for r in data)
print("kernel:", rows[0][1][:120] if rows[0] else "")
print("total samples", total, "instructions", len(data))
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: sum(num(r, h) for r in data) for h in reasons}
print("stall reasons:", ", ".join(f"{h[6:]}={v:.0f} ({100*v/max(total,1):.1f}%)" for h, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v > 0))
print("\ntop instructions by samples:")
for i, r in sorted(enumerate(data), key=lambda ir: -num(ir[1], "# Samples"))[:N]:
    top = sorted(((h[6:], num(r, h)) for h in reasons), key=lambda kv: -kv[1])[:2]
    print(f"{i:5d} {num(r,'# Samples'):7.0f} {100*num(r,'# Samples')/max(total,1):5.1f}%  exec={num(r,'Instructions Executed'):9.0f}  {r[col['Source']][:70]:70s} {top}")
