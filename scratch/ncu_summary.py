"""Markdown table from `ncu -i X.ncu-rep --page raw --csv`: one row per profiled launch.
python scratch/ncu_summary.py raw.csv [alg_bytes_json]"""
import csv, sys, re, json
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
alg = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else {}
def f(d, k, scale=1.0):
    try: return float(d[ix[k]].replace(",", "")) * scale
    except Exception: return float("nan")
stall_names = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
print("| kernel | grid x block | dur us | DRAM rd MB | DRAM wr MB | L2 traffic MB | warp-inst M | IPC/SM | warps active % | regs | top stalls (pc samples) |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
seen = set()
for d in data:
    name = re.sub(r"\(.*", "", d[ix["Kernel Name"]]).replace("void ", "").replace("pob::", "")
    key = (name, d[ix["Grid Size"]], d[ix["ID"]] if "ID" in ix else None)
    if key in seen: continue
    seen.add(key)
    dur = f(d, "gpu__time_duration.sum")
    unit = rows[1][ix["gpu__time_duration.sum"]]
    dur_us = dur / 1000 if unit in ("ns", "nsecond") else (dur if unit in ("us", "usecond") else dur * 1000)
    def mb(k):
        v = f(d, k); u = rows[1][ix[k]] if k in ix else ""
        return v / 1e6 if u == "byte" else (v / 1e3 if u == "Kbyte" else (v if u == "Mbyte" else v * 1e3))
    st = sorted(((f(d, s), s.replace("smsp__pcsamp_warps_issue_stalled_", "")) for s in stall_names), reverse=True)
    tot = sum(v for v, _ in st if v == v) or 1
    tops = ", ".join(f"{n} {100*v/tot:.0f}%" for v, n in st[:3])
    print(f"| `{name}` | {d[ix['Grid Size']]} x {d[ix['Block Size']]} | {dur_us:.1f} | {mb('dram__bytes_read.sum'):.1f} | {mb('dram__bytes_write.sum'):.1f} | "
          f"{f(d,'lts__t_sectors.sum')*32/1e6:.1f} | {f(d,'smsp__inst_executed.sum')/1e6:.2f} | {f(d,'sm__inst_executed.avg.per_cycle_active'):.2f} | "
          f"{f(d,'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {f(d,'launch__registers_per_thread'):.0f} | {tops} |")
