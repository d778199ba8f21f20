"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.
python tools/launch_summary.py launches.csv [rooms]"""
import csv, gzip, sys, re, collections
rooms = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = []
with (gzip.open(sys.argv[1], "rt", newline="") if sys.argv[1].endswith(".gz") else open(sys.argv[1], newline="")) as f:
    lines = [l for l in f if not l.startswith("==")]
rdr = csv.DictReader(lines)
agg = collections.OrderedDict()
for r in rdr:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)[:70]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1000.0)
    d = agg.setdefault(name, [0, 0.0])
    d[0] += 1; d[1] += us
tot = sum(v[1] for v in agg.values())
print(f"total device time {tot/rooms:.1f} us per room over {rooms:g} rooms, {sum(v[0] for v in agg.values())/rooms:.0f} launches per room")
for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{us/rooms:10.1f} us  {100*us/tot:5.1f}%  {cnt/rooms:7.1f} x  {name}")
