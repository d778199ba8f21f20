import csv, io, sys
txt = open(sys.argv[1]).read()
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sections = txt.split('"Kernel Name",')
seen = set()
for sec in sections[1:]:
    lines = sec.split('\n')
    name = lines[0][:100]
    rdr = csv.reader(io.StringIO('\n'.join(lines[1:])))
    hdr = next(rdr)
    iS = hdr.index("Warp Stall Sampling (All Samples)"); iI = hdr.index("Instructions Executed"); iSrc = hdr.index("Source")
    rows = [r for r in rdr if len(r) > iI]
    tot = sum(int(r[iS]) for r in rows)
    key = (name, tot)
    if key in seen: continue
    seen.add(key)
    print("==", name, "total samples", tot, "instr rows", len(rows))
    for r in sorted(rows, key=lambda r: -int(r[iS]))[:N]:
        print(f"{int(r[iS]):7d} {100*int(r[iS])/max(tot,1):5.1f}%  exec={r[iI]:>9s}  {r[iSrc].strip()[:100]}")
