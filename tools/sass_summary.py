"""Per-kernel SASS evidence from the built library: counts of the mnemonics that prove what a kernel uses
(bulk-async copies, mbarrier, warp reductions, vector atomics, DSMEM stores, local-memory spills).

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pointcloudpdf_b200", "lib", "libpointops_b200.so")
PATTERNS = [("UBLKCP", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS\."), ("ST.ASYNC / STAS (DSMEM)", r"\bSTAS\b|ST\.ASYNC|\bSTAS\."),
            ("REDUX / CREDUX", r"\bC?REDUX"), ("RED.*.F32x4", r"\bRED\.[A-Z0-9_.]*F32x4|\bREDG?\.[A-Z0-9_.]*\.128"),
            ("ATOM / RED (any)", r"\b(ATOM|ATOMG|ATOMS|RED|REDG)\b|\b(ATOM|ATOMG|ATOMS|RED|REDG)\."), ("UCGABAR (cluster barrier)", r"UCGABAR"),
            ("LDG.E.128", r"LDG\.E\.128|LDG\.E\.[A-Z.]*128"), ("STG.E.128", r"STG\.E\.128|STG\.E\.[A-Z.]*128"),
            ("LDS.128", r"LDS\.128"), ("STS.128", r"STS\.128"), ("STL/LDL (local)", r"\b(STL|LDL)\b|\b(STL|LDL)\."),
            ("FFMA", r"\bFFMA"), ("HMMA (mma.sync tensor core)", r"\bHMMA"), ("LDGSTS (cp.async)", r"\bLDGSTS"), ("BAR.SYNC", r"BAR\.SYNC")]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kern, counts, total = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "")
        counts[kern] = collections.Counter(); total[kern] = 0
        continue
    if kern and re.search(r"/\*[0-9a-f]{4}\*/", line):
        total[kern] += 1
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[kern][name] += 1
want = sys.argv[1:] or ["fps_merge", "fps_chain", "fps_cluster", "knn_grid", "aggregation_fwd_pipe", "aggregation_bwd_fast", "gather_rows_fast<float, 8, 8",
                        "group_xyz", "scatter_rows_fast<8", "reduce_neighbours_fast<8", "pt_layer_tile", "score_fused", "linear_tile_kernel<128", "linear_mma_kernel", "fps_curve"]
print("# SASS evidence per kernel (`cuobjdump -sass pointcloudpdf_b200/lib/libpointops_b200.so`, sm_100a)\n")
print("Counts of SASS instructions by mnemonic, per kernel instantiation (static counts, not executed counts).\n")
cols = [n for n, _ in PATTERNS]
print("| kernel | instr | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for k, c in counts.items():
    if any(w in k for w in want):
        print(f"| `{k[:110]}` | {total[k]} | " + " | ".join(str(c.get(n, 0)) for n in cols) + " |")
