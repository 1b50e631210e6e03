"""Per-source-line totals of an ncu capture taken with --import-source on (kernel built with -lineinfo):

    python scripts/ncu_lines.py gpurun_out/prof_ni_r2.ncu-rep [file-substring] [top]

Prints the source lines that execute the most warp instructions, with their share of the stall samples - the map from
a kernel's stages to where its instructions go."""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
path = ""
hdr = None
agg = defaultdict(lambda: [0, 0, ""])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        IE, S = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and r[0].isdigit():
        key = (path.split("/")[-1], int(r[0]))
        num = lambda v: int(v) if v.strip().lstrip("-").isdigit() else 0
        agg[key][0] += num(r[IE])
        agg[key][1] += num(r[S])
        agg[key][2] = r[1].strip()[:110]
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[1] for v in agg.values()) or 1
print(f"total warp instructions {tot_i}, samples {tot_s}")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if want in f:
        print(f"{f}:{ln:4d}  {100 * v[0] / tot_i:5.2f} % instr  {100 * v[1] / tot_s:5.2f} % samples  | {v[2]}")
