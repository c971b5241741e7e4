#!/usr/bin/env python
"""Join an ncu report's per-SASS-instruction metrics with nvdisasm line info -> per-source-line table.

usage: tools/ncu_lines.py <report.ncu-rep> <cubin> <kernel-mangled-substring> [top]
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# split into functions
line_of = {}
cur_fn = None
cur_line = None
in_fn = False
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        cur_fn = m.group(1)
        in_fn = kern in cur_fn
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur_line, m.group(2))

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
base = None
agg = defaultdict(lambda: defaultdict(float))
tot = defaultdict(float)
keys = ["# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive", "stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_math", "stall_wait", "stall_lg", "stall_not_selected", "stall_branch_resolving", "stall_no_inst"]
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    addr = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
    if base is None:
        base = addr
    off = addr - base
    src = line_of.get(off, ((None, -1), "?"))[0] or ("?", -1)
    for k in keys:
        if k in col:
            try:
                v = float(r[col[k]] or 0)
            except ValueError:
                v = 0
            agg[src][k] += v
            tot[k] += v
print("totals:", {k: int(v) for k, v in tot.items()})
print(f"{'file:line':34s} {'samp%':>6s} {'inst%':>6s} {'shWf%':>6s} {'exWf%':>6s}  top stalls")
for src, m in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    st = sorted(((k, v) for k, v in m.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:3]
    print(f"{src[0]}:{src[1]:<6d}".ljust(34), f"{100*m['# Samples']/max(tot['# Samples'],1):6.2f} {100*m['Instructions Executed']/max(tot['Instructions Executed'],1):6.2f} "
          f"{100*m['L1 Wavefronts Shared']/max(tot['L1 Wavefronts Shared'],1):6.2f} {100*m['L1 Wavefronts Shared Excessive']/max(tot['L1 Wavefronts Shared'],1):6.2f} ",
          " ".join(f"{k[6:]}={100*v/max(m['# Samples'],1):.0f}%" for k, v in st))
