#!/usr/bin/env python
"""Instruction count of a kernel per source region (nvdisasm -g line info): tools/sass_lines.py <cubin> <file-substring> [bucket]
Prints, in address order, runs of consecutive instructions attributed to the same bucket of source lines."""
import re
import subprocess
import sys

cubin, fsub = sys.argv[1:3]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cur = ("?", 0)
runs = []
for ln in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        key = (cur[0], cur[1] // bucket * bucket) if fsub in cur[0] else (cur[0], 0)
        if runs and runs[-1][0] == key:
            runs[-1][1] += 1
        else:
            runs.append([key, 1, int(m.group(1), 16)])
# merge tiny runs for readability
tot = {}
for key, n, addr in runs:
    tot[key] = tot.get(key, 0) + n
for key in sorted(tot):
    print(f"{key[0]}:{key[1]:<5d} {tot[key]:6d}")
print("total", sum(tot.values()))
