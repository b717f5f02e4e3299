#!/usr/bin/env python
"""Attribute the per-instruction counters of an ncu --set full --import-source capture to source lines.
usage: python scripts/ncu_lines.py <rep.ncu-rep> <object-or-cubin with the same build> <kernel mangled-name substring> [top]
Joins `ncu --page source --csv` (per SASS instruction: executed count, stall samples) with `nvdisasm -g` (offset -> file:line)
and prints instructions executed and stall samples per file, per line (top N) -- no GPU needed."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
with tempfile.TemporaryDirectory() as td:
    cubin = obj
    if not obj.endswith(".cubin"):
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
        cubin = os.path.join(td, [f for f in os.listdir(td) if f.endswith(".cubin")][0])
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of, cur, inside = {}, ("?", 0), False
for ln in dis.splitlines():
    if ln.startswith("//-----") and ".text." in ln:
        inside = kname in ln
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][ix["Address"]], 16)
by_file, by_line, samp_file, samp_line = Counter(), Counter(), Counter(), Counter()
for r in rows[2:]:
    off = int(r[ix["Address"]], 16) - base
    (f, l), _ = line_of.get(off, (("?", 0), ""))
    n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    by_file[f] += n; by_line[(f, l)] += n; samp_file[f] += s; samp_line[(f, l)] += s
tot, stot = sum(by_file.values()), sum(samp_file.values())
print(f"total warp instructions {tot}, stall samples {stot}")
for f, n in by_file.most_common():
    print(f"  {f:28s} {100*n/tot:6.2f}% instr  {100*samp_file[f]/max(1,stot):6.2f}% samples")
print("top lines:")
for (f, l), n in by_line.most_common(top):
    print(f"  {f}:{l:<5d} {100*n/tot:6.2f}% instr  {100*samp_line[(f,l)]/max(1,stot):6.2f}% samples")
