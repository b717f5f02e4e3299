#!/usr/bin/env python
"""SASS census of the shipped library: for every kernel matching a pattern, the instruction count, the registers, the
opcode histogram and the loops (backward branches) with their instruction mix -- the static evidence behind the
instruction-count statements of DESIGN.md / profiles/README.md (DFMA/DMUL per EM pass, instructions per individual of
the x87 loop, UBLKCP/SYNCS of the TMA row copies of the dense kernel).

  python scripts/sass_excerpt.py ngsld_b200/libngsld_b200.so 'em_cell_kernelILi6ELb1ELi4|em_warp_kernelILi6ELb0ELb0ELi1' > profiles/r2_sass_census.txt"""
import collections
import re
import subprocess
import sys


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, res = None, collections.OrderedDict()
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            res[cur] = []
            continue
        m = re.match(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
        if m and cur:
            res[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return res


def opcode(txt):
    txt = re.sub(r"^@!?U?P\d+\s+", "", txt)
    return txt.split()[0].split(".")[0]


def main():
    path, pat = sys.argv[1], sys.argv[2]
    regs = {}
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    name = None
    for ln in res.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+).*SHARED:(\d+)", ln)
        if m and name:
            regs[name] = (int(m.group(1)), int(m.group(2)))
    for fn, ins in functions(path).items():
        if not re.search(pat, fn):
            continue
        demangled = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
        print(f"=== {demangled}")
        r = regs.get(fn)
        print(f"    {len(ins)} instructions" + (f", {r[0]} registers, {r[1]} B static shared memory" if r else ""))
        hist = collections.Counter(opcode(t) for _, t in ins)
        print("    opcodes: " + ", ".join(f"{k} {v}" for k, v in hist.most_common(24)))
        wide = sum(1 for _, t in ins if "IMAD.WIDE" in t)
        print(f"    of which IMAD.WIDE {wide}, LDG.E.128 {sum(1 for _, t in ins if 'LDG.E.128' in t)}, "
              f"LDG.E.64 {sum(1 for _, t in ins if 'LDG.E.64' in t)}, UBLKCP {hist.get('UBLKCP', 0)}, SYNCS {hist.get('SYNCS', 0)}, "
              f"LDL/STL (spills) {hist.get('LDL', 0) + hist.get('STL', 0)}")
        idx = {a: i for i, (a, _) in enumerate(ins)}
        loops = []
        for i, (a, t) in enumerate(ins):
            m = re.search(r"\bBRA\S*\s+(?:.*?)(0x[0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in idx:
                    loops.append((idx[tgt], i))
        for s, e in sorted(loops):
            if e - s + 1 < 24:
                continue
            h = collections.Counter(opcode(t) for _, t in ins[s:e + 1])
            print(f"    loop {ins[s][0]:#07x}..{ins[e][0]:#07x}: {e - s + 1:5d} instr  " +
                  ", ".join(f"{k} {v}" for k, v in h.most_common(10)))
        print()


if __name__ == "__main__":
    main()
