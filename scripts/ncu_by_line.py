#!/usr/bin/env python
"""Aggregate an ncu SASS source page by CUDA source line.

usage: ncu_by_line.py REPORT.ncu-rep LIB.so KERNEL_SUBSTR [top_n]

The CSV source page of ncu carries per-SASS-instruction counters but no line numbers; nvdisasm -g
of the same cubin carries the line of every instruction.  Both list the kernel's instructions in
address order, so they are joined by index.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    rep, lib, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    # the page lists one section per profiled launch: "Kernel Name" row, header row, one row per SASS instruction
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    sect = next((i for i in starts if kern in rows[i][1]), None)
    if sect is None:
        print("no profiled kernel matches", kern, file=sys.stderr)
        return 1
    kname = rows[sect][1]
    hdr = rows[sect + 1]
    end = min([i for i in starts if i > sect] + [len(rows)])
    body = [r for r in rows[sect + 2:end] if len(r) == len(hdr)]
    col = {h: i for i, h in enumerate(hdr)}
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
        cubins = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")]
        dis = ""
        for c in cubins:
            dis += subprocess.run(["nvdisasm", "-g", "-c", c], capture_output=True, text=True).stdout
    # locate the function whose demangled name matches the profiled kernel: pick by instruction count
    funcs, cur, name, line = {}, None, None, ("?", 0)
    for ln in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            name, cur = m.group(1), []
            funcs[name] = cur
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), line, m.group(2).strip()))
    # mangle the template arguments of the profiled instance: <(int)256, (bool)1, (bool)0> -> ILi256ELb1ELb0E
    targs = "".join(("Li" if t == "int" else "Lb") + v + "E" for t, v in re.findall(r"\((int|bool)\)(\d+)", kname))
    cands = [(n, f) for n, f in funcs.items() if kern.split("<")[0] in n and len(f) == len(body)
             and (not targs or ("I" + targs + "E") in n)]
    if not cands:
        print("no function with", len(body), "instructions matches", kern, file=sys.stderr)
        print({n: len(f) for n, f in funcs.items() if kern.split('<')[0] in n}, file=sys.stderr)
        return 1
    fname, f = cands[0]
    print(f"# {kname}  ->  {fname}  ({len(body)} SASS instructions)")
    agg = defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0]
    for (off, line, txt), r in zip(f, body):
        ie = int(float(r[col["Instructions Executed"]] or 0))
        te = int(float(r[col["Predicated-On Thread Instructions Executed"]] or 0))
        ss = int(float(r[col["# Samples"]] or 0))
        a = agg[line]
        a[0] += ie
        a[1] += te
        a[2] += ss
        a[3] += 1
        tot[0] += ie
        tot[1] += te
        tot[2] += ss
    print(f"# total warp-instr {tot[0]:,}  thread-instr {tot[1]:,}  avg active threads {tot[1] / max(1, tot[0]):.2f}  samples {tot[2]:,}")
    print(f"{'file:line':28s} {'sass':>5s} {'warp-instr%':>11s} {'avg-thr':>8s} {'samples%':>9s}")
    for line, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{line[0] + ':' + str(line[1]):28s} {a[3]:5d} {100 * a[0] / tot[0]:10.2f}% {a[1] / max(1, a[0]):8.2f} {100 * a[2] / max(1, tot[2]):8.2f}%")
    # by file
    byf = defaultdict(lambda: [0, 0, 0])
    for line, a in agg.items():
        b = byf[line[0]]
        b[0] += a[0]
        b[1] += a[1]
        b[2] += a[2]
    print("# by file")
    for fn, b in sorted(byf.items(), key=lambda kv: -kv[1][0]):
        print(f"{fn:28s} {100 * b[0] / tot[0]:10.2f}% {b[1] / max(1, b[0]):8.2f} {100 * b[2] / max(1, tot[2]):8.2f}%")
    return 0


if __name__ == "__main__":
    sys.exit(main())
