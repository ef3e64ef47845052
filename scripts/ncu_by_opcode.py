#!/usr/bin/env python
"""Executed warp instructions of one profiled kernel by SASS opcode and by issue pipe.

usage: ncu_by_opcode.py REPORT.ncu-rep KERNEL_SUBSTR

Pipe classes follow /opt/skills/guides/B300_MICROARCH.md ("fma vs alu split"): FFMA/FMUL/FADD/IMAD/
HFMA2 on the fma pipe; IADD3/LOP3/SHF/PRMT/FMNMX/SEL/ISETP/FSETP/LEA/MOV/VIADD... on the half-rate alu pipe.
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

FMA = {"FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "DFMA", "DMUL", "DADD"}
LSU = {"LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL", "ATOMG", "ATOMS", "RED", "LDC", "LDCU", "ULDC"}
XU = {"MUFU", "I2F", "F2I", "I2FP", "F2F", "POPC", "FLO", "BREV", "F2FP"}
CTRL = {"BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "WARPSYNC", "NOP", "BAR", "BMOV", "YIELD", "BREAK", "BRX", "JMP"}
WARP = {"SHFL", "VOTE", "MATCH", "VOTEU", "REDUX", "S2R", "S2UR", "CS2R", "R2UR", "R2P", "P2R"}


def pipe(op):
    base = op.split(".")[0]
    if base in FMA:
        return "fma"
    if base in LSU:
        return "lsu"
    if base in XU:
        return "xu"
    if base in CTRL:
        return "ctrl"
    if base in WARP:
        return "warp/sreg"
    if base.startswith("U") and base not in ("UNPACK",):
        return "uniform"
    return "alu"


def main():
    rep, kern = sys.argv[1:3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    sect = next((i for i in starts if kern in rows[i][1]), None)
    hdr = rows[sect + 1]
    end = min([i for i in starts if i > sect] + [len(rows)])
    body = [r for r in rows[sect + 2:end] if len(r) == len(hdr)]
    col = {h: i for i, h in enumerate(hdr)}
    by_op, by_pipe = defaultdict(lambda: [0, 0]), defaultdict(lambda: [0, 0])
    tot = 0
    for r in body:
        src = r[col["Source"]].strip()
        src = re.sub(r"^@!?U?P\d+\s+", "", src)
        op = src.split()[0].rstrip(";") if src else "?"
        ie = int(float(r[col["Instructions Executed"]] or 0))
        te = int(float(r[col["Predicated-On Thread Instructions Executed"]] or 0))
        b = op.split(".")[0]
        by_op[b][0] += ie
        by_op[b][1] += te
        by_pipe[pipe(op)][0] += ie
        by_pipe[pipe(op)][1] += te
        tot += ie
    print(f"# {rows[sect][1]}: {tot:,} warp instructions")
    print("pipe          warp-instr%   avg-thr")
    for k, v in sorted(by_pipe.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:12s} {100 * v[0] / tot:10.2f}%  {v[1] / max(1, v[0]):8.2f}")
    print("opcode        warp-instr%   avg-thr  pipe")
    for k, v in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:32]:
        print(f"{k:12s} {100 * v[0] / tot:10.2f}%  {v[1] / max(1, v[0]):8.2f}  {pipe(k)}")


if __name__ == "__main__":
    sys.exit(main())
