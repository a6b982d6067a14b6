#!/usr/bin/env python
"""Executed warp-instructions of one kernel of an .ncu-rep, grouped by opcode (and a rough issue-pipe class).
    python tools/ncu_opmix.py rep [launch]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blk = raw.split('"Kernel Name",')[1 + launch]
lines = blk.split("\n")
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
isrc, iex, ithr = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
FMA = {"FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "FFMA32I", "FMUL32I", "FADD32I"}
LSU = {"LDG", "STG", "LDL", "STL", "LDS", "STS", "ATOMG", "RED", "LDC", "LDCU"}
XU = {"MUFU", "I2F", "F2I", "F2F", "I2FP", "F2FP"}
CTRL = {"BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "CALL", "RET", "NOP", "BAR"}
ops = collections.Counter(); cls = collections.Counter(); thr = collections.Counter()
for r in rows[1:]:
    if len(r) <= ithr or not r[iex]: continue
    try: e = int(r[iex])
    except ValueError: continue
    toks = r[isrc].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    base = op.split(".")[0]
    ops[base] += e; thr[base] += e * float(r[ithr] or 0)
    cls["fma" if base in FMA else "lsu" if base in LSU else "xu" if base in XU else "ctrl" if base in CTRL else "alu"] += e
tot = sum(ops.values())
print("total warp-instr %.3f G" % (tot / 1e9))
print("by class: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in cls.most_common()))
for op, e in ops.most_common(32):
    print("  %-10s %6.2f%%  lanes %.1f" % (op, 100 * e / tot, thr[op] / e))
