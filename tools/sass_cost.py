#!/usr/bin/env python3
"""tools/sass_cost.py -- static issue-cost model of a kernel's steady-state loop, from its SASS (no GPU needed).

The streaming scorer is bound by the issue port of the SM sub-partition, not by one pipe (profiles/r02_issue_sensitivity.md:
extra FMA-pipe instructions cost as much as their share of ALL issue cycles).  Fitted on the A/B builds of round 2:
    cost = 1 per instruction  +  1 more per half-rate ALU / DPX instruction  +  1 per register-bank conflict
with two register banks (even / odd register numbers) and a conflict = every non-`.reuse` source register beyond the first
that a bank has to deliver to one instruction.  cost / (cell pairs of the loop) tracks the measured TCUPS of the builds.

    python tools/sass_cost.py [--so lib.so | --cubin file.cubin] 'mm16g_kernel<2, 8, -6>' [more kernel substrings ...]
"""
import argparse
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_report as S  # noqa: E402

ALU2 = ("VIMNMX3", "VIADDMNMX", "VIMNMX", "SEL", "LOP3", "ISETP", "IADD3", "LEA", "SHF", "PRMT", "VIADD", "PLOP3", "MOV", "IABS", "VABSDIFF")


def steady_loop(ins):
    index = {a: i for i, (a, _) in enumerate(ins)}
    regions = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.[A-Z.]+)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m and t.startswith("@") and int(m.group(1), 16) <= a and int(m.group(1), 16) in index:
            regions.append((index[int(m.group(1), 16)], i))
    dpx = lambda blk: sum(1 for _, t in blk if S.opcode(t).startswith(S.ALU_HALF))
    loops = []
    for lo, hi in regions:
        inner = [r for r in regions if r != (lo, hi) and r[0] >= lo and r[1] <= hi and dpx(ins[r[0]:r[1] + 1]) > 0]
        if not inner:
            loops.append(ins[lo:hi + 1])
    return max(loops, key=dpx) if loops else []


def cost(blk):
    n = len(blk)
    alu = conflicts = mov = 0
    per = collections.Counter()
    for _, t in blk:
        body = t.split(None, 1)[1] if t.startswith("@") else t
        op = body.split()[0]
        base = op.split(".")[0]
        if base in ALU2 and not op.startswith("IMAD"):
            alu += 1
        if op.startswith(("IMAD.MOV", "MOV")):
            mov += 1
        ops = [x.strip() for x in body.split(None, 1)[1].split(",")] if " " in body else []
        regs = [(int(m.group(1)), bool(m.group(2))) for x in ops[1:] for m in [re.match(r"-?R(\d+)(\.reuse)?$", x)] if m]
        if op.startswith(("LDS", "STS", "LDG", "STG", "SHFL", "BRA", "S2R")):
            continue
        c = collections.Counter(r % 2 for r, ru in regs if not ru)
        k = sum(v - 1 for v in c.values() if v > 1)
        conflicts += k
        if k:
            per[base] += k
    dpx = sum(1 for _, t in blk if S.opcode(t).startswith(("VIMNMX3",)))
    return {"instr": n, "alu": alu, "mov": mov, "conflicts": conflicts, "by_op": dict(per), "cost": n + alu + conflicts, "cell_pairs": dpx,
            "cost_per_cell_pair": round((n + alu + conflicts) / max(dpx, 1), 3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernels", nargs="+")
    ap.add_argument("--so", default=S.SO)
    args = ap.parse_args()
    fns = S.functions(args.so)
    for want in args.kernels:
        for d, n, b in fns:
            if want in d:
                c = cost(steady_loop(S.instructions(b)))
                print("%-44s %s" % (d.split("(")[0].replace("void abv::", ""), c))


if __name__ == "__main__":
    main()
