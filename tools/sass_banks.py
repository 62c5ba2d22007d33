#!/usr/bin/env python3
"""tools/sass_banks.py -- register-bank pressure of the DPX instructions in a kernel's hottest basic block (a proxy for how
lucky ptxas' register allocation was: three-source DPX instructions whose non-reused sources share a bank (register number
mod 4) pay extra operand-fetch cycles).   python tools/sass_banks.py [lib.so ...]"""
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_report as S  # noqa: E402


def report(so, names):
    fns = S.functions(so)
    for want in names:
        for d, n, b in fns:
            if want not in d:
                continue
            ins = S.instructions(b)
            blk = max(S.blocks(ins), key=lambda k: sum(v for kk, v in S.histogram(k).items() if kk.startswith(S.ALU_HALF)))
            conf, reuse, ndpx = collections.Counter(), 0, 0
            for a, t in blk:
                if S.opcode(t).startswith(("VIMNMX3", "VIADDMNMX")):
                    ndpx += 1
                    srcs = re.findall(r"\bR(\d+)(\.reuse)?", t.split(None, 1)[1])[1:]
                    c = collections.Counter(int(r) % 4 for r, ru in srcs if not ru)
                    conf[max(c.values()) if c else 0] += 1
                    reuse += sum(1 for r, ru in srcs if ru)
            print("%-34s %-28s block %4d instr, %3d DPX: same-bank sources 2-way %3d, 3-way %2d, reuse flags %3d, MOV %2d" % (
                so.split("/")[-1], d.split("(")[0].replace("void abv::", ""), len(blk), ndpx, conf[2], conf[3], reuse,
                sum(1 for a, t in blk if S.opcode(t).startswith(("IMAD.MOV", "MOV")))))


if __name__ == "__main__":
    libs = sys.argv[1:] or [S.SO]
    for so in libs:
        report(so, ["mm16g_kernel<2, 7>", "mm16g_kernel<2, 8>", "mm16g_kernel<2, 9>", "mm16g_kernel<3, 7>", "mm16g_kernel<3, 8>"])
