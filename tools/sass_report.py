#!/usr/bin/env python3
"""tools/sass_report.py -- SASS evidence of the shipped library (no GPU needed: cuobjdump reads the .so).

For one kernel of ambivert_b200/align/align_c.b200.so (selected by a substring of its demangled name, e.g.
'mm16g_kernel<2, 8>' = global, 8 columns per lane) it prints
  * the opcode histogram of the whole kernel,
  * its basic blocks ranked by size, the steady-state loop being the largest block that branches back to itself,
  * that loop's opcode histogram, pipe split (ALU/DPX vs FMA-pipe adds vs shared memory vs shuffles) and listing,
  * the staging sequence (UBLKCP bulk copy + SYNCS mbarrier) if the kernel has one.

    python tools/sass_report.py 'mm16g_kernel<2, 8>' > profiles/r02_sass_mm16g_global_k8_steady.txt
    python tools/sass_report.py --list                    # every kernel of the library with its size
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "ambivert_b200", "align", "align_c.b200.so")

INST = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\*")
ALU_HALF = ("VIMNMX3", "VIADDMNMX", "VIMNMX")          # DPX / integer min-max: the half-rate pipe that bounds the kernels


def functions(so):
    txt = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True, check=True).stdout
    out, name, body = [], None, []
    for ln in txt.splitlines():
        m = re.match(r"\s+Function : (\S+)", ln)
        if m:
            if name:
                out.append((name, body))
            name, body = m.group(1), []
        elif name:
            body.append(ln)
    if name:
        out.append((name, body))
    names = subprocess.run(["c++filt"] + [n for n, _ in out], stdout=subprocess.PIPE, text=True).stdout.splitlines()
    return [(d, n, b) for (n, b), d in zip(out, names)]


def instructions(body):
    ins = []
    for ln in body:
        m = INST.match(ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def opcode(text):
    t = text
    if t.startswith("@"):
        t = t.split(None, 1)[1]
    return t.split()[0].rstrip(";")


def blocks(ins):
    """basic blocks: leaders are branch targets and the instructions after a branch"""
    addr = [a for a, _ in ins]
    leaders = {addr[0]}
    for i, (a, t) in enumerate(ins):
        op = opcode(t)
        if op.startswith(("BRA", "BRX", "JMP", "EXIT", "RET", "CALL", "BSYNC", "BREAK")) or op.startswith("WARPSYNC"):
            if i + 1 < len(ins):
                leaders.add(addr[i + 1])
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and op.startswith("BRA"):
                leaders.add(int(m.group(1), 16))
    out, cur = [], []
    for a, t in ins:
        if a in leaders and cur:
            out.append(cur)
            cur = []
        cur.append((a, t))
    if cur:
        out.append(cur)
    return out


def histogram(ins):
    h = collections.Counter()
    for _, t in ins:
        op = opcode(t)
        parts = op.split(".")
        key = parts[0]
        if key in ("VIMNMX3", "VIADDMNMX", "VIMNMX", "IMAD", "LDS", "STS", "SHFL", "LOP3", "IADD3", "SHF", "LDG", "STG"):
            key = ".".join(parts[:3]) if key in ("VIMNMX3", "VIADDMNMX", "VIMNMX") else ".".join(parts[:2])
        h[key] += 1
    return h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernel", nargs="?", help="substring of the demangled kernel name")
    ap.add_argument("--so", default=SO)
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--listing", type=int, default=400, help="print at most this many instructions of the steady-state loop")
    args = ap.parse_args()
    fns = functions(args.so)
    if args.list or not args.kernel:
        for d, n, b in fns:
            print("%6d instructions  %s" % (len(instructions(b)), d))
        return
    sel = [(d, n, b) for d, n, b in fns if args.kernel in d]
    if not sel:
        sys.exit("no kernel matches %r" % args.kernel)
    d, n, b = sel[0]
    ins = instructions(b)
    print("library : %s" % os.path.relpath(args.so, ROOT))
    print("kernel  : %s\n          %s\n          %d SASS instructions, arch sm_100a (cuobjdump -sass)" % (d, n, len(ins)))
    h = histogram(ins)
    print("\n== whole kernel, opcode histogram ==")
    for k, v in h.most_common(40):
        print("  %-28s %6d" % (k, v))
    # loops = regions [target, branch] of BACKWARD branches (the unrolled loops start with a BRA.DIV of the warp shuffle,
    # so they are not single basic blocks); keep the innermost ones and take the one with the most DPX instructions
    index = {a: i for i, (a, _) in enumerate(ins)}
    regions = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.[A-Z.]+)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m and t.startswith("@") and opcode(t).startswith("BRA") and int(m.group(1), 16) <= a and int(m.group(1), 16) in index:   # loops close with a predicated branch
            regions.append((index[int(m.group(1), 16)], i))
    def dpx_of(blk):
        return sum(v for k, v in histogram(blk).items() if k.startswith(ALU_HALF))
    loops = []
    for lo, hi in regions:
        inner = [r for r in regions if r != (lo, hi) and r[0] >= lo and r[1] <= hi and dpx_of(ins[r[0]:r[1] + 1]) > 0]
        if not inner:
            loops.append(ins[lo:hi + 1])
    print("\n== innermost loops (backward-branch regions), by DPX/min-max instructions ==")
    for blk in sorted(loops, key=dpx_of, reverse=True)[:8]:
        print("  0x%05x..0x%05x  %5d instructions, %4d DPX/min-max" % (blk[0][0], blk[-1][0], len(blk), dpx_of(blk)))
    if not loops:
        print("  (none)")
        return
    hot = max(loops, key=dpx_of)
    hh = histogram(hot)
    print("\n== steady-state loop 0x%05x..0x%05x: %d instructions ==" % (hot[0][0], hot[-1][0], len(hot)))
    for k, v in hh.most_common():
        print("  %-28s %6d" % (k, v))
    dpx = sum(v for k, v in hh.items() if k.startswith(ALU_HALF))
    sel_ = sum(v for k, v in hh.items() if k.split(".")[0] in ("SEL", "ISETP", "LOP3", "SHF", "PRMT", "IADD3", "LEA", "MOV", "VIADD") and not k.startswith("IMAD"))
    fma = sum(v for k, v in hh.items() if k.startswith("IMAD"))
    mem = sum(v for k, v in hh.items() if k.startswith(("LDS", "STS", "LDG", "STG")))
    shf = sum(v for k, v in hh.items() if k.startswith("SHFL"))
    print("\n  half-rate ALU pipe (DPX / packed min-max): %d   other integer-ALU-pipe ops: %d   FMA-pipe (IMAD.*): %d   shared/global memory: %d   shuffles: %d   other: %d" % (
        dpx, sel_, fma, mem, shf, len(hot) - dpx - sel_ - fma - mem - shf))
    print("\n== listing of the steady-state loop (first %d instructions) ==" % args.listing)
    for a, t in hot[:args.listing]:
        print("  /*%05x*/  %s" % (a, t))
    stage = [(a, t) for a, t in ins if opcode(t).startswith(("UBLKCP", "SYNCS", "UTMA"))]
    if stage:
        print("\n== staging: bulk async copies (TMA 1-D) and mbarrier operations in the kernel ==")
        for a, t in stage[:40]:
            print("  /*%05x*/  %s" % (a, t))
        print("  ... %d in total (UBLKCP %d, SYNCS %d)" % (len(stage), sum(1 for _, t in stage if opcode(t).startswith("UBLKCP")),
                                                           sum(1 for _, t in stage if opcode(t).startswith("SYNCS"))))


if __name__ == "__main__":
    main()
