#!/usr/bin/env python3
"""tools/ncu_source_table.py -- source-level table of one kernel from an `ncu --set full --import-source on` report.

    ncu -i report.ncu-rep --page source --csv --launch-skip K --launch-count 1 > src.csv
    python tools/ncu_source_table.py src.csv

Prints where the issued warp instructions and the stall samples of the kernel go: by opcode, by stall reason, and by code
region (runs of SASS instructions executed equally often = the loops of the kernel)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[0][1] if rows and len(rows[0]) > 1 else "")
    hdr, data = rows[h], []
    for r in rows[h + 1:]:                     # the first kernel of the file only
        if r and r[0] in ("Kernel Name", "Address"):
            break
        data.append(r)
    ix = {name: i for i, name in enumerate(hdr)}
    tot, by_op, samples, stall, recs = 0, collections.Counter(), collections.Counter(), collections.Counter(), []
    for r in data:
        if len(r) < len(hdr):
            continue
        src = r[ix["Source"]].strip()
        body = src.split(None, 1)[1] if src.startswith("@") else src
        op = body.split()[0]
        ex, sm = int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)
        tot += ex
        parts = op.split(".")
        key = ".".join(parts[:2]) if parts[0] in ("IMAD", "LDS", "SHFL", "VIMNMX3", "VIADDMNMX", "VIMNMX") else parts[0]
        by_op[key] += ex
        samples[key] += sm
        for name in hdr:
            if name.startswith("stall_") and "Not Issued" not in name:
                stall[name] += int(r[ix[name]] or 0)
        recs.append((int(r[ix["Address"]], 16), src, ex, sm))
    ts = sum(samples.values()) or 1
    print("warp instructions executed: %d   stall samples: %d" % (tot, ts))
    print("\nby opcode (share of executed warp instructions, share of stall samples):")
    for k, v in by_op.most_common(18):
        print("  %-18s %6.2f %%   %6.2f %%" % (k, 100.0 * v / tot, 100.0 * samples[k] / ts))
    tt = sum(stall.values()) or 1
    print("\nby stall reason (all samples): " + ", ".join("%s %.1f %%" % (k[6:], 100.0 * v / tt) for k, v in stall.most_common(9)))
    base, runs, cur = recs[0][0], [], None
    for a, src, ex, sm in recs:
        if cur and cur["ex"] == ex:
            cur["n"] += 1; cur["sm"] += sm; cur["end"] = a
            cur["ops"][src.split()[0] if not src.startswith("@") else src.split()[1]] += 1
        else:
            cur = {"start": a, "end": a, "ex": ex, "n": 1, "sm": sm, "ops": collections.Counter()}
            runs.append(cur)
    print("\nby code region (runs of instructions executed equally often; regions below 0.5 % of the instructions omitted):")
    print("  %-19s %6s %14s %9s %9s   %s" % ("SASS offsets", "instr", "executed each", "share", "samples", "DPX / SEL / SHFL in the run"))
    for r in runs:
        if r["ex"] * r["n"] < 0.005 * tot:
            continue
        dpx = sum(v for k, v in r["ops"].items() if k.startswith(("VIMNMX", "VIADDMNMX")))
        sel = sum(v for k, v in r["ops"].items() if k.startswith("SEL"))
        shf = sum(v for k, v in r["ops"].items() if k.startswith("SHFL"))
        print("  0x%05x..0x%05x %7d %14d %8.2f%% %8.2f%%   %d / %d / %d" % (r["start"] - base, r["end"] - base, r["n"], r["ex"], 100.0 * r["ex"] * r["n"] / tot,
                                                                           100.0 * r["sm"] / ts, dpx, sel, shf))


if __name__ == "__main__":
    main(sys.argv[1])
