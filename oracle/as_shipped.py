#!/usr/bin/env python3
"""as_shipped.py -- the reference's OWN Python + ctypes + C-extension path, timed as a user runs it.
TEST / BASELINE INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg and tests/ run it in a subprocess).

It imports the UNMODIFIED reference package vendored under oracle/_ref/pkg_cpu (its C extension built from
lib/align/*.c) -- or, with --pkg, any other vendored copy, e.g. oracle/_ref/pkg_gpu with align_c.b200.so in
the extension slot -- and times AmpliconData.match_to_reference(threads=N): one ctypes call per
(amplicon, target) pair from a multiprocessing.dummy.Pool created per amplicon (ambivert/ambivert.py:461-485).

  --bundled            the package's own test data (tests/data: 2580 read pairs, 26 manifest targets):
                       process_twofile_readpairs + merge_overlaps + add_references_from_manifest, then the match
  --queries F --targets F   one sequence per line: merged amplicons x targets (a sample of bench.py's workload)

Prints ONE JSON line: seconds, cells, GCUPS, threads, and the chosen target per amplicon (index or -1).
"""
import argparse
import io
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pkg", default=os.path.join(HERE, "_ref", "pkg_cpu"))
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--bundled", action="store_true")
    ap.add_argument("--threshold", type=int, default=0)
    ap.add_argument("--queries")
    ap.add_argument("--targets")
    args = ap.parse_args()
    sys.path.insert(0, args.pkg)
    import logging
    logging.disable(logging.CRITICAL)
    import ambivert.align as align_mod
    from ambivert import ambivert
    amp = ambivert.AmpliconData()
    if args.bundled:
        data = os.path.join(args.pkg, "ambivert", "tests", "data")
        amp.process_twofile_readpairs(open(os.path.join(data, "testdata_R1.fastq"), "rb"),
                                      open(os.path.join(data, "testdata_R2.fastq"), "rb"))
        amp.add_references_from_manifest(io.TextIOWrapper(open(os.path.join(data, "testdatamanifest.txt"), "rb")))
        amp.merge_overlaps(threshold=args.threshold)
        what = "bundled test data, threshold %d" % args.threshold
    else:
        queries = [ln.strip() for ln in open(args.queries) if ln.strip()]
        targets = [ln.strip() for ln in open(args.targets) if ln.strip()]
        for i, q in enumerate(queries):
            amp.merged["q%07d" % i] = q
        for j, t in enumerate(targets):
            amp.reference_sequences[("t%06d" % j, "chr1", str(j), str(j + len(t)), "+")] = t
        what = "%d amplicons x %d targets" % (len(queries), len(targets))
    keys = list(amp.merged)
    ref_keys = list(amp.reference_sequences)
    cells = float(sum(len(amp.merged[k]) for k in keys)) * float(sum(len(amp.reference_sequences[r]) for r in ref_keys))
    t0 = time.perf_counter()
    amp.match_to_reference(threads=args.threads)
    sec = time.perf_counter() - t0
    index = {r: j for j, r in enumerate(ref_keys)}
    chosen = [index[amp.reference[k]] if k in amp.reference else -1 for k in keys]
    print(json.dumps({"what": what, "seconds": sec, "cells": cells, "gcups": cells / sec / 1e9, "threads": args.threads,
                      "amplicons": len(keys), "targets": len(ref_keys), "chosen": chosen,
                      "extension": os.path.basename(align_mod.align_ctypes.align_c._name) if hasattr(align_mod, "align_ctypes") else align_mod.so_filename,
                      "api": "ambivert.ambivert.AmpliconData.match_to_reference(threads=%d): ctypes + multiprocessing.dummy.Pool per amplicon" % args.threads}))


if __name__ == "__main__":
    main()
