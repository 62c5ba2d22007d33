#!/usr/bin/env python3
"""tools/gpu_probe.py -- quick on-box measurements: integer-pipe microbenchmarks and first
throughput numbers of the two kernels.  Writes gpurun_out/probe.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ambivert_b200 import batch as B  # noqa: E402

NAMES = ["iadd32", "vimnmx3_s16x2", "viaddmnmx_s16x2", "viadd_16x2", "packed_cell_mix", "cell32_mix", "vimnmx3_s32", "packed_cell_iadd32_mix"]


def synth(n_t, n_q, seed=0xC4):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    tl = rng.integers(200, 261, n_t)
    targets = [bytes(rng.choice(acgt, int(L))) for L in tl]
    src = rng.integers(0, n_t, n_q)
    queries = []
    for i in range(n_q):
        s = bytearray(targets[src[i]])
        for _ in range(int(rng.poisson(1.5))):
            s[int(rng.integers(0, len(s)))] = b"ACGT"[int(rng.integers(0, 4))]
        queries.append(bytes(s))
    return queries, targets


def main():
    out = {"version": B.version(), "int_peak": {}}
    for w, name in enumerate(NAMES):
        ops, mhz = B.int_peak(w, 20000)
        out["int_peak"][name] = {"lane_ops_per_s": ops, "per_sm_per_clk_at_nominal": ops / 148 / (mhz * 1e6), "nominal_mhz": mhz}
        print(name, "%.3e lane-ops/s" % ops, "%.1f /SM/clk@%.0fMHz" % (ops / 148 / (mhz * 1e6), mhz), flush=True)
    for mode, mname in ((B.GLOBAL, "global"), (B.GLOCAL, "glocal"), (B.LOCAL, "local")):
        q, t = synth(2000, 6000)
        B.best_hits(q[:500], t, mode)
        t0 = time.time()
        bs, bi = B.best_hits(q, t, mode)
        wall = time.time() - t0
        tm = B.timing()
        out["best_hits_" + mname] = {"n_q": len(q), "n_t": len(t), "cells": tm["cells"], "kernel_ms": tm["kernel_ms"], "h2d_ms": tm["h2d_ms"],
                                    "wall_s": wall, "gcups_kernel": tm["cells"] / tm["kernel_ms"] / 1e6, "gcups_e2e": tm["cells"] / wall / 1e9}
        print(mname, out["best_hits_" + mname], flush=True)
    B.set_option("path", B.PATH_WF32)
    q, t = synth(2000, 300)
    B.best_hits(q, t, B.GLOBAL)
    tm = B.timing()
    out["best_hits_global_wf32"] = {"cells": tm["cells"], "kernel_ms": tm["kernel_ms"], "gcups_kernel": tm["cells"] / tm["kernel_ms"] / 1e6}
    print("wf32", out["best_hits_global_wf32"], flush=True)
    B.set_option("path", B.PATH_AUTO)
    rng = np.random.default_rng(1)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    fw, rv = [], []
    for _ in range(100000):
        amp = bytes(rng.choice(acgt, int(rng.integers(170, 281))))
        fw.append(amp[:150]); rv.append(amp[-150:])
    for mode, mname in ((B.LOCAL, "local"), (B.GLOBAL, "global")):
        B.align_pairs(fw[:1000], rv[:1000], mode)
        t0 = time.time()
        sc, st, cn, fr = B.align_pairs(fw, rv, mode)
        wall = time.time() - t0
        tm = B.timing()
        out["align_pairs_" + mname] = {"n": len(fw), "cells": tm["cells"], "kernel_ms": tm["kernel_ms"], "wall_s": wall,
                                      "gcups_kernel": tm["cells"] / tm["kernel_ms"] / 1e6, "frags": int(fr.shape[0])}
        print("pairs", mname, out["align_pairs_" + mname], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
