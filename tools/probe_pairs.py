#!/usr/bin/env python3
"""tools/probe_pairs.py -- where does the time of a big one-to-one batch (merge stage) go?"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ambivert_b200 import batch as B, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
t0 = time.time(); fwd, rev = synth.read_pairs(n); print("gen", time.time() - t0, flush=True)
B.align_pairs(fwd[:1000], rev[:1000], B.LOCAL)
for rep in range(3):
    t0 = time.time(); a = B.pack(fwd); b = B.pack(rev); t1 = time.time()
    sc, st, cn, fr = B.align_pairs(a, b, B.LOCAL); t2 = time.time()
    tm = B.timing()
    print("pack %.3f call %.3f | h2d %.1f ms kernel %.1f ms d2h %.1f ms | frags %d | kernel GCUPS %.1f e2e GCUPS %.1f" % (
        t1 - t0, t2 - t1, tm["h2d_ms"], tm["kernel_ms"], tm["d2h_ms"], fr.shape[0], tm["cells"] / tm["kernel_ms"] / 1e6, tm["cells"] / (t2 - t1) / 1e9), flush=True)
