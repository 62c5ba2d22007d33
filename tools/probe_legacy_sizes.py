#!/usr/bin/env python3
"""tools/probe_legacy_sizes.py -- latency of one call through the reference's own ABI against the two lengths"""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ambivert_b200.align as A
rng = random.Random(0)
for la, lb in ((8, 8), (32, 240), (64, 240), (128, 240), (256, 240), (240, 32), (240, 64), (240, 128), (512, 240), (240, 512)):
    q = "".join(rng.choice("ACGT") for _ in range(la)).encode()
    t = "".join(rng.choice("ACGT") for _ in range(lb)).encode()
    al = A.global_align(q, len(q), t, len(t), 5, A.DNA_MAP[1], A.DNA_SCORE, -7, -1); A.alignment_free(al)
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        al = A.global_align(q, len(q), t, len(t), 5, A.DNA_MAP[1], A.DNA_SCORE, -7, -1)
        A.alignment_free(al)
    dt = (time.perf_counter() - t0) / n
    print("global_align %dx%d: %.1f us per call" % (la, lb, dt * 1e6))
