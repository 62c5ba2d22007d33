#!/usr/bin/env python3
"""tools/probe_legacy_latency.py -- cost of ONE call through the reference's own ABI (global_align etc.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ambivert_b200.align as A
import random
rng = random.Random(0)
q = "".join(rng.choice("ACGT") for _ in range(240)).encode()
t = "".join(rng.choice("ACGT") for _ in range(230)).encode()
for name, fn in (("global_align", A.global_align), ("local_align", A.local_align)):
    al = fn(q, len(q), t, len(t), 5, A.DNA_MAP[1], A.DNA_SCORE, -7, -1); A.alignment_free(al)
    n = 300
    t0 = time.perf_counter()
    for _ in range(n):
        al = fn(q, len(q), t, len(t), 5, A.DNA_MAP[1], A.DNA_SCORE, -7, -1)
        A.alignment_free(al)
    dt = (time.perf_counter() - t0) / n
    print("%s 240x230: %.1f us per call (%.3f GCUPS)" % (name, dt * 1e6, 240 * 230 / dt / 1e9))

# the reference's own usage pattern: a thread pool mapping one ctypes call per target (ambivert.py:474-476)
from multiprocessing.dummy import Pool
targets = ["".join(rng.choice("ACGT") for _ in range(rng.randint(200, 260))).encode() for _ in range(300)]


def score(tt):
    al = A.global_align(q, len(q), tt, len(tt), 5, A.DNA_MAP[1], A.DNA_SCORE, -7, -1)
    s = al[0].score
    A.alignment_free(al)
    return s


for threads in (1, 2, 4, 8, 16):
    pool = Pool(threads)
    pool.map(score, targets[:32])
    t0 = time.perf_counter()
    res = pool.map(score, targets)
    dt = time.perf_counter() - t0
    pool.close()
    cells = sum(len(q) * len(tt) for tt in targets)
    print("Pool(%2d).map over 300 targets: %.1f ms, %.1f us per call, %.3f GCUPS" % (threads, dt * 1e3, dt / len(targets) * 1e6, cells / dt / 1e9))
