#!/usr/bin/env python3
"""tools/probe_merge_e2e.py -- end-to-end time of the one-to-one stage APIs on n read pairs with pinned host buffers"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ambivert_b200 import batch as B, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
a, b = synth.read_pairs_packed(n)
pin = lambda x: torch.from_numpy(x).pin_memory().numpy()
a, b = (pin(a[0]), a[1]), (pin(b[0]), b[1])
out = B.pinned_empty(a[0].size + b[0].size + 16)
B.align_pairs((a[0][:150000], a[1][:1001]), (b[0][:150000], b[1][:1001]), B.LOCAL)
bufs = B.align_pairs_buffers(n, 4 * n)
for rep in range(4):
    t0 = time.perf_counter(); sc, st, cn, fr = B.align_pairs(a, b, B.LOCAL); t1 = time.perf_counter(); tm = B.timing()
    t2 = time.perf_counter(); r = B.merge_pairs(a, b, 20, out=out); t3 = time.perf_counter(); tm2 = B.timing()
    t4 = time.perf_counter(); B.align_pairs(a, b, B.LOCAL, out=bufs); t5 = time.perf_counter(); tm3 = B.timing()
    print("align_pairs(out=pinned) %.1f ms (kernel %.1f d2h %.1f)" % (1e3 * (t5 - t4), tm3["kernel_ms"], tm3["d2h_ms"]), flush=True)
    print("align_pairs %.1f ms (h2d+plan %.1f kernel %.1f d2h %.1f) | merge_pairs %.1f ms (h2d+plan %.1f kernel %.1f d2h %.1f)" % (
        1e3 * (t1 - t0), tm["h2d_ms"], tm["kernel_ms"], tm["d2h_ms"], 1e3 * (t3 - t2), tm2["h2d_ms"], tm2["kernel_ms"], tm2["d2h_ms"]), flush=True)
