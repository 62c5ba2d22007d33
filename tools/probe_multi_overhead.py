#!/usr/bin/env python3
"""tools/probe_multi_overhead.py -- does the in-process multi-device call (one worker thread per shard) cost anything over the single
call?  best_hits vs best_hits_multi with all shards on device 0 (10k x 2k panel).  Measured: 122.5 / 122.5 / 125.0 ms."""
import os
import sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ambivert_b200 import batch as B, synth
targets = synth.panel_targets(2000)
queries, _ = synth.panel_queries(targets, 10000)
targets = [t.upper() for t in targets]
q = B.pack(queries); t = B.pack(targets)
for rep in range(4):
    t0 = time.perf_counter(); a = B.best_hits(q, t); t1 = time.perf_counter()
    b = B.best_hits_multi(q, t, [0]); t2 = time.perf_counter()
    c = B.best_hits_multi(q, t, [0, 0]); t3 = time.perf_counter()
    print("best_hits %.1f ms | multi([0]) %.1f ms | multi([0,0]) %.1f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2)), (a[0]==b[0]).all() and (a[1]==c[1]).all(), flush=True)
