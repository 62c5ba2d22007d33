#!/usr/bin/env python3
"""tools/probe_e2e.py -- where does the end-to-end time of a hashtable build go (host vs device)?"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ambivert_b200 import batch as B, synth

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
targets = synth.panel_targets(2000)
queries, _ = synth.panel_queries(targets, nq)
targets = [t.upper() for t in targets]
q = B.pack(queries); t = B.pack(targets)
pin = lambda a: torch.from_numpy(a).pin_memory()
pq, pqo, pt, pto = pin(q[0]), pin(q[1]), pin(t[0]), pin(t[1])
torch.cuda.synchronize()
for rep in range(5):
    t0 = time.perf_counter()
    job = B.BestHitsJob((pq.numpy(), pqo.numpy()), (pt.numpy(), pto.numpy()))
    t1 = time.perf_counter()
    job.launch()
    t2 = time.perf_counter()
    bs, bi = job.fetch()
    t3 = time.perf_counter()
    job.close()
    t4 = time.perf_counter()
    print("prepare %.1f ms  launch(call) %.1f ms  fetch(sync+d2h) %.1f ms  close %.1f ms  total %.1f ms" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t4 - t0)), flush=True)
t0 = time.perf_counter(); B.best_hits((pq.numpy(), pqo.numpy()), (pt.numpy(), pto.numpy())); print("one-call best_hits %.1f ms" % (1e3 * (time.perf_counter() - t0)), B.timing())
