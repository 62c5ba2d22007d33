#!/usr/bin/env python3
"""tools/probe_variants.py -- variant scan of n aligned amplicons: device + end-to-end time of
abv_call_variants next to the oracle restatement of the reference's Python loop on a bounded sample."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import variants_oracle as VO
from ambivert_b200 import batch as B
from ambivert_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
base = synth.aligned_rows(5000, 0xF4)
rows = [base[i % len(base)] for i in range(n)]
s = B.pack([r[0].encode() for r in rows]); r = B.pack([x[1].encode() for x in rows])
B.call_variants(s, r)
best = None
for _ in range(5):
    t0 = time.perf_counter(); status, start, count, recs = B.call_variants(s, r); t1 = time.perf_counter()
    tm = B.timing()
    if best is None or t1 - t0 < best[0]:
        best = (t1 - t0, tm)
t0 = time.perf_counter()
for a, b in base[:2000]:
    VO.scan(a, b)
py = (time.perf_counter() - t0) / 2000
row_bytes = float(s[1][-1] + r[1][-1])
print(json.dumps({"workload": "variant scan of %d aligned amplicons (%.0f sites each)" % (n, row_bytes / 2 / n),
                  "e2e_s": best[0], "h2d_ms": best[1]["h2d_ms"], "device_ms": best[1]["kernel_ms"], "d2h_ms": best[1]["d2h_ms"],
                  "records": int(len(recs)), "row_bytes": row_bytes,
                  "device_GBps_two_passes": 2 * row_bytes / best[1]["kernel_ms"] / 1e6,
                  "amplicons_per_s_e2e": n / best[0], "python_port_amplicons_per_s": 1 / py, "python_sample": 2000}))
