import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ambivert_b200 import batch as B, synth
n = 1000000
a, b = synth.read_pairs_packed(n)
pin = lambda x: torch.from_numpy(x).pin_memory().numpy()
a, b = (pin(a[0]), a[1]), (pin(b[0]), b[1])
for rep in range(3):
    print("---- rep", rep, file=sys.stderr, flush=True)
    t0 = time.perf_counter(); B.align_pairs(a, b, B.LOCAL); t1 = time.perf_counter()
    print("align_pairs %.1f ms" % (1e3 * (t1 - t0)), B.timing(), file=sys.stderr, flush=True)
