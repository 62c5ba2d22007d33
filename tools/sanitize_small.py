"""tools/sanitize_small.py -- a small many-to-many job per mode through the streaming kernels (global: several level
staircases per warp; glocal: deferred last rows) plus a small merge batch, compared with the 32-bit kernel.  Small enough
to run under compute-sanitizer (memcheck / racecheck) where that tool is available; the GPU pool of this project has it
closed, so the CPU emulation of the kernels (tests/emu, bounds-checked row reads) is what guards the pointer bookkeeping."""
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ambivert_b200 import batch as B  # noqa: E402

rng = random.Random(9)
dna = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
targets = [dna(rng.randint(32, 70)) for _ in range(700)]
queries = [dna(n) for n in (20, 60, 70, 100, 150, 200, 230, 280)]
q = [s.encode() for s in queries]
t = [s.encode() for s in targets]
for mode in (B.GLOBAL, B.GLOCAL, B.OVERLAP):          # overlap: the packed form of mm16_kernel
    got = B.score_matrix(q, t, mode)
    B.set_option("path", B.PATH_WF32)
    want = B.score_matrix(q, t, mode)
    B.set_option("path", B.PATH_AUTO)
    assert (got == want).all(), mode
# several strip widths and both instantiations per width: the packed launches chained by tail flags on their own streams
queries2 = [dna(n) for n in (200, 217, 218, 224, 225, 240, 248, 249, 256, 257, 260) for _ in range(3)]
hits = B.best_hits([s.encode() for s in queries2], t)
B.set_option("path", B.PATH_WF32)
hits32 = B.best_hits([s.encode() for s in queries2], t)
B.set_option("path", B.PATH_AUTO)
assert (hits[0] == hits32[0]).all() and (hits[1] == hits32[1]).all()
fwd = [dna(150) for _ in range(64)]
rev = [f[rng.randint(40, 120):] + dna(rng.randint(10, 60)) for f in fwd]
sc, st, cn, fr = B.align_pairs([s.encode() for s in fwd], [s.encode() for s in rev], B.LOCAL)
print("sanitize_small ok", int(np.sum(sc)))
