#!/usr/bin/env python3
"""oracle/make_golden_panel.py -- medium-size golden results from the REAL reference C code.

Run in the build container only (needs oracle/_ref built from /root/reference/lib/align):
    python oracle/make_golden_panel.py
Writes tests/golden/panel_results.npz:
  * config-3-shaped panel: 300 manifest-style targets x 4,000 unique amplicons from
    ambivert_b200.synth (seeded) -> per-query best hit (score, target) of the reference's global_align_raw
    under the selection rule of ambivert.py:478-485 (seed -1000000) and of glocal_align_raw (seed 0);
  * config-2-shaped read pairs: 20,000 F/R pairs -> local_align_raw score + fragment list of every pair
    (what merge_overlaps consumes), stored as arrays.
The inputs are not stored: the test regenerates them from the same seeds and checks their sha256.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from ambivert_b200 import synth  # noqa: E402

PANEL = dict(n_targets=300, n_queries=4000, target_seed=0xB2CA, query_seed=0xB2CB)
PAIRS = dict(n_pairs=20000, seed=0xA3B1)


def panel_inputs():
    targets = synth.panel_targets(PANEL["n_targets"], 170, 250, seed=PANEL["target_seed"])
    queries, _ = synth.panel_queries(targets, PANEL["n_queries"], seed=PANEL["query_seed"])
    return queries, [t.upper() for t in targets]


def pair_inputs():
    """(forward read, reverse read brought to the forward strand) -- what merge_overlaps hands to the
    local aligner after its reverse_complement (ambivert.py:321-323)"""
    fwd, rev = synth.read_pairs(PAIRS["n_pairs"], seed=PAIRS["seed"])
    return fwd, [synth.reverse_complement(r) for r in rev]


def digest(seqs):
    h = hashlib.sha256()
    for s in seqs:
        h.update(s); h.update(b"\n")
    return h.hexdigest()


def main():
    O.build(ref=True)
    port, ref = O.port(), O.ref()
    threads = os.cpu_count() or 1
    out = {}
    queries, targets = panel_inputs()
    out["panel_sha"] = np.frombuffer((digest(queries) + digest(targets)).encode(), np.uint8)
    q = [O.encode(s) for s in queries]
    t = [O.encode(s) for s in targets]
    for name, mode, seed in (("global", O.GLOBAL, -1000000), ("glocal", O.GLOCAL, 0)):
        sec, bs, bi = port.timed_best_hits(mode, q, t, 5, O.DNA_SCORE, -7, -1, seed, threads, fns=ref.fns(mode))
        print(name, "%.1f s" % sec, flush=True)
        out["panel_%s_score" % name] = bs
        out["panel_%s_idx" % name] = bi
    fwd, rev = pair_inputs()
    out["pairs_sha"] = np.frombuffer((digest(fwd) + digest(rev)).encode(), np.uint8)
    scores, counts, frags = [], [], []
    for a, b in zip(fwd, rev):
        sc, fr = ref.align(O.LOCAL, a, b)
        scores.append(sc); counts.append(len(fr)); frags.extend(fr)
    out["pairs_score"] = np.asarray(scores, np.int32)
    out["pairs_count"] = np.asarray(counts, np.int32)
    out["pairs_frags"] = np.asarray(frags, np.int32).reshape(-1, 4)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "panel_results.npz"), **out)
    print("wrote panel_results.npz")


if __name__ == "__main__":
    main()
