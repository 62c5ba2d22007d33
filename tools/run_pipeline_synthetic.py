#!/usr/bin/env python3
"""tools/run_pipeline_synthetic.py -- the three on-path pipeline stages end to end on a synthetic
TruSeq-style panel (BASELINE.json configs[2] shape: ~300 targets, ~20k unique read pairs), through the host
mirror of AmpliconData (ambivert_b200/host.py).  Prints stage times and sanity counts as one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ambivert_b200 import synth  # noqa: E402
from ambivert_b200.host import AmpliconData, reverse_complement  # noqa: E402


def main(n_targets=300, n_unique=20000, read_len=150, seed=0xB2CA):
    rng = np.random.default_rng(seed)
    targets = synth.panel_targets(n_targets, 170, 250, seed=seed)
    target_strs = [t.decode() for t in targets]
    amp = AmpliconData()
    for i, t in enumerate(target_strs):
        strand = "+" if i % 2 == 0 else "-"
        amp.reference_sequences[("target%03d" % i, "chr%d" % (1 + i % 2), str(1000 * i), str(1000 * i + len(t)), strand)] = t
    ref_keys = list(amp.reference_sequences)
    queries, src = synth.panel_queries(targets, n_unique, seed=seed + 1, p_unrelated=0.0)
    truth = {}
    for q, s in zip(queries, src):
        q = q.decode()                                 # stored targets and their amplicons are on the same strand
        if len(q) < read_len:
            continue
        f, r = q[:read_len], reverse_complement(q[-read_len:])
        count = int(rng.integers(1, 4))
        for _ in range(count):
            amp.add_reads("f", f, "I" * read_len, "r", r, "I" * read_len)
        truth[list(amp.data)[-1]] = ref_keys[s]
    out = {"targets": n_targets, "unique_pairs": len(amp.data), "read_pairs": amp.readpairs}
    t0 = time.perf_counter(); amp.merge_overlaps(threshold=0, minimum_overlap=20); t1 = time.perf_counter()
    amp.match_to_reference(); t2 = time.perf_counter()
    amp.align_to_reference(); t3 = time.perf_counter()
    out.update(merge_s=t1 - t0, match_s=t2 - t1, align_s=t3 - t2, merged=len(amp.merged), unmergable=len(amp.unmergable),
               matched=len(amp.reference), aligned=len(amp.aligned), potential_variants=len(amp.potential_variants))
    # minus-strand entries are flipped (amplicon and target) only for the final alignment (ambivert.py:559-565)
    out["matched_to_source"] = sum(1 for k, v in amp.reference.items() if truth.get(k) == v)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
