#!/usr/bin/env python3
"""tools/probe_pipeline_profile.py -- cProfile of the host mirror's stages (host.AmpliconData) at config-4 size: where the Python
time of add_read_pairs / merge_overlaps / align_to_reference goes.  Run on a GPU box."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ambivert_b200 import synth  # noqa: E402
from ambivert_b200.host import AmpliconData  # noqa: E402

n_unique = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
refs, pairs = synth.panel_pipeline_inputs(n_targets=2000, n_unique=n_unique, seed=0xC4A)
recs = [(("f%d" % i, f, "I" * len(f)), ("r%d" % i, r, "I" * len(r))) for i, (f, r, _c) in enumerate(pairs)]


def run(profile):
    amp = AmpliconData()
    for k, sq in refs:
        amp.reference_sequences[k] = sq
    for name, fn in (("add_read_pairs", lambda: amp.add_read_pairs(recs)), ("merge_overlaps", lambda: amp.merge_overlaps(0, 20)),
                     ("match_to_reference", lambda: amp.match_to_reference()), ("align_to_reference", lambda: amp.align_to_reference())):
        if profile and name != "match_to_reference":
            pr = cProfile.Profile()
            pr.enable()
            fn()
            pr.disable()
            print("=====", name)
            pstats.Stats(pr).sort_stats("tottime").print_stats(10)
        else:
            fn()


run(False)
run(True)
