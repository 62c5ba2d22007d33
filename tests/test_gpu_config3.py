"""GPU parity at BASELINE.json configs[2] scale: the three on-path stages of the pipeline on a synthetic
TruSeq-style panel (300 targets, ~20k unique read pairs) against the REFERENCE PIPELINE (Python + C) run in
the build container (tests/golden/config3_pipeline.json, made by oracle/make_golden_config3.py).  The digests
are the reference tests' own kind (md5 of repr of the sorted state, test_ambivert.py:237-323); identical
merged reads, hash table, aligned rows, locations and potential variants mean an identical VCF, which the
reference derives from them alone."""
import hashlib
import io
import json
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config3_pipeline.json")


def test_config3_pipeline_states_match_reference():
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/config3_pipeline.json not generated")
    import make_golden_config3 as M
    from ambivert_b200 import synth
    from ambivert_b200.host import AmpliconData
    with open(GOLD) as f:
        G = json.load(f)
    refs, pairs = synth.panel_pipeline_inputs()
    assert M.inputs_sha(refs, pairs) == G["inputs_sha256"], "synthetic inputs drifted from the golden run"
    amp = AmpliconData()
    for k, s in refs:
        amp.reference_sequences[k] = s
    for f, r, c in pairs:
        for _ in range(c):
            amp.add_reads("f", f, "I" * len(f), "r", r, "I" * len(r))
    assert len(amp.data) == G["n_unique_pairs"] and amp.readpairs == G["read_pairs"]
    amp.merge_overlaps(threshold=0, minimum_overlap=20)
    amp.match_to_reference(threads=8)
    amp.align_to_reference()
    got = M.state_digests(amp)
    for key in ("n_merged", "n_unmergable", "n_reference", "n_aligned", "n_potential_variants", "md5_merged", "md5_reference",
                "md5_aligned", "md5_location", "md5_potential_variants", "unmergable"):
        assert got[key] == G[key], key
    buf = io.BytesIO()
    buf.close = lambda: None
    amp.save_hash_table(buf)
    assert hashlib.sha256(buf.getvalue()).hexdigest() == G["hash_table_pickle_sha256"]      # byte-identical hash table
