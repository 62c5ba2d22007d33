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


def test_config3_variants_after_the_path_match_reference():
    """the step after the path at config-3 scale (19.7k potential variants): device variant scan + host consolidation
    against the reference's call_variants / consolidate_variants / VCF on the same aligned rows, including the amplicons
    on which the reference's own assert fires (call_variants.py:139: a substitution right before the trailing primer)"""
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/config3_pipeline.json not generated")
    import make_golden_config3 as M
    from ambivert_b200 import synth
    from ambivert_b200.call_variants import call_variants_many
    from ambivert_b200.host import AmpliconData
    with open(GOLD) as f:
        G = json.load(f)
    if "md5_called_variants" not in G:
        pytest.skip("golden predates the variant fields")
    refs, pairs = synth.panel_pipeline_inputs()
    amp = AmpliconData()
    for k, s in refs:
        amp.reference_sequences[k] = s
    bulk = []
    for f, r, c in pairs:
        bulk += [(("f", f, "I" * len(f)), ("r", r, "I" * len(r)))] * c
    amp.add_read_pairs(bulk)                                       # md5 keys on the device
    assert len(amp.data) == G["n_unique_pairs"] and amp.readpairs == G["read_pairs"]
    amp.merge_overlaps(threshold=0, minimum_overlap=20)
    amp.match_to_reference(threads=8)
    amp.align_to_reference()
    assert M.state_digests(amp)["md5_aligned"] == G["md5_aligned"]
    keys = list(amp.potential_variants)
    jobs = [(amp.aligned[k][0], amp.aligned[k][1], amp.reference[k][1], int(amp.reference[k][2])) for k in keys]
    results = call_variants_many(jobs)
    called = [(k, str(r)) for k, r in zip(keys, results) if not isinstance(r, Exception)]
    errors = {k: type(r).__name__ for k, r in zip(keys, results) if isinstance(r, Exception)}
    assert len(called) == G["n_called"] and M.md5(str(sorted(called))) == G["md5_called_variants"]
    E = G["call_variants_errors"]
    assert len(errors) == E["count"] and sorted(set(errors.values())) == E["types"] and M.md5(str(sorted(errors.items()))) == E["md5_keys"]
    if E["count"]:                                                  # call_amplicon_variants stops where the reference stops
        with pytest.raises(AssertionError):
            amp.call_amplicon_variants()
        assert next(k for k in keys if k in errors) == E["first_in_pipeline_order"]
        assert list(amp.called_variants) == keys[:keys.index(E["first_in_pipeline_order"])]
    amp.called_variants = {k: r for k, r in zip(keys, results) if not isinstance(r, Exception)}
    amp.consolidated_variants = []
    amp.consolidate_variants()
    assert len(amp.consolidated_variants) == G["n_consolidated"]
    assert M.md5(str(sorted(amp.consolidated_variants))) == G["md5_consolidated"]
    assert M.md5(str(list(amp.get_filtered_variants()))) == G["md5_filtered"]
    out = io.StringIO()
    amp.print_consolidated_vcf(outfile=out)
    assert hashlib.sha256(out.getvalue().encode()).hexdigest() == G["vcf_of_callable_sha256"]
