"""GPU parity of the three pipeline stages on the reference's bundled data set: the host mirror
(ambivert_b200/host.py, one batched CUDA call per stage) must reproduce, key for key and in the same
dict order, what the reference Python + reference C produced (tests/golden/pipeline_bundled.json, made
by oracle/make_golden.py) -- including the md5s the reference's own tests assert
(ambivert/tests/test_ambivert.py:204-323, :344-369)."""
import hashlib
import io

import pytest

from util import golden

pytestmark = pytest.mark.gpu

md5 = lambda s: hashlib.md5(s.encode("ascii")).hexdigest()


@pytest.fixture(scope="module")
def G():
    return golden("pipeline_bundled.json")


def fresh(G):
    from ambivert_b200.host import AmpliconData
    amp = AmpliconData()
    for p in G["pairs"]:
        amp.data[p["key"]] = [(("r", p["fwd"], "I" * len(p["fwd"])), ("r", p["rev"], "I" * len(p["rev"])))] * p["count"]
    for t in G["targets"]:
        amp.reference_sequences[tuple(t["key"])] = t["seq"]
    return amp


def as_tuples(d):
    return {k: tuple(v) for k, v in d.items()}


def test_merge_match_align_bundled_data(G):
    amp = fresh(G)
    amp.merge_overlaps(threshold=0, minimum_overlap=10)
    want = G["merge_threshold0_overlap10"]
    assert amp.merged == want["merged"]          # same keys and strings (the fixture is key-sorted, so order is
    # checked through the order-preserving lists: unmergable, potential_variants)
    assert amp.unmergable == want["unmergable"]
    # test_AmpliconData_match_to_reference (:237-270): five parameterisations and their md5s
    for case in G["match"]:
        amp.reference = as_tuples(case["preseed"])
        amp.match_to_reference(**case["kwargs"])
        assert amp.reference == as_tuples(case["reference"]), case["kwargs"]
        assert md5(str(sorted(list(amp.reference.items())))) == case["md5"]
    # test_AmpliconData_align_to_reference (:271-297)
    for tag, ga in (("global", True), ("local", False)):
        amp.aligned, amp.location, amp.potential_variants = {}, {}, []
        amp.align_to_reference(global_align=ga)
        w = G["align"][tag]
        assert amp.aligned == as_tuples(w["aligned"])
        assert amp.location == as_tuples(w["location"])
        assert amp.potential_variants == w["potential_variants"]
        assert md5(str(sorted(list(amp.aligned.items())))) == w["md5_aligned"]
        assert md5(str(sorted(list(amp.location.items())))) == w["md5_location"]


def test_process_amplicon_data_defaults_and_hash_table(G):
    """test_process_amplicon_data (:313-323) + save/load_hash_table (:344-369)"""
    w = G["process_threshold50_overlap20"]
    amp = fresh(G)
    amp.merge_overlaps(threshold=50, minimum_overlap=20)
    amp.match_to_reference()
    amp.align_to_reference()
    assert amp.merged == w["merged"] and amp.unmergable == w["unmergable"]
    assert amp.reference == as_tuples(w["reference"])
    assert amp.aligned == as_tuples(w["aligned"])
    assert amp.location == as_tuples(w["location"])
    assert amp.potential_variants == w["potential_variants"]
    assert md5(str(sorted(amp.potential_variants))) == w["md5_sorted_potential_variants"]
    buf = io.BytesIO()
    buf.close = lambda: None
    amp.save_hash_table(buf)
    assert hashlib.sha1(buf.getvalue()).hexdigest() == w["hash_table_pickle_sha1"]     # byte-identical hash table
    # reload: cached keys are skipped by match_to_reference (ambivert.py:516-523)
    amp2 = fresh(G)
    amp2.merge_overlaps(threshold=50, minimum_overlap=20)
    amp2.load_hash_table(io.BytesIO(buf.getvalue()))
    cached = dict(amp2.reference)
    amp2.match_to_reference()
    assert amp2.reference == as_tuples(w["reference"]) and all(amp2.reference[k] == v for k, v in cached.items())


def test_wrappers_follow_reference_error_behaviour():
    from ambivert_b200 import host as H
    with pytest.raises(RuntimeError):
        H.smith_waterman("AC-GT", "ACGT")                 # test_ambivert.py:141-143
    with pytest.raises(ValueError):
        H.needleman_wunsch("", "ACGT")                    # NULL from the C call -> ctypes ValueError
    assert H.needleman_wunsch("ACGT", "acgt") == ("ACGT", "acgt", 0, 0)
    a1, a2, s1, s2 = H.smith_waterman("TTTTACGTACGTAC", "ggACGTACGTACgg")
    assert (a1, a2.upper(), s1, s2) == ("ACGTACGTAC", "ACGTACGTAC", 4, 2)


def test_read_pairs_to_vcf_bundled_data(G):
    """the whole chain on the reference's bundled data, every stage on the device: cluster (md5 keys) -> merge -> match ->
    align -> variant scan -> consolidation -> VCF text, against the literal VCF the reference's own test asserts
    (tests/test_ambivert.py:532-557, stored by oracle/make_golden_variants.py) -- process_amplicon_data(threshold=50, overlap=20)"""
    import gzip
    import json
    import os
    from ambivert_b200.host import AmpliconData
    with gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "variants.json.gz"), "rb") as f:
        V = json.loads(f.read().decode())["bundled"]["threshold50_overlap20"]
    amp = AmpliconData()
    pairs = []
    for p in G["pairs"]:
        pairs += [(("r", p["fwd"], "I" * len(p["fwd"])), ("r", p["rev"], "I" * len(p["rev"])))] * p["count"]
    amp.add_read_pairs(pairs)                                     # keys computed on the device
    assert list(amp.data) == [p["key"] for p in G["pairs"]] and {k: len(v) for k, v in amp.data.items()} == V["counts"]
    for t in G["targets"]:
        amp.reference_sequences[tuple(t["key"])] = t["seq"]
    amp.merge_overlaps(threshold=50, minimum_overlap=20)
    amp.match_to_reference()
    amp.align_to_reference()
    assert {k: list(v) for k, v in amp.aligned.items()} == V["aligned"]
    assert amp.potential_variants == V["potential_variants"]
    amp.call_amplicon_variants()
    assert str(sorted(amp.called_variants.items())) == V["str_sorted_called_variants"]
    amp.consolidate_variants()
    assert str(sorted(amp.consolidated_variants)) == V["str_sorted_consolidated"]
    out = io.StringIO()
    amp.print_consolidated_vcf(outfile=out)
    assert out.getvalue() == V["vcf"]
    assert "chr17\t41244435\t.\tT\tC\t.\tPASS\tDP=193;AC=122;AF=0.632;ALTAMPS=1186cf52bcce24d9e5446675e896b90c,268014b6f669772a4e97b99427563d6b;REFAMPS=c3ee13ce7e3e7973b37714187bd2019d\n" in out.getvalue()
