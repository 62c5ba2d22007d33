"""CPU tests of the host-side glue in ambivert_b200/host.py that needs no GPU: the byte-level helpers
either side of the aligner, checked against what the reference Python produced
(tests/golden/merge_strings.json.gz, tests/golden/pipeline_bundled.json)."""
import gzip
import io
import json
import os
import pickle

import pytest

from util import golden

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "merge_strings.json.gz")


@pytest.fixture(scope="module")
def H():
    from ambivert_b200 import host
    return host


def test_flatten_and_assembly_reproduce_reference_merges(H):
    """given the reference's own gapped rows, the host helpers rebuild the reference's merged reads"""
    with gzip.open(GOLD, "rt") as f:
        G = json.load(f)
    n = 0
    for p in G["pairs"]:
        fa, ra, fs, rs = p["sw"]
        if len(fa) < G["minimum_overlap"]:
            assert p["merged"] is None
            continue
        merged = p["fwd"][:fs] + H.flatten_paired_alignment(fa, ra) + p["rev"][rs + len(ra.replace("-", "")):]
        assert merged == p["merged"]
        n += 1
    assert n > 500


def test_reverse_complement_and_iupac(H):
    assert H.reverse_complement("AACG") == "CGTT"
    assert H.reverse_complement("acgtN") == "Nacgt"
    assert H.reverse_complement("xZ") == "nn"                      # unknown characters become 'n'
    assert H.reverse_complement(H.reverse_complement("ACGTNRYSWKMBDHV")) == "ACGTNRYSWKMBDHV"
    assert H.encode_ambiguous(("A", "G")) == "R" and H.encode_ambiguous(("t", "C")) == "Y"
    assert H.encode_ambiguous(("a", "A")) == "A" and H.encode_ambiguous(("N", "C")) == "N"
    assert H.encode_ambiguous(("A", "C", "G", "T")) == "N"
    with pytest.raises(KeyError):
        H.encode_ambiguous(("R", "Y"))
    assert H.flatten_paired_alignment("AC-GT", "A-TGC") == "ACTGY"


def test_gapped_strings_from_fragments(H):
    """fragment list -> gapped rows exactly as ambivert.py:95-111 (seq2 keeps its own case)"""
    kat = golden("kat_reference_tests.json")
    seen = 0
    for c in kat:
        exp = c["expect"].get("nw")
        if not exp:
            continue
        frags = [tuple(f) for f in c["modes"]["global"]["frags"]]
        a1, a2 = H.gapped_strings(c["s1"], c["s2"], frags)
        if "aligned_s1" in exp:
            assert a1 == exp["aligned_s1"]
        if "aligned_s2" in exp:
            assert a2 == exp["aligned_s2"]
        seen += 1
    assert seen >= 5


def test_hash_table_format_round_trip(H):
    """save/load_hash_table keep the reference's pickle structure (ambivert.py:608-637)"""
    G = golden("pipeline_bundled.json")
    w = G["process_threshold50_overlap20"]
    amp = H.AmpliconData()
    for t in G["targets"]:
        amp.reference_sequences[tuple(t["key"])] = t["seq"]
    amp.reference = {k: tuple(v) for k, v in w["reference"].items()}
    amp.potential_variants = list(w["potential_variants"])
    buf = io.BytesIO()
    buf.close = lambda: None
    amp.save_hash_table(buf)
    sha, table = pickle.loads(buf.getvalue())
    assert len(sha) == 56 and set(table) == set(amp.reference) - set(amp.potential_variants)
    other = H.AmpliconData()
    other.reference_sequences = dict(amp.reference_sequences)
    other.load_hash_table(io.BytesIO(buf.getvalue()))
    assert other.reference == table
    mismatched = H.AmpliconData()
    mismatched.reference_sequences = {("x", "chr1", "1", "2", "+"): "ACGT"}
    with pytest.warns(UserWarning):
        mismatched.load_hash_table(io.BytesIO(buf.getvalue()))
    assert mismatched.reference == {}


def test_add_reads_keys_are_md5_of_trimmed_pair(H):
    import hashlib
    amp = H.AmpliconData(trim5=2, trim3=3)
    amp.add_reads("f", "AACCGGTT", "IIIIIIII", "r", "TTGGCCAA", "IIIIIIII")
    key = hashlib.md5(b"CCG" + b"GGC").hexdigest()
    assert list(amp.data) == [key] and amp.readpairs == 1
    assert amp.get_above_threshold(0) == [key] and amp.get_above_threshold(1) == []


def test_bench_merge_roofline_uses_the_committed_sass_count():
    """bench.py's issue-rate roofline of the traceback kernel divides by the instruction count of the fill loop in the committed
    SASS report (tools/sass_report.py): the two must not drift apart"""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bench = open(os.path.join(root, "bench.py")).read()
    rep = open(os.path.join(root, "profiles", "r02_sass_tb16_local_k5_step.txt")).read()
    want = int(re.search(r"steady-state loop 0x[0-9a-f]+\.\.0x[0-9a-f]+: (\d+) instructions", rep).group(1))
    got = int(re.search(r"^TB16_LOCAL_K5_LOOP_INSTR = (\d+)", bench, re.M).group(1))
    assert got == want


def test_bulk_string_helpers_match_the_per_sequence_forms():
    """reverse_complement_many / _reject_gapped / _ascii are bulk forms of per-sequence operations of the reference
    (sequence_utilities.py:96-128, ambivert.py:80-84): same results, same exceptions"""
    import random
    import numpy as np
    import pytest
    from ambivert_b200 import host as H
    rng = random.Random(5)
    seqs = ["".join(rng.choice("ACGTNacgtnRYKMxz.-") for _ in range(rng.randint(0, 40))) for _ in range(500)]
    assert H.reverse_complement_many(seqs) == [H.reverse_complement(s) for s in seqs]
    assert H.reverse_complement_many(iter(seqs[:7])) == [H.reverse_complement(s) for s in seqs[:7]]
    assert H.reverse_complement_many([]) == [] and H.reverse_complement_many([""]) == [""]
    assert H.reverse_complement_many(["A\nC", "G"]) == ["G\nT".replace("\n", "n"), "C"]      # a line break is an unknown character: 'n'
    H._reject_gapped(["ACGT", "AC"], ["A", "T"])
    with pytest.raises(RuntimeError) as e:
        H._reject_gapped(["ACGT", "A-C", "-"], ["A", "T", "G"])
    assert e.value.args[1:] == ("A-C", "T")                       # the FIRST offending pair, as the reference's loop reports it
    buf, off = H._ascii(["ACG", "", "TT"])
    assert bytes(buf) == b"ACGTT" and off.tolist() == [0, 3, 3, 5] and off.dtype == np.int64
    with pytest.raises(UnicodeEncodeError):
        H._ascii(["ACé"])
    with pytest.raises(TypeError):
        H._ascii([b"ACGT"])
