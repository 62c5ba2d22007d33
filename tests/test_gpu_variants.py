"""GPU parity for SURVEY.md section 8f row f4: the variant scan on the device (abv_call_variants) through the host
mirror of ambivert/call_variants.py, against the reference's own outputs (tests/golden/variants.json.gz, made by
oracle/make_golden_variants.py from the unmodified reference Python) and, at full size, against the oracle
restatement on a sample plus size-independent properties."""
import gzip
import io
import json
import os
import random
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    with gzip.open(os.path.join(HERE, "golden", "variants.json.gz"), "rb") as f:
        return json.loads(f.read().decode())


def all_cases(gold):
    return gold["kat_caller"] + gold["kat_call_variants"] + gold["fuzz"]


def test_caller_matches_reference_outputs(gold):
    from ambivert_b200 import call_variants as CV
    for softmask in (True, False):
        cases = [c for c in all_cases(gold) if c["softmask"] == softmask]
        got = CV.caller_many([(c["sample"], c["ref"]) for c in cases], softmask=softmask)
        assert len(got) == len(cases)
        for c, g in zip(cases, got):
            if isinstance(g, Exception):
                assert c.get("caller_error") == type(g).__name__, c
            else:
                assert "caller_error" not in c and [list(v) for v in g] == c["caller"], (c, g)


def test_known_answers_of_the_reference_tests():
    """literal answers of tests/test_call_variants.py:41-79 and :95-105, one call each as the reference makes them"""
    from ambivert_b200 import call_variants as CV
    assert list(CV.caller('AAAA', 'AAGA')) == [('X', 2, 3, 'A')]
    assert list(CV.caller('AATA', 'AC-A')) == [('X', 1, 1, 'A'), ('I', 2, 3, 'T')]
    assert list(CV.caller('CAA-AT-T', 'gAAAAttt', softmask=False)) == [('X', 0, 1, 'C'), ('D', 3, 4, 'A'), ('D', 6, 7, 't')]
    assert list(CV.caller('CGGGGGCCATCAT', 'c--gggcCTTCAT', softmask=True)) == [('X', 8, 7, 'A')]
    assert list(CV.caller('CGGGGG---ATCAT', 'CGGGGGCCC-TCAT', softmask=True)) == [('D', 6, 7, 'CCC'), ('I', 9, 10, 'A')]
    assert list(CV.caller('CGGGGGCCATCA-', 'CGGGGGCCATCAG', softmask=True)) == [('D', 12, 13, 'G')]
    gen = CV.caller('CGGGGGCCATCAT', 'cggg-TTCAT', softmask=True)          # raises when iterated, like the reference's generator
    with pytest.raises(RuntimeError):
        next(gen)
    result = CV.call_variants(
        'CAGGACTGGCTGCCGGCCCTTCTCTCCAGGTACTGGCCCCACGGCCTGAAGACTTCACGCGGCCCAGACGTG-TCAGCGGGCAG---GTACCCCGGGCATGTGCA',
        'CAGGACTGGCTGCCGGCCCTTCTCTCCAGGTACTGGCCCCACGGCCTGAAGACTTCATGCGGCCCAGACGTGTTCAGCGG-CAGCTCGTACCCCGGG---GTGCA',
        'X', ref_start=153457150)
    assert repr(result) == (
        "[AmpliconVariant(chromosome='X', start=153457207, end=153457207, vcf_start=153457207, ref_allele='T', alt_allele='C'), "
        "AmpliconVariant(chromosome='X', start=153457222, end=153457222, vcf_start=153457221, ref_allele='GT', alt_allele='G'), "
        "AmpliconVariant(chromosome='X', start=153457230, end=153457230, vcf_start=153457229, ref_allele='G', alt_allele='GG'), "
        "AmpliconVariant(chromosome='X', start=153457233, end=153457235, vcf_start=153457232, ref_allele='GCTC', alt_allele='G'), "
        "AmpliconVariant(chromosome='X', start=153457246, end=153457246, vcf_start=153457245, ref_allele='G', alt_allele='GCAT')]")


def test_call_variants_matches_reference_outputs(gold):
    from ambivert_b200 import call_variants as CV
    for softmask in (True, False):
        cases = [c for c in all_cases(gold) if c["softmask"] == softmask]
        got = CV.call_variants_many([(c["sample"], c["ref"], c["chromosome"], c["ref_start"]) for c in cases], softmask=softmask)
        for c, g in zip(cases, got):
            if isinstance(g, Exception):
                assert c.get("call_variants_error") == type(g).__name__, c
            else:
                assert "call_variants_error" not in c and [list(v) for v in g] == c["call_variants"], (c, g)


@pytest.mark.parametrize("name", ["threshold50_overlap20", "threshold0_overlap10"])
def test_bundled_pipeline_after_alignment_matches_reference(gold, name):
    """from the reference's aligned state of its bundled data to the VCF text: device scan + host consolidation;
    the strings are those whose md5s tests/test_ambivert.py:399, :414, :454, :475, :491 assert"""
    from test_variants_cpu import State
    from ambivert_b200 import variants as V
    g = gold["bundled"][name]
    amp = State(g)
    amp.called_variants = {}
    V.call_amplicon_variants(amp)
    assert str(sorted(amp.called_variants.items())) == g["str_sorted_called_variants"]
    assert str(V.get_variant_positions(amp)) == g["str_variant_positions"]
    V.consolidate_variants(amp)
    assert str(sorted(amp.consolidated_variants)) == g["str_sorted_consolidated"]
    if "vcf" in g:
        amp = State(g)
        amp.called_variants = {}
        out = io.StringIO()
        V.print_consolidated_vcf(amp, outfile=out)
        assert out.getvalue() == g["vcf"]


def _synthetic_rows(n, seed):
    from ambivert_b200 import synth
    return synth.aligned_rows(n, seed)


def test_full_size_scan_properties_and_oracle_sample():
    """100k aligned amplicons (the size of BASELINE.json's hash-table job): every copy of a row pair gets the same
    records wherever it sits in the batch; identical rows give none; the distinct pairs match the oracle"""
    import variants_oracle as VO
    from ambivert_b200 import batch as B
    base = _synthetic_rows(2000, 0xF4)
    rng = random.Random(7)
    order = [rng.randrange(len(base)) for _ in range(100000)]
    samples = [base[i][0] for i in order]
    refs = [base[i][1] for i in order]
    status, start, count, recs = B.call_variants([s.encode() for s in samples], [r.encode() for r in refs])
    assert (status == 0).all() and int(count.sum()) == len(recs)
    first = {}
    for k, i in enumerate(order):
        got = recs[start[k]:start[k] + count[k]]
        if i not in first:
            first[i] = got
            want = VO.scan(base[i][0], base[i][1])
            row = lambda kind: base[i][1] if kind == 1 else base[i][0]
            assert [("XDI"[a], b, c, row(a)[e:e + d]) for a, b, c, d, e in got.tolist()] == want
        else:
            assert np.array_equal(got, first[i])
    plain = [r.replace("-", "") for r in refs[:1000]]
    st2, _s2, c2, _r2 = B.call_variants([r.upper().encode() for r in plain], [r.encode() for r in plain])
    assert (st2 == 0).all() and int(c2.sum()) == 0
    assert B.call_variants([], [])[3].shape[0] == 0
