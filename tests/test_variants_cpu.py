"""CPU tests for SURVEY.md section 8f row f4 (variant scan + consolidation), no GPU needed:

  * the oracle restatement (oracle/variants_oracle.py) against the reference's own outputs
    (tests/golden/variants.json.gz: known answers of tests/test_call_variants.py, 1500 fuzz cases);
  * the DEVICE walk (dp_core.cuh::variants_walk, compiled for the host by tests/emu) against the same vectors;
  * the host logic after the scan (ambivert_b200/variants.py: positions, overlap queries, consolidation, filters,
    VCF text) on the reference's bundled-data state, against the strings whose md5s the reference's tests assert
    (tests/test_ambivert.py:399-491, :544-557).
"""
import ctypes as C
import gzip
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import variants_oracle as VO  # noqa: E402

KIND = "XDI"


@pytest.fixture(scope="module")
def gold():
    with gzip.open(os.path.join(HERE, "golden", "variants.json.gz"), "rb") as f:
        return json.loads(f.read().decode())


@pytest.fixture(scope="module")
def emu():
    d = os.path.join(HERE, "emu")
    so, src = os.path.join(d, "libemu.so"), os.path.join(d, "emu.cpp")
    hdr = os.path.join(ROOT, "ambivert_b200", "csrc", "dp_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-o", so, src])
    lib = C.CDLL(so)
    lib.emu_variants.restype = C.c_int
    return lib


def all_cases(gold):
    return gold["kat_caller"] + gold["kat_call_variants"] + gold["fuzz"]


def test_golden_covers_the_edge_cases(gold):
    cases = all_cases(gold)
    errs = [c.get("caller_error") for c in cases]
    assert errs.count("RuntimeError") >= 10 and errs.count("IndexError") >= 10
    assert sum(1 for c in cases if c.get("call_variants_error") == "AssertionError") >= 1
    kinds = {v[0] for c in cases for v in c.get("caller", [])}
    assert kinds == {"X", "D", "I"}
    assert any(not c["softmask"] for c in cases) and any(c["sample"] == "" for c in cases)


def test_oracle_restatement_matches_reference_outputs(gold):
    for c in all_cases(gold):
        try:
            got = [list(v) for v in VO.scan(c["sample"], c["ref"], softmask=c["softmask"])]
        except (RuntimeError, IndexError) as e:
            assert c.get("caller_error") == type(e).__name__, c
            continue
        assert "caller_error" not in c and got == c["caller"], (c, got)


def emu_scan(lib, sample, ref, softmask):
    s, r = sample.encode("ascii"), ref.encode("ascii")
    cap = len(s) + 2
    recs = np.zeros((cap, 5), np.int32)
    n = lib.emu_variants(s, C.c_int(len(s)), r, C.c_int(len(r)), C.c_int(1 if softmask else 0), C.c_int(ord("-")),
                         recs.ctypes.data_as(C.c_void_p), C.c_int(cap))
    if n < 0:
        return {-1: "RuntimeError", -2: "IndexError"}[n]
    assert n <= cap
    out = []
    for kind, start, pos, length, at in recs[:n].tolist():
        row = ref if kind == 1 else sample
        out.append([KIND[kind], start, pos, row[at:at + length]])
    return out


def test_device_walk_on_host_matches_reference_outputs(gold, emu):
    for c in all_cases(gold):
        got = emu_scan(emu, c["sample"], c["ref"], c["softmask"])
        if isinstance(got, str):
            assert c.get("caller_error") == got, c
        else:
            assert "caller_error" not in c and got == c["caller"], (c, got)


def test_record_arithmetic_of_call_variants(gold):
    """ambivert_b200.call_variants._to_amplicon_variants (call_variants.py:134-163) on the reference's caller output"""
    from ambivert_b200.call_variants import _to_amplicon_variants
    n = 0
    for c in all_cases(gold):
        if "caller" not in c:
            continue
        try:
            got = [list(v) for v in _to_amplicon_variants(c["sample"], c["ref"], c["chromosome"], c["ref_start"], [tuple(v) for v in c["caller"]])]
        except (AssertionError, IndexError) as e:
            assert c.get("call_variants_error") == type(e).__name__, c
            continue
        assert got == c["call_variants"], c
        n += len(got)
    assert n > 1000


class State(object):
    """the attributes of the reference's AmpliconData that the post-scan host logic reads"""

    def __init__(self, g):
        from ambivert_b200.call_variants import AmpliconVariant
        self.threshold = g["threshold"]
        self.data = defaultdict(list, {k: [None] * n for k, n in g["counts"].items()})
        self.aligned = {k: tuple(v) for k, v in g["aligned"].items()}
        self.location = {k: tuple(v) for k, v in g["location"].items()}
        self.reference = {k: tuple(v) for k, v in g["reference"].items()}
        self.reference_sequences = {tuple(k): v for k, v in g["reference_sequences"]}
        self.potential_variants = list(g["potential_variants"])
        self.called_variants = {k: [AmpliconVariant(*v) for v in vs] for k, vs in g["called_variants"].items()}
        self.consolidated_variants = []


@pytest.mark.parametrize("name", ["threshold50_overlap20", "threshold0_overlap10"])
def test_host_consolidation_matches_reference(gold, name):
    from ambivert_b200 import variants as V
    g = gold["bundled"][name]
    amp = State(g)
    assert str(sorted(amp.called_variants.items())) == g["str_sorted_called_variants"]     # the namedtuple prints alike
    assert str(V.get_variant_positions(amp)) == g["str_variant_positions"]
    V.consolidate_variants(amp)
    assert str(sorted(amp.consolidated_variants)) == g["str_sorted_consolidated"]
    assert str(list(V.get_filtered_variants(amp, min_freq=0.5))) == g["str_filtered_min_freq_0.5"]
    assert str(list(V.get_filtered_variants(amp, min_reads=0, min_cover=0, min_freq=0.00001))) == g["str_filtered_all"]
    amp.consolidated_variants = []
    V.consolidate_variants(amp, exclude_softmasked_coverage=False)
    assert str(sorted(amp.consolidated_variants)) == g["str_sorted_consolidated_with_softmasked"]


def test_host_overlap_queries_vcf_and_homopolymer(gold):
    from ambivert_b200 import variants as V
    g = gold["bundled"]["threshold50_overlap20"]
    amp = State(g)
    assert sorted(V.get_amplicons_overlapping(amp, 'chr17', '41245364')) == g["overlapping"]["chr17:41245364"]
    assert sorted(V.get_amplicons_overlapping_without_softmasking(amp, 'chr17', '41245364')) == g["overlapping"]["chr17:41245364:nosoftmask"]
    # the made-up amplicon of tests/test_ambivert.py:427-437
    amp.location['testampliconidentifier'] = ('chrY', 1, 50, '+')
    amp.aligned['testampliconidentifier'] = ('AAAAAGACATGACATGACATGACATGACATGACATGACATGACATGACATTTTTT',
                                             'aaaaaGACATGACATGACATGACAT-----GACATGACATGACATGACATttttt')
    assert V.get_amplicons_overlapping(amp, 'chrY', '20') == ['testampliconidentifier']
    assert V.get_amplicons_overlapping_without_softmasking(amp, 'chrY', '6') == ['testampliconidentifier']
    assert V.get_amplicons_overlapping_without_softmasking(amp, 'chrY', '3') == []
    del amp.location['testampliconidentifier'], amp.aligned['testampliconidentifier']
    for key, want in g["homopolymer"].items():
        chrom, pos, minimum = key.split(":")
        assert V.is_homopolymer_at_position(amp, chrom, int(pos), minimum=int(minimum)) == want
    out = io.StringIO()
    V.print_consolidated_vcf(amp, outfile=out)           # called_variants present: no device call
    assert out.getvalue() == g["vcf"]
    # the literal lines of tests/test_ambivert.py:558-560
    assert "chr17\t41245466\t.\tG\tA\t.\tPASS\tDP=133;AC=61;AF=0.459;ALTAMPS=357199abf30c529a625b24916a341ff8;REFAMPS=32ff3ea55601305b5e8b3bd266c4080a\n" in out.getvalue()
    amp0 = State(gold["bundled"]["threshold0_overlap10"])
    out = io.StringIO()
    V.print_consolidated_vcf(amp0, min_freq=0.01, outfile=out)
    assert out.getvalue() == gold["bundled"]["threshold0_overlap10"]["vcf_min_freq_0.01"]
