"""CPU tests: pin the oracle restatement (oracle/align_oracle.c) to the reference.

 * against the committed golden vectors produced by the real reference (oracle/make_golden.py),
 * against the known answers the reference's own tests assert (test_ambivert.py:111-192),
 * and, where oracle/_ref was built, live against the compiled reference on fresh random inputs.
"""
import random

import numpy as np
import pytest

import oracle as O
from util import as_result, case_scoring, gapped_strings, golden, mutate, rand_dna

MODES = [O.OVERLAP, O.LOCAL, O.GLOBAL, O.GLOCAL]


def test_kat_vectors(oracle_port):
    for c in golden("kat_reference_tests.json"):
        for m in MODES:
            want = as_result(c["modes"][O.MODE_NAMES[m]])
            got = oracle_port.align(m, c["s1"].encode(), c["s2"].upper().encode())
            assert got == want, (c["name"], O.MODE_NAMES[m])


def test_kat_strings_match_reference_test_assertions(oracle_port):
    """Rebuild the gapped strings the way ambivert.py:95-111 does and compare with the literals the
    reference tests assert."""
    n = 0
    for c in golden("kat_reference_tests.json"):
        for tag, mode in (("sw", O.LOCAL), ("nw", O.GLOBAL)):
            exp = c["expect"].get(tag)
            if not exp:
                continue
            score, frags = oracle_port.align(mode, c["s1"].encode(), c["s2"].upper().encode())
            a1, a2 = gapped_strings(c["s1"], c["s2"], frags)
            if "aligned_s1" in exp:
                assert a1 == exp["aligned_s1"], c["name"]
            if "aligned_s2" in exp:
                assert a2 == exp["aligned_s2"], c["name"]
            if tag == "sw" and "start_s1" in exp:
                assert frags[0][1] == exp["start_s1"] and frags[0][2] == exp["start_s2"]
            n += 1
    assert n >= 9


def test_survey_edge_vectors(oracle_port):
    """SURVEY.md section 8a golden vectors (captured from the reference)."""
    p = oracle_port
    assert p.align(O.LOCAL, b"AAAA", b"CCCC") == (0, [])
    assert p.align(O.GLOCAL, b"AAAA", b"CCCC") == (0, [(O.MATCH, 0, 0, 1)])
    assert p.align(O.GLOBAL, b"AAAA", b"CCCC") == (-16, [(O.MATCH, 0, 0, 4)])
    assert p.align(O.GLOBAL, b"ANNA", b"ACGA")[0] == 2
    assert p.align(O.GLOBAL, b"ARGA", b"ACGA")[0] == 3
    assert p.align(O.GLOBAL, b"acga", b"ACGA")[0] == 4
    assert p.align(O.GLOBAL, b"A", b"ACGT") == (-8, [(O.MATCH, 0, 0, 1), (O.A_GAP, 1, 1, 3)])
    assert p.align(O.GLOBAL, b"ACGT", b"A") == (-8, [(O.MATCH, 0, 0, 1), (O.B_GAP, 1, 1, 3)])
    assert p.align(O.GLOBAL, b"ACGT", b"ACGT", go=1) == O.ERR_ARGS
    assert p.align(O.GLOBAL, b"", b"ACGT") == O.ERR_ARGS
    assert p.align(O.GLOCAL, b"ACGT", b"A") == O.ERR_UNDEFINED
    assert p.align(O.OVERLAP, b"A", b"ACGT") == O.ERR_UNDEFINED


def test_fuzz_vectors(oracle_port):
    g = golden("fuzz_ref_vectors.json")
    assert len(g["cases"]) >= 400
    for i, c in enumerate(g["cases"]):
        alpha, mp, M = case_scoring(c)
        for m in MODES:
            want = as_result(c["modes"][O.MODE_NAMES[m]])
            got = oracle_port.align(m, c["a"].encode(), c["b"].encode(), alpha, mp, M, c["go"], c["ge"])
            assert got == want, (i, c["kind"], O.MODE_NAMES[m])
    for c in g["invalid"]:
        for m in MODES:
            want = as_result(c["modes"][O.MODE_NAMES[m]])
            if want == O.ERR_UNDEFINED:
                continue
            got = oracle_port.align(m, c["a"].encode(), c["b"].encode(), go=c["go"], ge=c["ge"])
            assert got == O.ERR_ARGS and want == O.ERR_ARGS


def test_live_differential_vs_compiled_reference(oracle_port, oracle_ref):
    rng = random.Random(20261019)
    for it in range(3000):
        m = rng.choice(MODES)
        alpha = rng.choice([2, 4, 5, 5, 5, 7])
        la, lb = rng.randint(1, 70), rng.randint(1, 70)
        sa = bytes(rng.randrange(alpha) for _ in range(la))
        if rng.random() < 0.5:
            sb = bytearray(sa)
            for _ in range(rng.randint(0, 5)):
                if len(sb) > 2 and rng.random() < 0.5:
                    del sb[rng.randrange(len(sb))]
                else:
                    sb.insert(rng.randrange(len(sb) + 1), rng.randrange(alpha))
            sb = bytes(sb)
        else:
            sb = bytes(rng.randrange(alpha) for _ in range(lb))
        M = list(O.DNA_SCORE) if (alpha == 5 and rng.random() < 0.6) else [rng.randint(-5, 5) for _ in range(alpha * alpha)]
        go, ge = -rng.randint(0, 9), -rng.randint(0, 3)
        assert oracle_port.align_raw(m, sa, sb, alpha, M, go, ge) == oracle_ref.align_raw(m, sa, sb, alpha, M, go, ge)


def test_best_hits_rule(oracle_port):
    """first maximum wins, strict '>' against the seed (ambivert.py:478-485, :504-511)"""
    rng = random.Random(7)
    targets = [rand_dna(rng, rng.randint(30, 60)) for _ in range(12)]
    targets.insert(5, targets[2])               # duplicate target: index 2 must win over index 5
    queries = [mutate(rng, targets[rng.randrange(len(targets))], rng.randint(0, 3)) for _ in range(20)]
    queries[0] = targets[2]
    q = [O.encode(s) for s in queries]
    t = [O.encode(s) for s in targets]
    bs, bi = oracle_port.best_hits(O.GLOBAL, q, t, 5, O.DNA_SCORE, -7, -1, -1000000)
    for i in range(len(q)):
        scores = [oracle_port.align_raw(O.GLOBAL, q[i], tj, 5, O.DNA_SCORE, -7, -1)[0] for tj in t]
        assert bs[i] == max(scores) and bi[i] == int(np.argmax(scores))
    assert bi[0] == 2
    # seed 0: queries whose best score is <= 0 report idx -1 and the seed
    bs0, bi0 = oracle_port.best_hits(O.GLOBAL, [O.encode("A" * 40)], [O.encode("C" * 40)], 5, O.DNA_SCORE, -7, -1, 0)
    assert bs0[0] == 0 and bi0[0] == -1


def test_harness_threads_agree(oracle_port):
    rng = random.Random(11)
    targets = [O.encode(rand_dna(rng, rng.randint(40, 80))) for _ in range(9)]
    queries = [O.encode(rand_dna(rng, rng.randint(40, 80))) for _ in range(17)]
    bs, bi = oracle_port.best_hits(O.GLOBAL, queries, targets, 5, O.DNA_SCORE, -7, -1, -1000000)
    for th in (1, 3):
        sec, hs, hi = oracle_port.timed_best_hits(O.GLOBAL, queries, targets, 5, O.DNA_SCORE, -7, -1, -1000000, th)
        assert sec > 0 and (hs == bs).all() and (hi == bi).all()
    if O.Ref.available():
        sec, hs, hi = oracle_port.timed_best_hits(O.GLOBAL, queries, targets, 5, O.DNA_SCORE, -7, -1, -1000000, 2,
                                                  fns=O.ref().fns(O.GLOBAL))
        assert (hs == bs).all() and (hi == bi).all()
