"""Shared helpers for the test-suite: golden fixture loading and seeded sequence generators."""
import json
import os
import random

import numpy as np

import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def case_scoring(case):
    """(alpha, map256, matrix) of a fuzz_ref_vectors.json case"""
    if case["alphabet"] == "DNA":
        return O.DNA_ALPHA, O.DNA_MAP, O.DNA_SCORE
    alpha, mp = O.make_map(case["alphabet"], case["unknown"], True)
    assert alpha == case["alpha"]
    return alpha, mp, np.asarray(case["matrix"], dtype=np.int32)


def as_result(entry):
    """golden {'score','frags'} / {'error'} -> the tuple/int shape oracle.Port.align returns"""
    if "error" in entry:
        return entry["error"]
    return entry["score"], [tuple(f) for f in entry["frags"]]


def rand_dna(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def mutate(rng, s, nmut, alphabet="ACGT"):
    s = list(s)
    for _ in range(nmut):
        k = rng.random()
        if k < 0.6 and s:
            s[rng.randrange(len(s))] = rng.choice(alphabet)
        elif k < 0.8 and len(s) > 4:
            p = rng.randrange(len(s)); del s[p:p + rng.randint(1, 5)]
        else:
            p = rng.randrange(len(s) + 1); s[p:p] = list(rand_dna(rng, rng.randint(1, 5), alphabet))
    return "".join(s) or "A"


def gapped_strings(seq1, seq2, frags):
    """fragments -> the two gapped strings, as ambivert.py:95-111 builds them"""
    a1, a2 = [], []
    for typ, sa, sb, ln in frags:
        if typ == O.MATCH:
            a1.append(seq1[sa:sa + ln]); a2.append(seq2[sb:sb + ln])
        elif typ == O.A_GAP:
            a1.append("-" * ln); a2.append(seq2[sb:sb + ln])
        else:
            a1.append(seq1[sa:sa + ln]); a2.append("-" * ln)
    return "".join(a1), "".join(a2)
