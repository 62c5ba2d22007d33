#!/usr/bin/env python3
"""oracle/make_golden.py -- regenerate tests/golden/*.json from the REAL reference.

Run in the build container only (needs /root/reference):   python oracle/make_golden.py

  * kat_reference_tests.json : the known-answer alignments the reference's own tests assert
        (ambivert/tests/test_ambivert.py:111-192) plus SURVEY.md section 8a's edge cases, each with the
        full (score, fragments) the reference C extension returns for all four modes.
  * fuzz_ref_vectors.json    : seeded random cases (all modes, DNA/IUPAC/lower case, random matrices,
        random gap penalties incl. 0, a 24-letter alphabet) with the reference C extension's outputs.
  * pipeline_bundled.json    : the reference PYTHON pipeline (ambivert.ambivert.AmpliconData, imported
        from a scratch copy of /root/reference with the compiled reference library beside it) run on
        its bundled test data: inputs of the hot path (unique read pairs, targets) and its outputs
        (merged reads, chosen targets under the five parameterisations of test_ambivert.py:237-270,
        aligned strings, locations, potential variants).
Nothing here is read at test time on the GPU box; only the JSON files are.
"""
import hashlib
import io
import json
import os
import random
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def frags(res):
    if isinstance(res, int):
        return {"error": res}
    return {"score": res[0], "frags": [list(f) for f in res[1]]}


def kat():
    r = O.ref()
    cases = []

    def add(name, s1, s2, expect=None):
        c = {"name": name, "s1": s1, "s2": s2, "expect": expect or {}, "modes": {}}
        for m in range(4):
            # pipeline call sites upper-case the target (ambivert.py:85, :126, :169)
            c["modes"][O.MODE_NAMES[m]] = frags(r.align(m, s1.encode(), s2.upper().encode()))
        cases.append(c)

    # test_smithwaterman_align_DNA (:111-143) / test_needlemanwunsch_align_DNA (:145-192)
    a = 'GCGCCAGCCACTCAACTATATTTTGGTTAATATCTCCTTGGCTGGTTTCATCTACTGCATCTTCTCGGTCTTCACTGTCTTCATCTCTAGCTCCCAAGGC'
    b = 'GCTGGTTTCATCTACTGCATCTTCTCGGTCTTCACTGTCTTCATCTCTAGCTCCCAAGGCTACTTCATCTTTGGCCGCCATGTTTGTGCTATGGAGGGTT'
    add("sw1_nw1", a, b, {
        "sw": {"aligned_s1": a[40:], "aligned_s2": a[40:], "start_s1": 40, "start_s2": 0},
        "nw": {"aligned_s1": a + '-' * 40, "aligned_s2": '-' * 40 + b, "start_s1": 0, "start_s2": 0}})
    a = 'GCGCCAGCCACTCAACTATATTTTGGTTAATATCTCCTTCACACACACACACTTCTCGGTCGTCGTCGTCTTCACTGTCTTCATCTCTAGCTCCCAAGGC'
    b = 'GCGCCAGCCACTCAACTATATTTGGTTAATATCTCCTTCACACACACACTTCTCGGTCGTCGTCTTCACTGTCTTCATCTCTAGCTCCCAAGGC'
    add("sw2", a, b, {"sw": {"aligned_s2": 'GCGCCAGCCACTCAACTATA-TTTGGTTAATATCTCCTT--CACACACACACTTCTCG---GTCGTCGTCTTCACTGTCTTCATCTCTAGCTCCCAAGGC'}})
    add("nw2", 'GC' + a, b, {"nw": {"aligned_s2": '--GCGCCAGCCACTCAACTATA-TTTGGTTAATATCTCCTT--CACACACACACTTCTCG---GTCGTCGTCTTCACTGTCTTCATCTCTAGCTCCCAAGGC'}})
    b = 'GCGCCAGCCACTCAACTATATTTGGTTAATATCTCCTGCACACTTCTCGGTCGTCGTCTTCACTGTCTTCATCTCTAGCTCCCAAGGC'
    g3 = 'GCGCCAGCCACTCAACTATA-TTTGGTTAATATCTCCT--------GCACACTTCTCG---GTCGTCGTCTTCACTGTCTTCATCTCTAGCTCCCAAGGC'
    add("sw3_nw3", a, b, {"sw": {"aligned_s2": g3}, "nw": {"aligned_s2": g3}})
    a = 'GCGCCAGCCACTCAACTATAGGTTAATATCTC'
    b = 'GCGCCAGCCACTCAACTATATTTGGTTAATATCTC'
    g4 = 'GCGCCAGCCACTCAACTATA---GGTTAATATCTC'
    add("sw4_nw6", a, b, {"sw": {"aligned_s1": g4}, "nw": {"aligned_s1": g4}})
    b = 'GAAAAAGATACAACCGTATTCAGAATACTGTATTTTATGTTTTTTTCTCATTGGTGATATAATTTATTTTCTTAAAATAGCCATGCTGGCTGGAGCCACAGCAGTTTACTCCCAGTTCATTACTCAGCTAACAGACGAAAACCAGTCATGTTGCCCCGTTTGTCAGAGA'
    a = 'tagaaaaagatacaaccgtattcagaatacTGTATTTTATGTTTTTTTCTCATTGGTGATATAATTTATTTTCTTAAAATAGCCATGCTGGCTGGAGCCACAGCAGTTTACTCCCAGTTCATTACTCAGCTAACAGACGAAAACcagtcatgttgccccgtttgtcagaga'
    add("nw4_softmask", a, b, {"nw": {"aligned_s1": a, "aligned_s2": '--' + b}})
    a = 'taaaaaagatacaaccgtattcagaatacTGTATTTTATGTTTTTTTCTCATTGGTGATATAATTTATTTTCTTAAAATAGCCATGCTGGCTGGAGCCACAGCAGTTTACTCCCAGTTCATTACTCAGCTAACAGACGAAAACcagtcatgttgccccgtttgtcagaga'
    add("nw5_softmask", a, b, {"nw": {"aligned_s1": a, "aligned_s2": '-' + b}})
    # SURVEY 8a edge cases
    for n, (x, y) in {"all_mismatch": ("AAAA", "CCCC"), "n_in_query": ("ANNA", "ACGA"), "iupac": ("ARGA", "ACGA"),
                      "lower": ("acga", "ACGA"), "short_a": ("A", "ACGT"), "short_b": ("ACGT", "A"),
                      "dash": ("G-C", "G-C"), "single": ("A", "A"), "nn": ("NN", "NN")}.items():
        add(n, x, y)
    return cases


def rand_seq(rng, n, alphabet):
    return "".join(rng.choice(alphabet) for _ in range(n))


def mutate(rng, s, alphabet, nmut):
    s = list(s)
    for _ in range(nmut):
        k = rng.random()
        if k < 0.5 and s:
            s[rng.randrange(len(s))] = rng.choice(alphabet)
        elif k < 0.75 and len(s) > 2:
            p = rng.randrange(len(s)); del s[p:p + rng.randint(1, 4)]
        else:
            p = rng.randrange(len(s) + 1); s[p:p] = list(rand_seq(rng, rng.randint(1, 4), alphabet))
    return "".join(s) or rng.choice(alphabet)


def fuzz(n_cases=500, seed=0xA17B):
    rng = random.Random(seed)
    r = O.ref()
    out = []
    blos_alpha = 24
    for i in range(n_cases):
        kind = rng.choice(["dna", "dna", "dna", "dna_iupac", "homopolymer", "randmat", "alpha24"])
        if kind in ("dna", "dna_iupac", "homopolymer"):
            alphabet = {"dna": "ACGT", "dna_iupac": "ACGTNRYacgtn-", "homopolymer": "AAAAAC"}[kind]
            alpha, mp, M = O.DNA_ALPHA, O.DNA_MAP, [int(x) for x in O.DNA_SCORE]
        elif kind == "randmat":
            alpha = rng.randint(2, 6)
            alphabet = "ACGTNRY"[:alpha]
            _, mp = O.make_map(alphabet, alphabet[-1], True)
            M = [rng.randint(-6, 6) for _ in range(alpha * alpha)]
        else:
            alpha = blos_alpha
            alphabet = "ARNDCQEGHILKMFPSTWYVBZX*"
            _, mp = O.make_map(alphabet, "X", True)
            M = [rng.randint(-4, 4) if a != b else rng.randint(4, 11) for a in range(alpha) for b in range(alpha)]
        la = rng.choice([1, 2, 3, rng.randint(4, 40), rng.randint(4, 40), rng.randint(20, 120), rng.randint(100, 300)])
        a = rand_seq(rng, la, alphabet)
        if rng.random() < 0.6:
            b = mutate(rng, a, alphabet, rng.randint(0, 6))
            if rng.random() < 0.3:
                b = rand_seq(rng, rng.randint(0, 30), alphabet) + b
            if rng.random() < 0.3:
                b = b + rand_seq(rng, rng.randint(0, 30), alphabet)
            if rng.random() < 0.3 and len(b) > 10:
                b = b[rng.randint(0, len(b) // 2):]
        else:
            b = rand_seq(rng, rng.choice([1, 2, rng.randint(3, 60), rng.randint(50, 300)]), alphabet)
        if kind == "dna" and rng.random() < 0.7:
            go, ge = -7, -1
        else:
            go, ge = -rng.randint(0, 12), -rng.randint(0, 4)
        case = {"kind": kind, "a": a, "b": b, "alpha": alpha, "alphabet": alphabet if kind in ("randmat", "alpha24") else "DNA",
                "unknown": {"randmat": alphabet[-1], "alpha24": "X"}.get(kind, "N"),
                "matrix": M if kind in ("randmat", "alpha24") else "DNA", "go": go, "ge": ge, "modes": {}}
        for m in range(4):
            case["modes"][O.MODE_NAMES[m]] = frags(r.align(m, a.encode(), b.encode(), alpha, mp, M, go, ge))
        out.append(case)
    # argument validation (global_align.c:118-122)
    bad = []
    for go, ge, a, b in [(1, -1, "ACGT", "ACGT"), (-1, 1, "ACGT", "ACGT"), (-7, -1, "", "ACGT"), (-7, -1, "ACGT", "")]:
        bad.append({"a": a, "b": b, "go": go, "ge": ge,
                    "modes": {O.MODE_NAMES[m]: frags(r.align(m, a.encode(), b.encode(), go=go, ge=ge)) for m in range(4)}})
    return {"cases": out, "invalid": bad}


def pipeline():
    tmp = tempfile.mkdtemp(prefix="ambivert_ref_")
    try:
        shutil.copytree(os.path.join(REF, "ambivert"), os.path.join(tmp, "ambivert"))
        subprocess.check_call("gcc -O2 -fPIC -shared -std=gnu99 -w -o %s %s/lib/align/*.c" % (
            os.path.join(tmp, "ambivert", "align", "align_c.so"), REF), shell=True)
        sys.path.insert(0, tmp)
        import logging
        import warnings
        warnings.simplefilter("ignore")
        from ambivert import ambivert as A
        logging.disable(logging.CRITICAL)
        data = os.path.join(tmp, "ambivert", "tests", "data")

        def fresh(threshold=0, overlap=10):
            amp = A.AmpliconData()
            amp.process_twofile_readpairs(open(os.path.join(data, "testdata_R1.fastq"), "rb"), open(os.path.join(data, "testdata_R2.fastq"), "rb"))
            amp.add_references_from_manifest(open(os.path.join(data, "testdatamanifest.txt")))
            amp.merge_overlaps(threshold=threshold, minimum_overlap=overlap)
            return amp

        md5 = lambda s: hashlib.md5(s.encode("ascii")).hexdigest()
        amp = fresh()
        out = {
            "source": "ambivert.ambivert.AmpliconData on ambivert/tests/data (reference python + reference C)",
            "pairs": [{"key": k, "count": len(v), "fwd": v[0][0][1], "rev": v[0][1][1]} for k, v in amp.data.items()],
            "targets": [{"key": list(k), "seq": s} for k, s in amp.reference_sequences.items()],
            "merge_threshold0_overlap10": {"merged": dict(amp.merged), "unmergable": list(amp.unmergable)},
            "match": [], "align": {},
        }
        # the five parameterisations of test_AmpliconData_match_to_reference (:237-270) with their md5s
        seeded = {'002b0ade8cda6a7bdc15f09ed5812126': ('BRCA1_Exon9_UserDefined_(9825051)_7473614_chr17_41243841_41244065', 'chr17', '41243841', '41244065', '-'),
                  'randomfoo': ('MadeUp', 'chr99', '1', '100', '+')}
        for kw, pre, want in [
                (dict(min_score=0.1, trim_primers=0, global_align=True), {}, 'e08f4ac8d71a067547f43746972e86dc'),
                (dict(min_score=0.1, trim_primers=0, global_align=False), {}, 'e08f4ac8d71a067547f43746972e86dc'),
                (dict(min_score=0.1, trim_primers=10, global_align=False), {}, 'e08f4ac8d71a067547f43746972e86dc'),
                (dict(min_score=10, trim_primers=50, global_align=True), {}, 'fbfdb58228dcdabe280e742a56127b91'),
                (dict(min_score=10, trim_primers=0, global_align=True), seeded, 'a963bb251be3c27f4780b8292cdc3e13')]:
            amp.reference = dict(pre)
            amp.match_to_reference(**kw)
            got = md5(str(sorted(list(amp.reference.items()))))
            assert got == want, (kw, got, want)
            out["match"].append({"kwargs": kw, "preseed": {k: list(v) for k, v in pre.items()},
                                 "reference": {k: list(v) for k, v in amp.reference.items()}, "md5": want})
        # test_AmpliconData_align_to_reference (:271-297)
        for ga, want_al, want_loc in [(True, 'ffb3aa886476f377b485c9d967a844c7', 'cd58f75e64192227c8c2c6beca7564cf'),
                                      (False, '0bec5d7ad6ca3e6f0a5f3bd366ab1c66', '1058c4388032a963aec5a03679d5a899')]:
            amp.aligned, amp.location, amp.potential_variants = {}, {}, []
            amp.align_to_reference(global_align=ga)
            assert md5(str(sorted(list(amp.aligned.items())))) == want_al
            assert md5(str(sorted(list(amp.location.items())))) == want_loc
            out["align"]["global" if ga else "local"] = {
                "aligned": {k: list(v) for k, v in amp.aligned.items()},
                "location": {k: list(v) for k, v in amp.location.items()},
                "potential_variants": list(amp.potential_variants),
                "md5_aligned": want_al, "md5_location": want_loc}
        # test_process_amplicon_data (:313-323): threshold 50, overlap 20, defaults
        amp = A.process_amplicon_data(os.path.join(data, "testdata_R1.fastq"), os.path.join(data, "testdata_R2.fastq"),
                                      manifest=open(os.path.join(data, "testdatamanifest.txt")), threshold=50, overlap=20)
        assert md5(str(sorted(amp.potential_variants))) == '1a28f10a7a1e2ea430e79453e367a342'
        out["process_threshold50_overlap20"] = {
            "merged": dict(amp.merged), "unmergable": list(amp.unmergable),
            "reference": {k: list(v) for k, v in amp.reference.items()},
            "aligned": {k: list(v) for k, v in amp.aligned.items()},
            "location": {k: list(v) for k, v in amp.location.items()},
            "potential_variants": list(amp.potential_variants),
            "md5_sorted_potential_variants": '1a28f10a7a1e2ea430e79453e367a342'}
        buf = io.BytesIO()
        buf.close = lambda: None
        amp.save_hash_table(buf)
        out["process_threshold50_overlap20"]["hash_table_pickle_sha1"] = hashlib.sha1(buf.getvalue()).hexdigest()
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    O.build(ref=True)
    os.makedirs(GOLD, exist_ok=True)
    for name, fn in [("kat_reference_tests.json", kat), ("fuzz_ref_vectors.json", fuzz), ("pipeline_bundled.json", pipeline)]:
        obj = fn()
        with open(os.path.join(GOLD, name), "w") as f:
            json.dump(obj, f, separators=(",", ":"), sort_keys=True)
        print("wrote", name, os.path.getsize(os.path.join(GOLD, name)), "bytes")


if __name__ == "__main__":
    main()
