#!/usr/bin/env python3
"""oracle/make_golden_variants.py -- golden vectors for SURVEY.md section 8f row f4 (variant scan + consolidation),
produced by the UNMODIFIED reference Python (ambivert/call_variants.py, ambivert/ambivert.py) in the build container.

Run here only (needs /root/reference; a minute):
    python oracle/make_golden_variants.py
Writes tests/golden/variants.json.gz:
  kat        the known answers of the reference's own tests (tests/test_call_variants.py:41-85 caller(),
             :95-134 call_variants()), re-evaluated through the reference and stored with its results
  fuzz       seeded random gapped row pairs (primer flanks of any length incl. none, gaps inside / abutting / after
             the primers, gap runs at the row ends, both rows gapped, N and lower-case sample bases, rows of unequal
             length, softmask on and off) -> list(caller(...)) and call_variants(...), or the exception type raised
  bundled    the reference pipeline on its bundled test data (process_amplicon_data, threshold 50, overlap 20; the
             run behind tests/test_ambivert.py:386-509, :532-557): called_variants, get_variant_positions,
             consolidated_variants for both settings of exclude_softmasked_coverage, get_filtered_variants, the VCF
             text, each checked against the md5 the reference's tests assert; and the same with threshold 0,
             overlap 10 (:478-491)
"""
import gzip
import hashlib
import importlib.util
import io
import json
import os
import random
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
md5 = lambda s: hashlib.md5(s.encode("ascii")).hexdigest()


def load_call_variants():
    spec = importlib.util.spec_from_file_location("ref_call_variants", os.path.join(REF, "ambivert", "call_variants.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def evaluate(CV, sample, ref, softmask=True, chromosome="chrT", ref_start=1000):
    case = {"sample": sample, "ref": ref, "softmask": softmask, "chromosome": chromosome, "ref_start": ref_start}
    try:
        case["caller"] = [list(x) for x in CV.caller(sample, ref, softmask=softmask)]
    except Exception as e:
        case["caller_error"] = type(e).__name__
    try:
        case["call_variants"] = [list(x) for x in CV.call_variants(sample, ref, chromosome, ref_start, softmask=softmask)]
    except Exception as e:
        case["call_variants_error"] = type(e).__name__
    return case


KAT_CALLER = [  # tests/test_call_variants.py:41-85 (inputs only; the answers are re-derived and asserted below)
    ('AAAA', 'AAGA', True), ('AAAA', 'AGGA', True), ('AAAA', 'AA-A', True), ('AA-A', 'AAAA', True), ('AAAA', 'ACGA', True),
    ('AATA', 'AC-A', True), ('AA-A', 'ACAA', True), ('AAAAA', 'AAGCA', True), ('AAAAA', 'AA-CA', True), ('AA-AA', 'AAACA', True),
    ('GAAAA', 'gAAGA', True), ('GAAAA', 'gAA-A', True), ('GAA-A', 'gAAAA', True), ('CAA-A', 'gAAAA', True), ('CAA-A', 'gAAAA', False),
    ('CAA-ATCT', 'gAAAAttt', True), ('CAA-ATCT', 'gAAAAttt', False), ('CAA-AT-T', 'gAAAAttt', True), ('CAA-AT-T', 'gAAAAttt', False),
    ('CAA-ATTT', 'gAAAAt-t', True), ('CAA-ATTT', 'gAAAAt-t', False),
    ('CGGGGGCCATCAT', 'c-ggggcCTTCAT', False), ('CGGGGGCCATCAT', 'c-ggggcCTTCAT', True), ('CGGGGGCCATCAT', 'cggggg-CTTCAT', False),
    ('CGGGGGCCATCAT', 'cggggg-CTTCAT', True), ('CGGGGGCCATCAT', 'c--gggcCTTCAT', False), ('CGGGGGCCATCAT', 'c--gggcCTTCAT', True),
    ('CGGGGGCCATCAT', 'cggggg--TTCAT', False), ('CGGGGGCCATCAT', 'cggggg--TTCAT', True),
    ('CGGGGGC-TCAT', 'CGGGGG-ATCAT', True), ('CGGGGGCC-TCAT', 'CGGGGG--ATCAT', True), ('CGGGGG-ATCAT', 'CGGGGGC-TCAT', True),
    ('CGGGGG--ATCAT', 'CGGGGGCC-TCAT', True), ('CGGGGG---ATCAT', 'CGGGGGCCC-TCAT', True),
    ('CGGGGGCCATCAT', 'CGGGGGCCATCAG', True), ('CGGGGGCCATCA-', 'CGGGGGCCATCAG', True), ('CGGGGGCCATCAT', 'CGGGGGCCATCA-', True),
    ('CGGGGGCCATCAT', 'cggg-TTCAT', True),      # unequal lengths: RuntimeError (:81-85)
]
KAT_CALLER_ANSWERS = {  # a few of the literal answers of test_call_variants.py, to prove the harness evaluates the reference
    ('AAAA', 'AAGA', True): [['X', 2, 3, 'A']],
    ('AATA', 'AC-A', True): [['X', 1, 1, 'A'], ['I', 2, 3, 'T']],
    ('CAA-AT-T', 'gAAAAttt', False): [['X', 0, 1, 'C'], ['D', 3, 4, 'A'], ['D', 6, 7, 't']],
    ('CGGGGGCCATCAT', 'c--gggcCTTCAT', False): [['I', 1, 2, 'GG'], ['X', 8, 7, 'A']],
    ('CGGGGG---ATCAT', 'CGGGGGCCC-TCAT', True): [['D', 6, 7, 'CCC'], ['I', 9, 10, 'A']],
    ('CGGGGGCCATCAT', 'CGGGGGCCATCA-', True): [['I', 12, 13, 'T']],
}
MUT = 'CAGGACTGGCTGCCGGCCCTTCTCTCCAGGTACTGGCCCCACGGCCTGAAGACTTCACGCGGCCCAGACGTG-TCAGCGGGCAG---GTACCCCGGGCATGTGCA'
REF1 = 'CAGGACTGGCTGCCGGCCCTTCTCTCCAGGTACTGGCCCCACGGCCTGAAGACTTCATGCGGCCCAGACGTGTTCAGCGG-CAGCTCGTACCCCGGG---GTGCA'
KAT_CALL = [  # tests/test_call_variants.py:95-134
    (MUT, REF1, 'X', 153457150), (MUT, 'caggact' + REF1[7:], 'X', 153457150),
    ('CAGGGGGACTGGCTGTCGGCC', 'cagG---ACTGGCTGCCGGCC', 'X', 153457150),
    ('CAGGGACTGGCTGTCGGCC', 'cagg-actGGCTGCCGGCC', 'X', 153457150),
    ('CAGGGGGACTGGCTGTCGGCC', 'cagg---actGGCTGCCGGCC', 'X', 153457150),
]


def fuzz_cases(CV, n_cases=1500, seed=0xF4CA):
    rng = random.Random(seed)
    out = []
    for _ in range(n_cases):
        core = rng.choice([0, 1, 2, 5, 20, 60, 120])
        l5 = rng.choice([0, 0, 1, 3, 8, 25])
        l3 = rng.choice([0, 0, 1, 3, 8, 25])
        ref = [rng.choice("acgt") for _ in range(l5)] + [rng.choice("ACGT") for _ in range(core)] + [rng.choice("acgt") for _ in range(l3)]
        if rng.random() < 0.1:      # case noise: upper-case islands in primers, lower-case in the core, N
            for _ in range(rng.randint(1, 4)):
                if ref:
                    k = rng.randrange(len(ref))
                    ref[k] = rng.choice([ref[k].swapcase(), "N", "n"])
        sample_row, ref_row = [], []
        rate = rng.choice([0.0, 0.02, 0.1, 0.3])
        for b in ref:
            x = rng.random()
            if x < rate / 3:        # deletion in the sample
                sample_row.append("-"); ref_row.append(b)
            elif x < 2 * rate / 3:  # insertion in the sample (1-3 bases), then the base
                for _ in range(rng.randint(1, 3)):
                    sample_row.append(rng.choice("ACGT")); ref_row.append("-")
                sample_row.append(b.upper()); ref_row.append(b)
            elif x < rate:          # substitution
                sample_row.append(rng.choice([c for c in "ACGTN" if c != b.upper()])); ref_row.append(b)
            else:
                sample_row.append(b.upper()); ref_row.append(b)
        y = rng.random()
        if y < 0.08:                # trailing insertion: the reference row ends in a gap run
            for _ in range(rng.randint(1, 3)):
                sample_row.append(rng.choice("ACGT")); ref_row.append("-")
        elif y < 0.14:              # leading insertion
            for _ in range(rng.randint(1, 3)):
                sample_row.insert(0, rng.choice("ACGT")); ref_row.insert(0, "-")
        elif y < 0.17 and sample_row:   # both rows gapped at one site
            k = rng.randrange(len(sample_row))
            sample_row.insert(k, "-"); ref_row.insert(k, "-")
        elif y < 0.20 and sample_row:   # lower-case sample base
            k = rng.randrange(len(sample_row))
            sample_row[k] = sample_row[k].lower()
        elif y < 0.22:              # rows of unequal length
            sample_row.append("A")
        out.append(evaluate(CV, "".join(sample_row), "".join(ref_row), softmask=rng.random() < 0.8,
                            chromosome=rng.choice(["chr1", "chrX"]), ref_start=rng.choice([1, 1000, 41243125])))
    return out


def jsonable(x):
    if isinstance(x, tuple) and hasattr(x, "_fields"):
        return [jsonable(v) for v in x]
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, dict):
        return {k: jsonable(v) for k, v in x.items()}
    return x


def bundled():
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

        def fresh(threshold, overlap):
            return A.process_amplicon_data(os.path.join(data, "testdata_R1.fastq"), os.path.join(data, "testdata_R2.fastq"),
                                           manifest=open(os.path.join(data, "testdatamanifest.txt")), threshold=threshold, overlap=overlap)

        def state(amp):
            return {"version": A.__version__, "threshold": amp.threshold,
                    "counts": {k: len(v) for k, v in amp.data.items()},
                    "aligned": {k: list(v) for k, v in amp.aligned.items()},          # insertion order matters downstream
                    "location": {k: list(v) for k, v in amp.location.items()},
                    "reference": {k: list(v) for k, v in amp.reference.items()},
                    "reference_sequences": [[list(k), v] for k, v in amp.reference_sequences.items()],
                    "potential_variants": list(amp.potential_variants)}

        def downstream(amp, out):
            amp.call_amplicon_variants()
            out["called_variants"] = jsonable(amp.called_variants)
            out["str_sorted_called_variants"] = str(sorted(amp.called_variants.items()))
            out["str_variant_positions"] = str(amp.get_variant_positions())
            amp.consolidate_variants()
            out["str_sorted_consolidated"] = str(sorted(amp.consolidated_variants))
            out["str_filtered_min_freq_0.5"] = str(list(amp.get_filtered_variants(min_freq=0.5)))
            out["str_filtered_all"] = str(list(amp.get_filtered_variants(min_reads=0, min_cover=0, min_freq=0.00001)))
            amp.consolidated_variants = []
            amp.consolidate_variants(exclude_softmasked_coverage=False)
            out["str_sorted_consolidated_with_softmasked"] = str(sorted(amp.consolidated_variants))

        amp = fresh(50, 20)
        out = state(amp)
        downstream(amp, out)
        assert md5(out["str_sorted_called_variants"]) == '395ea8c7d8c55a4a752d9fa47dae8c8a'              # test_ambivert.py:399
        assert md5(out["str_variant_positions"]) == '563614d673fc502e6a191b7d37dd2070'                   # :414
        assert md5(out["str_sorted_consolidated"]) == 'a38f22f05f6c7d91f6615c404b206b2a'                 # :454
        assert md5(out["str_sorted_consolidated_with_softmasked"]) == 'a38f22f05f6c7d91f6615c404b206b2a' # :459
        assert md5(out["str_filtered_min_freq_0.5"]) == '870ffb1202aa1f3acde75c18a6b4ff1f'               # :475
        out["overlapping"] = {                                                                         # :432-433
            "chr17:41245364": sorted(amp.get_amplicons_overlapping('chr17', '41245364')),
            "chr17:41245364:nosoftmask": sorted(amp.get_amplicons_overlapping_without_softmasking('chr17', '41245364'))}
        out["homopolymer"] = {"chr17:41243290:5": amp.is_homopolymer_at_position('chr17', 41243290),   # :526-528
                              "chr17:41243290:6": amp.is_homopolymer_at_position('chr17', 41243290, minimum=6),
                              "chr17:41243290:7": amp.is_homopolymer_at_position('chr17', 41243290, minimum=7)}
        amp = fresh(50, 20)
        vcf = io.StringIO()
        amp.print_consolidated_vcf(outfile=vcf)                                                          # :544-557
        out["vcf"] = vcf.getvalue()
        assert out["vcf"].count("\n") == 17 and "chr17\t41244435\t.\tT\tC\t.\tPASS\tDP=193;AC=122;AF=0.632" in out["vcf"]
        amp = fresh(0, 10)                                                                               # :478-491
        out0 = state(amp)
        downstream(amp, out0)
        assert md5(out0["str_filtered_all"]) == '4c5d87513811af17fe80bb7029490edf'                       # :491
        amp = fresh(0, 10)
        vcf = io.StringIO()
        amp.print_consolidated_vcf(min_freq=0.01, outfile=vcf)
        out0["vcf_min_freq_0.01"] = vcf.getvalue()
        out = {"threshold50_overlap20": out, "threshold0_overlap10": out0}
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    CV = load_call_variants()
    kat = [evaluate(CV, s, r, softmask=m) for s, r, m in KAT_CALLER]
    for c in kat:
        want = KAT_CALLER_ANSWERS.get((c["sample"], c["ref"], c["softmask"]))
        if want is not None:
            assert c["caller"] == want, (c, want)
    assert kat[-1].get("caller_error") == "RuntimeError"
    kat_call = [evaluate(CV, s, r, True, chrom, start) for s, r, chrom, start in KAT_CALL]
    assert kat_call[0]["call_variants"][1] == ['X', 153457222, 153457222, 153457221, 'GT', 'G']      # test_call_variants.py:101
    assert kat_call[2]["call_variants"][0] == ['X', 153457154, 153457154, 153457153, 'G', 'GGGG']    # :122
    obj = {"kat_caller": kat, "kat_call_variants": kat_call, "fuzz": fuzz_cases(CV), "bundled": bundled()}
    n_err = sum(1 for c in obj["fuzz"] if "caller_error" in c or "call_variants_error" in c)
    path = os.path.join(GOLD, "variants.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(obj, separators=(",", ":"), sort_keys=False).encode())
    print("wrote", path, os.path.getsize(path), "bytes;", len(obj["fuzz"]), "fuzz cases,", n_err, "raising")


if __name__ == "__main__":
    main()
