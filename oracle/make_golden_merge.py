#!/usr/bin/env python3
"""oracle/make_golden_merge.py -- golden merged reads / gapped rows from the reference PYTHON + C.

Run in the build container only (needs /root/reference and oracle/_ref):
    python oracle/make_golden_merge.py
Writes tests/golden/merge_strings.json.gz: 800 synthetic forward / (reverse-complemented) reverse read pairs
with substitutions, small indels, N calls and unrelated pairs, and for each what the reference computes:
  * merge: smith_waterman + flatten_paired_alignment + the assembly of merge_overlaps (ambivert.py:320-327),
    or null when the overlap is shorter than minimum_overlap=20 ("unmergable");
  * nw / sw: the gapped rows + start offsets needleman_wunsch / smith_waterman return for (read, soft-masked
    target) pairs (ambivert.py:75-155), as align_to_reference uses them.
"""
import gzip
import json
import os
import random
import shutil
import sys
import tempfile
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tmp = tempfile.mkdtemp(prefix="abv_ref_")
    try:
        shutil.copytree("/root/reference/ambivert", os.path.join(tmp, "ambivert"))
        shutil.copy(os.path.join(ROOT, "oracle", "_ref", "libalign_ref.so"), os.path.join(tmp, "ambivert", "align", "align_c.so"))
        sys.path.insert(0, tmp)
        warnings.simplefilter("ignore")
        from ambivert import ambivert as A
        from ambivert.sequence_utilities import flatten_paired_alignment
        rng = random.Random(0xF1)
        dna = lambda n: "".join(rng.choice("ACGT") for _ in range(n))

        def noisy(s, psub, pn, pindel):
            out = []
            for ch in s:
                r = rng.random()
                if r < psub:
                    out.append(rng.choice("ACGT"))
                elif r < psub + pn:
                    out.append("N")
                elif r < psub + pn + pindel:
                    if rng.random() < 0.5:
                        continue
                    out.append(ch); out.append(dna(rng.randint(1, 3)))
                else:
                    out.append(ch)
            return "".join(out)
        pairs = []
        for i in range(800):
            amp = dna(rng.randint(160, 290))
            fwd = noisy(amp[:150], 0.01, 0.003, 0.002)
            rev = dna(150) if rng.random() < 0.03 else noisy(amp[-150:], 0.01, 0.003, 0.002)
            if not fwd or not rev:
                continue
            fa, ra, fs, rs = A.smith_waterman(fwd, rev)
            merged = None
            if len(fa) >= 20:
                merged = fwd[:fs] + flatten_paired_alignment(fa, ra) + rev[rs + len(ra.replace('-', '')):]
            pairs.append({"fwd": fwd, "rev": rev, "merged": merged, "sw": [fa, ra, fs, rs]})
        aligns = []
        for i in range(300):
            ref = dna(rng.randint(150, 260))
            ref = ref[:25].lower() + ref[25:-25] + ref[-25:].lower()          # soft-masked primers
            q = noisy(ref.upper(), 0.01, 0.002, 0.004)
            if rng.random() < 0.2:
                q = q[rng.randint(0, 10):len(q) - rng.randint(0, 10)]
            aligns.append({"q": q, "ref": ref, "nw": list(A.needleman_wunsch(q, ref)), "sw": list(A.smith_waterman(q, ref))})
        with gzip.open(os.path.join(ROOT, "tests", "golden", "merge_strings.json.gz"), "wt") as f:
            json.dump({"minimum_overlap": 20, "pairs": pairs, "aligns": aligns}, f, separators=(",", ":"))
        print("pairs", len(pairs), "unmergable", sum(p["merged"] is None for p in pairs), "aligns", len(aligns))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
