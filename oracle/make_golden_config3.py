#!/usr/bin/env python3
"""oracle/make_golden_config3.py -- the reference PIPELINE (Python + C) on a config-3-shaped synthetic panel.

Run in the build container only (needs /root/reference and oracle/_ref; ~10 minutes on 8 cores):
    python oracle/make_golden_config3.py
Inputs: ambivert_b200.synth.panel_pipeline_inputs (300 targets, ~20k unique read pairs, seeded; not stored).
Runs the reference's own AmpliconData: add_reads -> merge_overlaps(threshold=0, minimum_overlap=20) ->
match_to_reference(threads=ncores) -> align_to_reference -> print_consolidated_vcf, and writes
tests/golden/config3_pipeline.json: md5s of merged / reference (the hash table) / aligned / location /
potential_variants in the reference tests' own style (test_ambivert.py:237-323), the sha256 of the pickled hash
table and of the VCF text, plus a sha256 of the inputs.
"""
import hashlib
import io
import json
import os
import shutil
import sys
import tempfile
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ambivert_b200 import synth  # noqa: E402

md5 = lambda s: hashlib.md5(s.encode("ascii")).hexdigest()


def inputs_sha(refs, pairs):
    h = hashlib.sha256()
    for k, s in refs:
        h.update(repr(k).encode()); h.update(s.encode())
    for f, r, c in pairs:
        h.update(f.encode()); h.update(r.encode()); h.update(bytes([c]))
    return h.hexdigest()


def state_digests(amp):
    return {"n_merged": len(amp.merged), "n_unmergable": len(amp.unmergable), "n_reference": len(amp.reference),
            "n_aligned": len(amp.aligned), "n_potential_variants": len(amp.potential_variants),
            "md5_merged": md5(str(sorted(amp.merged.items()))),
            "md5_reference": md5(str(sorted(list(amp.reference.items())))),
            "md5_aligned": md5(str(sorted(list(amp.aligned.items())))),
            "md5_location": md5(str(sorted(list(amp.location.items())))),
            "md5_potential_variants": md5(str(sorted(amp.potential_variants))),
            "unmergable": list(amp.unmergable)[:50]}


def main():
    tmp = tempfile.mkdtemp(prefix="abv_ref_")
    try:
        shutil.copytree("/root/reference/ambivert", os.path.join(tmp, "ambivert"))
        shutil.copy(os.path.join(ROOT, "oracle", "_ref", "libalign_ref.so"), os.path.join(tmp, "ambivert", "align", "align_c.so"))
        sys.path.insert(0, tmp)
        warnings.simplefilter("ignore")
        import logging
        from ambivert import ambivert as A
        logging.disable(logging.CRITICAL)
        refs, pairs = synth.panel_pipeline_inputs()
        amp = A.AmpliconData()
        for k, s in refs:
            amp.reference_sequences[k] = s
        for f, r, c in pairs:
            for _ in range(c):
                amp.add_reads("f", f, "I" * len(f), "r", r, "I" * len(r))
        out = {"inputs_sha256": inputs_sha(refs, pairs), "n_targets": len(refs), "n_unique_pairs": len(amp.data), "read_pairs": amp.readpairs}
        t0 = time.time(); amp.merge_overlaps(threshold=0, minimum_overlap=20); t1 = time.time()
        amp.match_to_reference(threads=os.cpu_count() or 1); t2 = time.time()
        amp.align_to_reference(); t3 = time.time()
        out["reference_cpu_seconds"] = {"merge": t1 - t0, "match": t2 - t1, "align": t3 - t2, "cores": os.cpu_count()}
        out.update(state_digests(amp))
        buf = io.BytesIO(); buf.close = lambda: None
        amp.save_hash_table(buf)
        out["hash_table_pickle_sha256"] = hashlib.sha256(buf.getvalue()).hexdigest()
        # the step after the path: call_variants on every potential variant, one by one, so that
        # amplicons on which the reference raises (its own assert, call_variants.py:139) are recorded instead of stopping
        from ambivert.call_variants import call_variants
        called = []
        errors = {}
        for key in amp.potential_variants:
            sample, ref = amp.aligned[key]
            name, chromosome, position, end, strand = amp.reference[key]
            try:
                called.append((key, str(call_variants(sample, ref, chromosome, int(position)))))
            except Exception as e:
                errors[key] = type(e).__name__
        out["n_called"] = len(called)
        out["md5_called_variants"] = md5(str(sorted(called)))
        out["call_variants_errors"] = {"count": len(errors), "types": sorted(set(errors.values())),
                                       "md5_keys": md5(str(sorted(errors.items()))), "first_in_pipeline_order":
                                       next((k for k in amp.potential_variants if k in errors), None)}
        # consolidation over the amplicons the reference can call (the failing ones left out)
        amp.called_variants = {k: call_variants(amp.aligned[k][0], amp.aligned[k][1], amp.reference[k][1], int(amp.reference[k][2]))
                               for k in amp.potential_variants if k not in errors}
        t4 = time.time(); amp.consolidate_variants(); t5 = time.time()
        out["reference_cpu_seconds"]["consolidate"] = t5 - t4
        out["n_consolidated"] = len(amp.consolidated_variants)
        out["md5_consolidated"] = md5(str(sorted(amp.consolidated_variants)))
        out["md5_filtered"] = md5(str(list(amp.get_filtered_variants())))
        vcf_ok = io.StringIO()
        amp.print_consolidated_vcf(outfile=vcf_ok)
        out["vcf_of_callable_sha256"] = hashlib.sha256(vcf_ok.getvalue().encode()).hexdigest()
        out["vcf_of_callable_lines"] = vcf_ok.getvalue().count("\n")
        amp.called_variants, amp.consolidated_variants = {}, []
        vcf = io.StringIO()
        try:
            amp.print_consolidated_vcf(outfile=vcf)
            out["vcf_sha256"] = hashlib.sha256(vcf.getvalue().encode()).hexdigest()
            out["vcf_lines"] = vcf.getvalue().count("\n")
        except Exception as e:   # the VCF writer is outside the hot path; record why if it cannot run here
            out["vcf_error"] = repr(e)
        with open(os.path.join(ROOT, "tests", "golden", "config3_pipeline.json"), "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps({k: v for k, v in out.items() if k != "unmergable"}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
