"""Host mirror of ambivert/call_variants.py over the device variant scan.

Same names, argument meaning and error behaviour as the reference module:
  caller(variant_amplicon, reference, gap='-', softmask=True)      call_variants.py:32-128  (a generator)
  call_variants(variant_amplicon, reference, chromosome, ref_start=1, **kw)   :130-164
  format_as_vcf / call_variants_to_vcf                              :166-179
plus the batched forms the pipeline uses (one device call for every amplicon):
  caller_many(pairs, ...)          -> per pair: list of (type, start, pos, bases) or an exception instance
  call_variants_many(jobs, ...)    -> per job:  list of AmpliconVariant or an exception instance

The per-site walk runs in the CUDA library (abv_call_variants -> variants_kernel); this module only turns the
records into the reference's tuples by slicing the rows.  There is no CPU implementation of the walk here.
"""
import sys
from collections import namedtuple

from . import batch as B

AmpliconVariant = namedtuple('AmpliconVariant', 'chromosome, start, end, vcf_start, ref_allele, alt_allele')
_KIND = "XDI"


def _rows(seq):
    """the reference takes any indexable of single characters; rows reach the device as ASCII bytes"""
    if isinstance(seq, bytes):
        return seq
    if not isinstance(seq, str):
        seq = "".join(seq)
    return seq.encode("ascii")


def caller_many(pairs, gap='-', softmask=True):
    """pairs: sequence of (variant_amplicon, reference) gapped rows.  One device call; per pair the list of
    (type, start, pos, bases) the reference's caller() yields, or the exception it raises (not raised here)."""
    pairs = list(pairs)
    if not pairs:
        return []
    status, start, count, recs = B.call_variants([_rows(p[0]) for p in pairs], [_rows(p[1]) for p in pairs], softmask=softmask, gap=gap)
    out = []
    recs = recs.tolist()
    for k, (sample, ref) in enumerate(pairs):
        st = int(status[k])
        if st == B.VAR_ERR_LENGTH:
            out.append(RuntimeError('Sequence lengths are not equal:\n{0}\n{1}'.format(sample, ref)))
            continue
        if st == B.VAR_ERR_INDEX:
            out.append(IndexError('string index out of range'))
            continue
        got = []
        s = int(start[k])
        for kind, vstart, pos, length, at in recs[s:s + int(count[k])]:
            row = ref if kind == B.VAR_D else sample
            bases = row[at:at + length]
            if not isinstance(bases, str):
                bases = "".join(bases) if not isinstance(bases, bytes) else bases.decode("ascii")
            got.append((_KIND[kind], vstart, pos, bases))
        out.append(got)
    return out


def caller(variant_amplicon, reference, gap='-', softmask=True):
    """Mutation calling (call_variants.py:32-128): yields (type 'X'/'D'/'I', zero-based start in the gapped rows,
    one-based position in the ungapped reference, bases).  Lower-case reference is soft-masked primer sequence."""
    result = caller_many([(variant_amplicon, reference)], gap=gap, softmask=softmask)[0]
    if isinstance(result, Exception):
        raise result
    for variant in result:
        yield variant


def _to_amplicon_variants(variant_amplicon, reference, chromosome, ref_start, variants):
    """call_variants.py:134-163: the record arithmetic and allele slices, Python indexing semantics included"""
    result = []
    for kind, vstart, pos, bases in variants:
        start = ref_start + pos - 1
        if kind == 'X':
            assert variant_amplicon[vstart] == bases
            vcf_start, end = start, start
            ref_allele, alt_allele = reference[vstart], bases
        elif kind == 'D':
            vcf_start, end = start - 1, start + len(bases) - 1
            ref_allele, alt_allele = reference[vstart - 1:vstart + len(bases)], reference[vstart - 1]
        else:
            vcf_start, end = start - 1, start
            ref_allele, alt_allele = reference[vstart - 1], reference[vstart - 1] + bases
        result.append(AmpliconVariant(chromosome, start, end, vcf_start, ref_allele, alt_allele))
    return result


def call_variants_many(jobs, **kw):
    """jobs: sequence of (variant_amplicon, reference, chromosome, ref_start).  One device call; per job the list of
    AmpliconVariant that call_variants() returns, or the exception instance it raises (RuntimeError, IndexError,
    AssertionError -- the reference's own `assert`, call_variants.py:139)."""
    jobs = list(jobs)
    out = []
    for (sample, ref, chromosome, ref_start), variants in zip(jobs, caller_many([(j[0], j[1]) for j in jobs], **kw)):
        if isinstance(variants, Exception):
            out.append(variants)
            continue
        try:
            out.append(_to_amplicon_variants(sample, ref, chromosome, ref_start, variants))
        except (AssertionError, IndexError) as e:
            out.append(e)
    return out


def call_variants(variant_amplicon, reference, chromosome, ref_start=1, **kw):
    """Variants of one aligned amplicon as AmpliconVariant tuples (call_variants.py:130-164)"""
    result = call_variants_many([(variant_amplicon, reference, chromosome, ref_start)], **kw)[0]
    if isinstance(result, Exception):
        raise result
    return result


def format_as_vcf(variants, outfile=sys.stdout):
    for variant in variants:
        print(variant.chromosome, variant.vcf_start, ".", variant.ref_allele, variant.alt_allele, '.', 'PASS', '.', sep='\t', file=outfile)


def call_variants_to_vcf(variant_amplicon, reference, chromosome, ref_start=1, outfile=sys.stdout, **kw):
    format_as_vcf(call_variants(variant_amplicon, reference, chromosome, ref_start=ref_start, **kw), outfile=outfile)
