"""The steps after the path: calling variants in the aligned amplicons and
consolidating them per reference position, re-hosted from the reference's AmpliconData
(ambivert/ambivert.py:666-745 call_amplicon_variants, get_variant_positions, consolidate_variants; :747-774
get_filtered_variants; :776-836 print_consolidated_vcf; :838-906 the overlap queries; :992-1019
is_homopolymer_at_position).

  * call_amplicon_variants makes ONE device call for all potential variants (abv_call_variants) instead of a
    Python loop over every site of every amplicon.
  * consolidate_variants asks "which amplicons overlap this position" once per variant position; the reference
    answers by scanning every amplicon (ambivert.py:838-859), quadratic at 1e5 amplicons.  Here the amplicon
    intervals sit in per-chromosome numpy arrays in the insertion order of `location`, so that one vectorised
    comparison returns the same ids in the same order.
Results, container types, ordering and error behaviour are the reference's (tests/golden/variants.json.gz holds
the reference's own outputs, checked against the md5s its tests assert).

The functions take any object with the reference AmpliconData's attributes, like those of host.py.
"""
import sys

import numpy as np

from .call_variants import call_variants_many
from .host import reverse_complement

__version__ = "0.5.2"      # ##source line of the VCF header (ambivert.py __version__, echoed by print_consolidated_vcf)


def get_amplicon_count(self, key):
    return len(self.data[key])


def call_amplicon_variants(self):
    """ambivert.py:666-676 with one batched scan; an amplicon on which the reference raises, raises here too
    (in the reference's iteration order: the first failing key stops the loop, earlier keys are stored)"""
    keys = list(self.potential_variants)
    jobs = []
    for key in keys:
        aligned_sample_seq, aligned_ref_seq = self.aligned[key]
        name, chromosome, amplicon_position, end, strand = self.reference[key]
        jobs.append((aligned_sample_seq, aligned_ref_seq, chromosome, int(amplicon_position)))
    for key, result in zip(keys, call_variants_many(jobs)):
        if isinstance(result, Exception):
            raise result
        self.called_variants[key] = result


def get_variant_positions(self):
    """sorted unique (chromosome, start, end) of every called variant (ambivert.py:678-688)"""
    positions = set()
    for amplicon_id in self.called_variants:
        for variant in self.called_variants[amplicon_id]:
            positions.add((variant.chromosome, variant.start, variant.end))
    return sorted(positions)


class _IntervalIndex(object):
    """closed intervals (chromosome, start, end) of `items` in iteration order, grouped by chromosome"""

    def __init__(self, items):
        self.keys = []
        by_chrom = {}
        for i, (key, chrom, start, end) in enumerate(items):
            self.keys.append(key)
            rows = by_chrom.setdefault(chrom, ([], [], []))
            rows[0].append(i); rows[1].append(int(start)); rows[2].append(int(end))
        self.by_chrom = {c: (np.asarray(r[0], np.int64), np.asarray(r[1], np.int64), np.asarray(r[2], np.int64))
                         for c, r in by_chrom.items()}

    def overlapping(self, chrom, pos, length=1):
        """keys whose interval meets [pos, pos+length-1], in insertion order (the test of ambivert.py:856)"""
        rows = self.by_chrom.get(chrom)
        if rows is None:
            return []
        pos = int(pos)
        length = int(length)
        idx, starts, ends = rows
        hit = idx[~((pos + length - 1 < starts) | (pos > ends))]
        return [self.keys[i] for i in hit.tolist()]


def _amplicon_index(self):
    return _IntervalIndex((key, loc[0], loc[1], loc[2]) for key, loc in self.location.items())


def get_amplicons_overlapping(self, chrom, pos, length=1):
    """ids of the amplicons that overlap a reference position (ambivert.py:838-859)"""
    return _amplicon_index(self).overlapping(chrom, pos, length)


def get_reference_overlapping(self, chrom, pos, length=1):
    """keys of the reference sequences that overlap a position (ambivert.py:861-883)"""
    return _IntervalIndex((key, key[1], key[2], key[3]) for key in self.reference_sequences).overlapping(chrom, pos, length)


def _without_softmasking(self, amplicon_ids, pos, length, ungapped):
    result = []
    for amplicon_id in amplicon_ids:
        start = self.location[amplicon_id][1]
        ref = ungapped.get(amplicon_id)
        if ref is None:
            ref = ungapped[amplicon_id] = self.aligned[amplicon_id][1].replace('-', '')
        window = ref[pos - start:pos - start + length + 1]          # Python slice semantics, as ambivert.py:903
        if not any(base.islower() for base in window):
            result.append(amplicon_id)
    return result


def get_amplicons_overlapping_without_softmasking(self, chrom, pos, length=1):
    """as get_amplicons_overlapping, keeping the amplicons whose reference bases there are not lower case
    (ambivert.py:885-906)"""
    pos = int(pos)
    length = int(length)
    return _without_softmasking(self, get_amplicons_overlapping(self, chrom, pos, length), pos, length, {})


def consolidate_variants(self, exclude_softmasked_coverage=True):
    """ambivert.py:690-745: per variant position the supporting / non-supporting amplicons and read depths"""
    index = _amplicon_index(self)
    ungapped = {}
    count = {}

    def depth(ids):
        total = 0
        for amplicon_id in ids:
            c = count.get(amplicon_id)
            if c is None:
                c = count[amplicon_id] = get_amplicon_count(self, amplicon_id)
            total += c
        return total

    for (chrom, start, end) in get_variant_positions(self):
        amplicon_ids = index.overlapping(chrom, start, end - start)
        if exclude_softmasked_coverage:
            amplicon_ids = _without_softmasking(self, amplicon_ids, int(start), int(end - start), ungapped)
        ref = []
        alt = {}
        for amplicon_id in amplicon_ids:
            overlapping_variant = False
            for variant in self.called_variants.get(amplicon_id, ()):
                if not ((variant.end < start) or (variant.start > end)):
                    overlapping_variant = True
                    alt.setdefault(variant, []).append(amplicon_id)
            if not overlapping_variant:
                ref.append(amplicon_id)
        total_depth = depth(amplicon_ids)
        for variant, ids in alt.items():         # indels first, then substitutions, each in first-seen order (:724-738)
            if len(variant.alt_allele) > 1 or len(variant.ref_allele) > 1:
                variant_depth = depth(ids)
                self.consolidated_variants.append((variant, variant_depth, total_depth, variant_depth / total_depth, tuple(ref), tuple(ids)))
        for variant, ids in alt.items():
            if len(variant.alt_allele) == 1 and len(variant.ref_allele) == 1:
                variant_depth = depth(ids)
                self.consolidated_variants.append((variant, variant_depth, total_depth, variant_depth / total_depth,
                                                   tuple(sorted(ref)), tuple(sorted(ids))))
    self.consolidated_variants = sorted(set(self.consolidated_variants))


def get_filtered_variants(self, min_cover=0, min_reads=0, min_freq=0.1):
    """consolidated variants that pass the depth / cover / frequency filters and touch neither soft-masked nor
    ambiguous bases (ambivert.py:747-774)"""
    for variant in self.consolidated_variants:
        alleles = variant[0].alt_allele + variant[0].ref_allele
        if variant[1] >= min_reads and variant[2] >= min_cover and variant[3] >= min_freq and \
                not any(base.islower() for base in alleles) and \
                all(base in '-GATC' for base in variant[0].alt_allele):
            yield variant


def is_homopolymer_at_position(self, chrom, pos, minimum=5):
    """does a homopolymer of at least `minimum` bases start at this reference position (ambivert.py:992-1019)"""
    reference_ids = get_reference_overlapping(self, chrom, pos, length=1)
    if not reference_ids:
        raise IndexError('No reference amplicons overlapping {0} {1}'.format(chrom, pos))
    ref = sorted(reference_ids)[-1]
    ref_seq = self.reference_sequences[ref]
    if ref[4] == '-':
        ref_seq = reverse_complement(ref_seq)
    first = pos - int(ref[2])
    base = ref_seq[first].upper()
    i = first
    while i < len(ref_seq) and ref_seq[i].upper() == base:
        i += 1
    return i - first >= minimum


_VCF_HEADER = (
    "##fileformat=VCFv4.2",
    "##source=AmBiVeRT{version}",
    '##FILTER=<ID=depth,Description="more than {depth} variant supporting reads">',
    '##FILTER=<ID=cover,Description="more than {cover} reads at variant position">',
    '##FILTER=<ID=freq,Description="more than {freq}% of reads support variant">',
    '##FILTER=<ID=primer,Description="involves primer sequence">',
    '##FILTER=<ID=homopoly5,Description="homopolymer of length > 5 at mutation site">',
    '##FILTER=<ID=homopoly10,Description="homopolymer of length > 10 at mutation site">',
    '##INFO=<ID=DP,Number=1,Type=Integer,Description="Read depth excluding soft masked primers at variant site">',
    '##INFO=<ID=AC,Number=1,Type=Integer,Description="Alt allele supporting read count">',
    '##INFO=<ID=AF,Number=1,Type=Float,Description="Alt allele frequency">',
    '##INFO=<ID=ALTAMPS,Number=.,Type=String,Description="Unique identifiers for amplicons supporting alt allele">',
    '##INFO=<ID=REFAMPS,Number=.,Type=String,Description="Unique identifiers for amplicons supporting other alleles">',
    "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO",
)


def print_consolidated_vcf(self, min_cover=0, min_reads=0, min_freq=0.1, outfile=sys.stdout):
    """the VCF of the filtered consolidated variants, byte-compatible with ambivert.py:776-836"""
    if not self.called_variants:
        call_amplicon_variants(self)
    if not self.consolidated_variants:
        consolidate_variants(self)
    header = "\n".join(_VCF_HEADER).format(version=__version__, depth=max(self.threshold, min_reads),
                                           cover=max(self.threshold, min_cover), freq=min_freq * 100)
    print(header, file=outfile)
    for variant in get_filtered_variants(self, min_cover=min_cover, min_reads=min_reads, min_freq=min_freq):
        v = variant[0]
        soft_filter_name = 'PASS'
        if is_homopolymer_at_position(self, v.chromosome, v.start, minimum=5):
            soft_filter_name = 'homopoly5'
        if is_homopolymer_at_position(self, v.chromosome, v.start, minimum=10):
            soft_filter_name = 'homopoly10'
        info = 'DP={DP};AC={AC};AF={AF:.3};ALTAMPS={ALTAMPS};REFAMPS={REFAMPS}'.format(
            DP=variant[2], AC=variant[1], AF=variant[3], ALTAMPS=",".join(variant[5]), REFAMPS=",".join(variant[4]))
        print(v.chromosome, v.vcf_start, ".", v.ref_allele, v.alt_allele, '.', soft_filter_name, info, sep='\t', file=outfile)
