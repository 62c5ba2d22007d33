"""Host side of the hot path: the three pipeline stages of the reference's AmpliconData that call
the aligner, re-hosted so that each stage makes ONE batched call into the CUDA library instead of one
ctypes call per pair.

  merge_overlaps      ambivert/ambivert.py:301-331   local alignment + traceback of every unique F/R pair
  match_to_reference  ambivert/ambivert.py:365-541   every merged amplicon x every target -> hash table
  align_to_reference  ambivert/ambivert.py:543-574   final global/local alignment to the chosen target

The functions take any object with the reference AmpliconData's attributes (`data`, `merged`,
`unmergable`, `reference`, `reference_sequences`, `aligned`, `location`, `potential_variants`,
`threshold`), so they can be bound onto the reference class
(`ambivert.ambivert.AmpliconData.match_to_reference = ambivert_b200.host.match_to_reference`) or used
through the small `AmpliconData` below.  Names, argument meaning, result containers, iteration order
and error behaviour follow the reference, including its quirks.
"""
import hashlib
import logging
import pickle
import warnings
from collections import defaultdict
from itertools import compress

import numpy as np

from . import batch as B

_COMP = {ord(a): b for a, b in zip("acgtACGT-.nNUMRWSYKVHDBumrwsykvhdb", "tgcaTGCA-.nNAKYWSRMBDHVakywsrmbdhv")}
_IUPAC = {"AG": "R", "CT": "Y", "CG": "S", "AT": "W", "GT": "K", "AC": "M",
          "CGT": "B", "AGT": "D", "ACT": "H", "ACG": "V", "ACGT": "N"}


class _CompTable(dict):
    def __missing__(self, key):
        return "n"


_COMP_TABLE = _CompTable(_COMP)


def complement(seq):
    """case-sensitive IUPAC complement; unknown characters become 'n' (sequence_utilities.py:96-119)"""
    return seq.translate(_COMP_TABLE)


def reverse_complement(seq):
    return complement(seq)[::-1]


_COMP_TABLE_NL = _CompTable(_COMP)
_COMP_TABLE_NL[10] = "\n"


def reverse_complement_many(seqs):
    """[reverse_complement(s) for s in seqs] with three string operations for the whole list (join, translate, reverse)
    instead of two per sequence"""
    seqs = list(seqs)
    joined = "\n".join(seqs)
    if joined.count("\n") != max(len(seqs) - 1, 0):          # a sequence with a line break in it: one by one
        return [reverse_complement(s) for s in seqs]
    return joined.translate(_COMP_TABLE_NL)[::-1].split("\n")[::-1] if seqs else []


def encode_ambiguous(bases):
    """IUPAC symbol of a set of bases; N if any is N (sequence_utilities.py:130-156)"""
    s = "".join(sorted(set(b.upper() for b in bases)))
    if len(s) == 1:
        return s
    if "N" in s:
        return "N"
    return _IUPAC[s]          # KeyError for non-ACGT input, as in the reference


def flatten_paired_alignment(seq1, seq2, gap="-"):
    """merge two gapped rows: a gap takes the other base, a disagreement its IUPAC code
    (sequence_utilities.py:158-169)"""
    out = []
    for a, b in zip(seq1, seq2):
        if a == gap:
            out.append(b)
        elif b == gap or a == b:
            out.append(a)
        else:
            out.append(encode_ambiguous((a, b)))
    return "".join(out)


def gapped_strings(seq1, seq2, frags):
    """fragment list -> the two gapped rows, as ambivert.py:95-111 / :136-152 build them
    (slices of seq2 keep seq2's own case)"""
    r1, r2 = [], []
    for typ, sa, sb, n in frags:
        if typ == B.MATCH:
            r1.append(seq1[sa:sa + n]); r2.append(seq2[sb:sb + n])
        elif typ == B.A_GAP:
            r1.append("-" * n); r2.append(seq2[sb:sb + n])
        elif typ == B.B_GAP:
            r1.append(seq1[sa:sa + n]); r2.append("-" * n)
    a, b = "".join(r1), "".join(r2)
    assert len(a) == len(b)
    return a, b


def _ascii(seqs):
    """str sequences -> the packed (buffer, offsets) form of the batched calls with ONE join and ONE encode for the whole set (a
    bytes object per sequence costs more than the device work of the stage at 1e5 sequences); same conversion and same
    UnicodeEncodeError / TypeError as bytes(seq, "ascii") at the reference's call sites (ambivert.py:84)"""
    return B.pack_text(seqs if isinstance(seqs, list) else list(seqs))


def _reject_gapped(seqs1, seqs2):
    """ambivert.py:80-81, :121-122: the first pair with a gap character raises (one scan of the joined text unless there is one)"""
    if "-" in "".join(seqs1) or "-" in "".join(seqs2):
        for s1, s2 in zip(seqs1, seqs2):
            if "-" in s1 or "-" in s2:
                raise RuntimeError("Aligning Sequences with gaps is not supported", s1, s2)


def batched_pairwise(seqs1, seqs2, local, raw=False):
    """smith_waterman (local=True, ambivert.py:75-114) / needleman_wunsch (ambivert.py:116-155) for a
    whole batch: -> list of (align_seq1, align_seq2, start_seq1, start_seq2).  The aligner sees
    seq2.upper() in the reference; DNA_MAP is case-insensitive, so seq2 goes in as it is and the gapped
    rows, built on the device, keep its case as the reference's slices do."""
    _reject_gapped(seqs1, seqs2)
    try:
        sc, status, st1, st2, off, rows1, rows2 = B.gapped_pairs(_ascii(seqs1), _ascii(seqs2), B.LOCAL if local else B.GLOBAL)
    except B.AlignError as e:
        if e.code in (B.ERR_ARGS, B.ERR_UNDEFINED):
            raise ValueError("NULL pointer access") from e     # what ctypes raises on the reference's NULL result
        raise
    if local and (status < 0).any():
        raise ValueError("NULL pointer access")                # alignment.contents.align_frag.contents on an empty list
    t1, t2 = rows1.tobytes().decode("ascii"), rows2.tobytes().decode("ascii")
    offs = off.tolist()
    if raw:            # the two texts of all gapped rows, their offsets and the start positions: for callers that slice themselves
        return t1, t2, offs, (st1.tolist() if local else None), (st2.tolist() if local else None)
    if local:
        return [(t1[offs[k]:offs[k + 1]], t2[offs[k]:offs[k + 1]], int(st1[k]), int(st2[k])) for k in range(len(seqs1))]
    return [(t1[offs[k]:offs[k + 1]], t2[offs[k]:offs[k + 1]], 0, 0) for k in range(len(seqs1))]


def smith_waterman(seq1, seq2):
    return batched_pairwise([seq1], [seq2], True)[0]


def needleman_wunsch(seq1, seq2):
    return batched_pairwise([seq1], [seq2], False)[0]


def get_above_threshold(self, threshold=1):
    return [x for x in self.data if len(self.data[x]) > threshold]


def merge_overlaps(self, threshold=0, minimum_overlap=10):
    """ambivert.py:301-331 with ONE batched call: local alignment with traceback of every unique pair above
    threshold and assembly of the merged read on the device (abv_merge_pairs)"""
    self.threshold = threshold
    keys = get_above_threshold(self, threshold)
    logging.info("Merging overlaps for {0} unique pairs".format(len(keys)))
    if keys:
        fwds = [self.data[k][0][0][1] for k in keys]
        revs = reverse_complement_many(self.data[k][0][1][1] for k in keys)
        _reject_gapped(fwds, revs)
        try:
            sc, status, off, merged = B.merge_pairs(_ascii(fwds), _ascii(revs), minimum_overlap)
        except B.AlignError as e:
            if e.code in (B.ERR_ARGS, B.ERR_UNDEFINED):
                raise ValueError("NULL pointer access") from e
            raise
        if (status == -1).any():
            raise ValueError("NULL pointer access")            # smith_waterman on an empty alignment (ambivert.py:93)
        if (status == -2).any():
            k = int(np.argmax(status == -2))
            raise KeyError("mismatching bases outside ACGTN cannot be encoded (sequence_utilities.py:152-156)", fwds[k], revs[k])
        text = merged.tobytes().decode("ascii")
        offs = off.tolist()
        for k, key in enumerate(keys):
            if status[k] == 1:
                self.merged[key] = text[offs[k]:offs[k + 1]]
            else:
                self.unmergable.append(key)
        # The merged reads left the device packed (one buffer + offsets): keep them that way for match_to_reference, which
        # would otherwise join and encode the strings again.  Unmergable pairs have length 0 in the buffer, so the kept reads are
        # adjacent.  (keys, buffer, offsets); used only while self.merged still holds exactly these reads in this order.
        ok = status == 1
        self._merged_pack = ([key for key, good in zip(keys, ok.tolist()) if good],
                             np.ascontiguousarray(merged, dtype=np.uint8), np.concatenate([off[:-1][ok], off[-1:]]).astype(np.int64))
    logging.info("\nSuccessfully merged {0} reads. {1} reads could not be merged".format(len(self.merged), len(self.unmergable)))


def match_to_reference(self, min_score=0.1, trim_primers=0, global_align=True, show_progress=False, threads=1, gpus=1):
    """ambivert.py:365-541: the exact-match hash table.  Every merged amplicon that is not already
    in the table is scored against every reference target (global alignment in both settings of
    `global_align` -- the reference's calculate_local_score calls global_align too, ambivert.py:190; the
    settings differ by the arg-max seed, 0 vs -1000000) and keeps the first best-scoring target if its
    score exceeds min_score.  `threads` is accepted for signature compatibility; the GPUs do the work:
    `gpus` (the successor of --threads, ambivert.py:474-476, :1092-1095) = how many GPUs of the box, or a list of device
    ordinals -- the amplicons are sharded over them inside the library (abv_best_hits_multi), one process."""
    reference = self.reference
    keys = [k for k in self.merged if k not in reference]
    ref_keys = list(self.reference_sequences)
    if not keys:
        return
    merged = self.merged
    pack = getattr(self, "_merged_pack", None)
    if pack is not None and not trim_primers and len(keys) == len(merged) == len(pack[0]) and keys == pack[0]:
        seqs = None                      # the packed merged reads of merge_overlaps are exactly the amplicons to score
    else:
        pack = None
        seqs = [merged[k][trim_primers:-trim_primers] for k in keys] if trim_primers else [merged[k] for k in keys]
    targets = [self.reference_sequences[x].upper() for x in ref_keys]
    seed = B.SEED_GLOBAL if global_align else B.SEED_LOCAL
    logging.info("Matching read bins to references - {0} amplicons x {1} targets on the GPU".format(len(keys), len(ref_keys)))
    many = not isinstance(gpus, int) or gpus > 1
    try:
        # one join + one encode per set, not one bytes object per amplicon; the amplicons not even that when they come packed
        q, t = (pack[1], pack[2]) if pack is not None else B.pack_text(seqs), B.pack_text(targets)
        if many:
            best_score, best_idx = B.best_hits_multi(q, t, gpus, B.GLOBAL, seed)
        else:
            best_score, best_idx = B.best_hits(q, t, B.GLOBAL, seed)
    except B.AlignError as e:
        if e.code in (B.ERR_ARGS, B.ERR_UNDEFINED):
            raise ValueError("NULL pointer access") from e     # what ctypes raises on the reference's NULL result
        raise
    # best_hit and best_score > min_score (ambivert.py:528), for all amplicons at once; insertion order = order of `merged`
    truthy = np.fromiter((bool(r) for r in ref_keys), dtype=bool, count=len(ref_keys))
    hit = (best_idx >= 0) & (best_score > min_score)
    hit[hit] &= truthy[best_idx[hit]]
    which = best_idx[hit].tolist()
    reference.update(zip(compress(keys, hit.tolist()), (ref_keys[i] for i in which)))
    for key in compress(keys, (~hit).tolist()):
        logging.warning("\nWARNING NO MATCH FOR {0}\n Use ambivert --fastqamplicon {1} to retrieve reads from this amplicon".format(
            merged[key], key))


def align_to_reference(self, global_align=True):
    """ambivert.py:543-574 with one batched alignment (+ traceback) for all matched amplicons"""
    keys = [k for k in self.reference if k in self.merged]
    if not keys:
        return
    ref_keys = [self.reference[k] for k in keys]
    queries = [self.merged[k] for k in keys]
    refs = [self.reference_sequences[rk] for rk in ref_keys]
    minus = [i for i, rk in enumerate(ref_keys) if rk[-1] == "-"]      # minus-strand probes: both sequences to the plus strand
    if minus:
        rc = reverse_complement_many([queries[i] for i in minus] + [refs[i] for i in minus])
        for j, i in enumerate(minus):
            queries[i] = rc[j]
            refs[i] = rc[len(minus) + j]
    t1, t2, offs, _st1, st2 = batched_pairwise(queries, refs, not global_align, raw=True)
    t2_upper = t2.upper()                           # ref.upper() != sample (ambivert.py:566) on one upper-cased text, not per row
    place = {}                                      # (chromosome, start, end) of a target, converted once per target
    aligned, location, potential = self.aligned, self.location, self.potential_variants
    for i, k in enumerate(keys):
        a, b = offs[i], offs[i + 1]
        sample = t1[a:b]
        if t2_upper[a:b] != sample:
            potential.append(k)
        aligned[k] = (sample, t2[a:b])
        rk = ref_keys[i]
        loc = place.get(rk)
        if loc is None:
            loc = place[rk] = (rk[1], int(rk[2]), int(rk[3]))
        location[k] = loc + (st2[i] if st2 is not None else 0,)


def _reference_sha224(self):
    return hashlib.sha224(repr(sorted(self.reference_sequences)).encode("latin-1")).hexdigest()


def save_hash_table(self, newhashfile):
    """the on-disk product of match_to_reference, byte-compatible with ambivert.py:608-619"""
    with newhashfile as outfile:
        table = {k: self.reference[k] for k in self.reference if k not in self.potential_variants}
        pickle.dump((_reference_sha224(self), table), outfile)


def load_hash_table(self, hashfile):
    """ambivert.py:621-637"""
    with hashfile as infile:
        sha, table = pickle.load(infile)
        if sha == _reference_sha224(self):
            self.reference = table
        else:
            warnings.warn("WARNING: loaded read to reference hash library does not match reference sequences\n"
                          "I really hope you know what you are doing... Check and if in doubt use --savehashtable\n"
                          "without specifying an existing hash file.", UserWarning)


def _parse_fastq(handle):
    """(name, sequence, quality) of every 4-line FASTQ record, name without its leading character; the same
    non-validating reading as the reference's parse_fastq (sequence_utilities.py:51-69): lines end at '\n'
    only, bytes are latin-1, the file ends at the first empty name line"""
    with handle as fastqfile:
        text = fastqfile.read()
    if isinstance(text, bytes):
        text = text.decode("latin-1")
    lines = text.split("\n")
    for i in range(0, len(lines), 4):
        name = lines[i]
        if not name:
            return
        seq = lines[i + 1] if i + 1 < len(lines) else ""
        qual = lines[i + 3] if i + 3 < len(lines) else ""
        yield (name[1:], seq, qual)


def _from_variants(name):
    """method of ambivert_b200.variants, looked up at call time (that module imports this one)"""
    def method(self, *args, **kwargs):
        from . import variants
        return getattr(variants, name)(self, *args, **kwargs)
    method.__name__ = name
    return method


class AmpliconData(object):
    """The state the hot path reads and writes (ambivert.py:202-228) with the on-path stages bound to
    the batched GPU implementations, and the steps either side of it (clustering, variant calling and
    consolidation: SURVEY.md section 8f).  Manifest parsing and the SAM/FASTQ writers stay with the reference."""

    def __init__(self, trim5=None, trim3=None):
        self.data = defaultdict(list)      # {md5hexdigest: [((f_name, f_seq, f_qual), (r_name, r_seq, r_qual)), ...]}
        self.merged = {}
        self.unmergable = []
        self.aligned = {}
        self.location = {}
        self.reference = {}
        self.reference_sequences = {}      # {(name, chromosome, start, end, strand): sequence}
        self.potential_variants = []
        self.called_variants = {}          # {md5hexdigest: [AmpliconVariant, ...]}
        self.consolidated_variants = []
        self.trim5 = trim5
        self.trim3 = None if trim3 is None else -1 * abs(trim3)
        self.threshold = 0
        self.readpairs = 0

    def add_reads(self, f_name, f_seq, f_qual, r_name, r_seq, r_qual):
        """the cluster key is the md5 of the (trimmed) forward + reverse read (ambivert.py:259)"""
        key = hashlib.md5(bytes(f_seq[self.trim5:self.trim3] + r_seq[self.trim5:self.trim3], "ascii")).hexdigest()
        self.data[key].append(((f_name, f_seq, f_qual), (r_name, r_seq, r_qual)))
        self.readpairs += 1

    def add_read_pairs(self, pairs):
        """Bulk add_reads: `pairs` is a sequence of ((f_name, f_seq, f_qual), (r_name, r_seq, r_qual)).  The md5
        keys and the grouping are computed on the device in one call (abv_cluster_pairs); self.data ends up exactly
        as after calling add_reads on every pair in order (same keys, same insertion order, same lists)."""
        pairs = list(pairs)
        if not pairs:
            return
        cluster_of, keys, _count, _first = B.cluster_pairs(_ascii(p[0][1] for p in pairs), _ascii(p[1][1] for p in pairs),
                                                            self.trim5, self.trim3)
        for pair, k in zip(pairs, cluster_of.tolist()):
            self.data[keys[k]].append((tuple(pair[0]), tuple(pair[1])))
        self.readpairs += len(pairs)

    def process_twofile_readpairs(self, forward_file, reverse_file, parse_fastq=None):
        """ambivert.py:264-289 with the keying done in bulk: reads both FASTQ files (4-line records, text or
        bytes) and adds the pairs in file order."""
        parse = parse_fastq or _parse_fastq
        self.add_read_pairs(zip(parse(forward_file), parse(reverse_file)))

    get_above_threshold = get_above_threshold
    merge_overlaps = merge_overlaps
    match_to_reference = match_to_reference
    align_to_reference = align_to_reference
    save_hash_table = save_hash_table
    load_hash_table = load_hash_table
    get_amplicon_count = _from_variants("get_amplicon_count")
    call_amplicon_variants = _from_variants("call_amplicon_variants")
    get_variant_positions = _from_variants("get_variant_positions")
    consolidate_variants = _from_variants("consolidate_variants")
    get_filtered_variants = _from_variants("get_filtered_variants")
    print_consolidated_vcf = _from_variants("print_consolidated_vcf")
    get_amplicons_overlapping = _from_variants("get_amplicons_overlapping")
    get_reference_overlapping = _from_variants("get_reference_overlapping")
    get_amplicons_overlapping_without_softmasking = _from_variants("get_amplicons_overlapping_without_softmasking")
    is_homopolymer_at_position = _from_variants("is_homopolymer_at_position")
