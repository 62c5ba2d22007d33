"""Synthetic amplicon data of the shapes BASELINE.json names.  Seeded and
deterministic, so every rank of a multi-GPU run builds the same set and takes its own shard."""
import random

import numpy as np

_ACGT = np.frombuffer(b"ACGT", np.uint8)
_IUPAC = b"NRYSWKM"
_COMP = bytes.maketrans(b"ACGTNacgtn", b"TGCANtgcan")


def reverse_complement(seq):
    return seq.translate(_COMP)[::-1]


def _random_dna(rng, n):
    return _ACGT[rng.integers(0, 4, int(n))].tobytes()


def panel_targets(n_targets=2000, lo=200, hi=260, probe=27, seed=0xC4):
    """manifest-style targets: upper-case core with lower-case primer flanks, as
    truseq_manifest.make_sequences(with_probes, softmask_probes) produces (truseq_manifest.py:148-152)"""
    rng = np.random.default_rng(seed)
    out = []
    for L in rng.integers(lo, hi + 1, n_targets):
        s = _random_dna(rng, L)
        out.append(s[:probe].lower() + s[probe:len(s) - probe] + s[len(s) - probe:].lower())
    return out


def panel_queries(targets, n_queries=100000, seed=0xC4 + 1, snv_mean=1.5, p_indel=0.10, p_iupac=0.005, p_unrelated=0.01):
    """unique amplicons: mutated copies of uniformly chosen targets (upper-cased), k~Poisson SNVs, an
    occasional 1-5 bp indel, a rare IUPAC/N base, 1 % unrelated sequences; de-duplicated.
    -> (queries, source target index or -1)"""
    rng = np.random.default_rng(seed)
    up = [t.upper() for t in targets]
    seen = set()
    queries, source = [], []
    while len(queries) < n_queries:
        m = n_queries - len(queries)
        src = rng.integers(0, len(up), m)
        nsnv = rng.poisson(snv_mean, m)
        indel = rng.random(m) < p_indel
        iupac = rng.random(m) < p_iupac
        unrelated = rng.random(m) < p_unrelated
        for i in range(m):
            if unrelated[i]:
                s = bytearray(_random_dna(rng, len(up[src[i]])))
                origin = -1
            else:
                s = bytearray(up[src[i]])
                origin = int(src[i])
                for p in rng.integers(0, len(s), int(nsnv[i])):
                    s[p] = b"ACGT"[int(rng.integers(0, 4))]
                if indel[i]:
                    p = int(rng.integers(1, len(s) - 6))
                    n = int(rng.integers(1, 6))
                    if rng.random() < 0.5:
                        del s[p:p + n]
                    else:
                        s[p:p] = _random_dna(rng, n)
                if iupac[i]:
                    s[int(rng.integers(0, len(s)))] = _IUPAC[int(rng.integers(0, len(_IUPAC)))]
            b = bytes(s)
            if b in seen:
                continue
            seen.add(b)
            queries.append(b)
            source.append(origin)
    return queries, np.asarray(source, np.int64)


def fasta_targets(n_targets=10000, length=250, seed=0xD5):
    rng = np.random.default_rng(seed)
    return [_random_dna(rng, length) for _ in range(n_targets)]


def read_pairs(n_pairs=1000000, read_len=150, seed=0xA3B1, p_sub=0.005, p_n=0.001, p_unrelated=0.02):
    """forward / reverse read pairs of amplicons 170..280 bp (overlap 20..130 bp): F = first 150 bases,
    R = reverse complement of the last 150; per-base substitutions and N calls; 2 % of pairs carry an
    unrelated reverse read (the `unmergable` path, ambivert.py:324-325).  -> (fwd, rev) as sequenced"""
    rng = np.random.default_rng(seed)
    lens = rng.integers(read_len + 20, 2 * read_len - 19, n_pairs)
    unrelated = rng.random(n_pairs) < p_unrelated
    fwd, rev = [], []
    for i in range(n_pairs):
        amp = _ACGT[rng.integers(0, 4, int(lens[i]))]
        f = amp[:read_len].copy()
        r = (_ACGT[rng.integers(0, 4, read_len)] if unrelated[i] else amp[-read_len:].copy())
        for a in (f, r):
            nsub = rng.binomial(read_len, p_sub)
            if nsub:
                a[rng.integers(0, read_len, nsub)] = _ACGT[rng.integers(0, 4, nsub)]
            nn = rng.binomial(read_len, p_n)
            if nn:
                a[rng.integers(0, read_len, nn)] = ord("N")
        fwd.append(f.tobytes())
        rev.append(reverse_complement(r.tobytes()))
    return fwd, rev


def read_pairs_packed(n_pairs=1000000, read_len=150, seed=0xA3B1, p_sub=0.005, p_n=0.001, p_unrelated=0.02):
    """the read-pair workload of read_pairs() generated with array operations and returned already
    packed for the C ABI: ((fwd_buffer, offsets), (rev_buffer, offsets)), every read `read_len` long.
    (Same distribution, different random stream: use read_pairs() where golden results are pinned.)"""
    rng = np.random.default_rng(seed)
    max_len = 2 * read_len - 20
    lens = rng.integers(read_len + 20, max_len + 1, n_pairs)
    amp = _ACGT[rng.integers(0, 4, (n_pairs, max_len), dtype=np.uint8)]
    fwd = amp[:, :read_len].copy()
    cols = (lens - read_len)[:, None] + np.arange(read_len)[None, :]
    rev = np.take_along_axis(amp, cols, axis=1)                      # last read_len bases of each amplicon (plus strand)
    unrelated = rng.random(n_pairs) < p_unrelated
    rev[unrelated] = _ACGT[rng.integers(0, 4, (int(unrelated.sum()), read_len), dtype=np.uint8)]
    for a in (fwd, rev):
        sub = rng.random(a.shape) < p_sub
        a[sub] = _ACGT[rng.integers(0, 4, int(sub.sum()), dtype=np.uint8)]
        a[rng.random(a.shape) < p_n] = ord("N")
    off = np.arange(n_pairs + 1, dtype=np.int64) * read_len
    return (np.ascontiguousarray(fwd).reshape(-1), off), (np.ascontiguousarray(rev).reshape(-1), off.copy())


def panel_pipeline_inputs(n_targets=300, n_unique=20000, read_len=150, seed=0xB2CA):
    """BASELINE.json configs[2] shape: a TruSeq-style panel and the unique read pairs of its amplicons.
    -> (references: list of ((name, chrom, start, end, strand), sequence str),
        pairs: list of (forward read, reverse read as sequenced, copies))
    Targets are 170..250 bp with soft-masked 27 bp probes, alternating strand; amplicons are mutated copies
    (SNVs, occasional indel, rare IUPAC base); F = first read_len bases, R = reverse complement of the last."""
    rng = np.random.default_rng(seed + 7)
    targets = panel_targets(n_targets, 170, 250, seed=seed)
    refs = []
    for i, t in enumerate(targets):
        refs.append((("target%03d" % i, "chr%d" % (1 + i % 2), str(1000 * i), str(1000 * i + len(t)), "+" if i % 2 == 0 else "-"), t.decode()))
    # reads carry ACGT only: two different IUPAC codes facing each other make the reference's own
    # encode_ambiguous raise KeyError (sequence_utilities.py:152-156)
    queries, src = panel_queries(targets, n_unique, seed=seed + 1, p_unrelated=0.0, p_iupac=0.0)
    pairs = []
    for q in queries:
        if len(q) < read_len:
            continue
        pairs.append((q[:read_len].decode(), reverse_complement(q[-read_len:]).decode(), int(rng.integers(1, 4))))
    return refs, pairs


def aligned_rows(n, seed=0xF4):
    """aligned rows of config-4-like amplicons: 27 bp lower-case primer flanks, ~200 bp core, a few events each"""
    rng = random.Random(seed)
    rows = []
    for _ in range(n):
        core = rng.randint(150, 210)
        ref = [rng.choice("acgt") for _ in range(27)] + [rng.choice("ACGT") for _ in range(core)] + [rng.choice("acgt") for _ in range(27)]
        s, r = [], []
        for b in ref:
            x = rng.random()
            if x < 0.004:
                s.append("-"); r.append(b)
            elif x < 0.008:
                s.append(rng.choice("ACGT")); r.append("-"); s.append(b.upper()); r.append(b)
            elif x < 0.02:
                s.append(rng.choice("ACGTN")); r.append(b)
            else:
                s.append(b.upper()); r.append(b)
        rows.append(("".join(s), "".join(r)))
    return rows
