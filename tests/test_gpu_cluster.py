"""GPU parity of the clustering front end: md5 keys of read pairs + grouping on the
device against hashlib (what the reference uses, ambivert.py:259) and against the keys the reference produced
for its bundled FASTQ data (tests/golden/pipeline_bundled.json)."""
import hashlib
import io
import random

import numpy as np
import pytest

from util import golden, rand_dna

pytestmark = pytest.mark.gpu


def ref_key(f, r, trim5, trim3):
    return hashlib.md5(bytes(f[trim5:trim3] + r[trim5:trim3], "ascii")).hexdigest()


def check(B, fwds, revs, trim5, trim3):
    cluster_of, keys, count, first = B.cluster_pairs([s.encode() for s in fwds], [s.encode() for s in revs], trim5, trim3)
    want = {}
    for i, (f, r) in enumerate(zip(fwds, revs)):
        want.setdefault(ref_key(f, r, trim5, trim3), []).append(i)
    assert keys == list(want)                                   # same keys in order of first occurrence
    assert [int(c) for c in count] == [len(v) for v in want.values()]
    assert [int(x) for x in first] == [v[0] for v in want.values()]
    for k, members in enumerate(want.values()):
        assert (cluster_of[members] == k).all()


def test_md5_known_answers():
    from ambivert_b200 import batch as B
    # RFC 1321 test suite, through the pair interface (forward + reverse concatenated)
    cases = [("", "", "d41d8cd98f00b204e9800998ecf8427e"), ("a", "", "0cc175b9c0f1b6a831c399e269772661"), ("a", "bc", "900150983cd24fb0d6963f7d28e17f72"),
             ("message ", "digest", "f96b697d7cb7938d525a2f31aaf161d0"), ("abcdefghijklm", "nopqrstuvwxyz", "c3fcd3d76192e4007dfb496cca67e13b"),
             ("ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz", "0123456789", "d174ab98d277d9f5a5611c2c9f419d9f"),
             ("1234567890" * 4, "1234567890" * 4, "57edf4a22be3c955ac49da2e2107b67a")]
    cluster_of, keys, count, first = B.cluster_pairs([c[0].encode() for c in cases], [c[1].encode() for c in cases])
    assert keys == [c[2] for c in cases] and list(cluster_of) == list(range(len(cases)))


def test_cluster_pairs_random_vs_hashlib():
    from ambivert_b200 import batch as B
    rng = random.Random(77)
    uniq = [(rand_dna(rng, rng.choice([0, 1, 27, 28, 55, 56, 57, 63, 64, 65, 119, 120, 150, 151, 300])),
             rand_dna(rng, rng.choice([0, 1, 30, 64, 150, 151, 250]))) for _ in range(400)]
    pairs = [uniq[rng.randrange(len(uniq))] for _ in range(5000)]
    fwds, revs = [p[0] for p in pairs], [p[1] for p in pairs]
    for trim5, trim3 in ((None, None), (0, None), (5, -7), (10, -200), (-20, None), (400, -1), (None, -3)):
        check(B, fwds, revs, trim5, trim3)
    check(B, [], [], None, None)
    check(B, ["ACGT"], ["TTGA"], 1, -1)


def test_bundled_fastq_keys_and_bulk_add_reads():
    """the keys of the reference's bundled read pairs, and AmpliconData.add_read_pairs == repeated add_reads"""
    from ambivert_b200.host import AmpliconData
    G = golden("pipeline_bundled.json")
    reads = []
    rng = random.Random(1)
    for p in G["pairs"]:
        reads += [(("n", p["fwd"], "I" * len(p["fwd"])), ("n", p["rev"], "I" * len(p["rev"])))] * p["count"]
    rng.shuffle(reads)
    one_by_one, bulk = AmpliconData(), AmpliconData()
    for f, r in reads:
        one_by_one.add_reads(*f, *r)
    bulk.add_read_pairs(reads)
    assert list(bulk.data) == list(one_by_one.data) and bulk.data == one_by_one.data and bulk.readpairs == one_by_one.readpairs
    assert set(bulk.data) == {p["key"] for p in G["pairs"]}      # the reference's own md5 keys
    assert {k: len(v) for k, v in bulk.data.items()} == {p["key"]: p["count"] for p in G["pairs"]}
    # FASTQ text in, same clustering out
    fq = lambda idx: io.BytesIO("".join("@r%d\n%s\n+\n%s\n" % (i, rd[idx][1], rd[idx][2]) for i, rd in enumerate(reads)).encode())
    parsed = AmpliconData()
    parsed.process_twofile_readpairs(fq(0), fq(1))
    assert list(parsed.data) == list(bulk.data) and [len(v) for v in parsed.data.values()] == [len(v) for v in bulk.data.values()]
    trimmed, trimmed_ref = AmpliconData(trim5=3, trim3=4), AmpliconData(trim5=3, trim3=4)
    trimmed.add_read_pairs(reads[:500])
    for f, r in reads[:500]:
        trimmed_ref.add_reads(*f, *r)
    assert trimmed.data == trimmed_ref.data and list(trimmed.data) == list(trimmed_ref.data)
