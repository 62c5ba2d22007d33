"""GPU parity of the device-side string assembly (merged reads, gapped rows) against what the reference
PYTHON + C produce (tests/golden/merge_strings.json.gz, made by oracle/make_golden_merge.py):
smith_waterman / needleman_wunsch rows (ambivert.py:75-155) and the merged read of merge_overlaps
(ambivert.py:320-327 with sequence_utilities.flatten_paired_alignment / encode_ambiguous)."""
import gzip
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "merge_strings.json.gz")


@pytest.fixture(scope="module")
def G():
    with gzip.open(GOLD, "rt") as f:
        return json.load(f)


def test_merged_reads_match_reference_python(G):
    from ambivert_b200 import batch as B
    pairs = G["pairs"]
    sc, status, off, merged = B.merge_pairs([p["fwd"] for p in pairs], [p["rev"] for p in pairs], G["minimum_overlap"])
    text = merged.tobytes().decode("ascii")
    n_un = 0
    for k, p in enumerate(pairs):
        if p["merged"] is None:
            assert status[k] == 0 and off[k + 1] == off[k], k
            n_un += 1
        else:
            assert status[k] == 1 and text[off[k]:off[k + 1]] == p["merged"], k
    assert n_un > 10 and n_un < len(pairs) // 2


def test_merge_overlaps_host_stage(G):
    """through the host mirror: keys, merged dict, unmergable list in data order"""
    from ambivert_b200.host import AmpliconData, reverse_complement
    amp = AmpliconData()
    want_merged, want_un = {}, []
    for i, p in enumerate(G["pairs"]):
        key = "k%04d" % i
        amp.data[key] = [(("r", p["fwd"], ""), ("r", reverse_complement(p["rev"]), ""))] * (1 + i % 3)
        if i % 3 == 0:
            continue                         # count 1 is not above threshold 1
        if p["merged"] is None:
            want_un.append(key)
        else:
            want_merged[key] = p["merged"]
    amp.merge_overlaps(threshold=1, minimum_overlap=G["minimum_overlap"])
    assert amp.merged == want_merged and list(amp.merged) == list(want_merged)
    assert amp.unmergable == want_un


def test_gapped_rows_match_reference_python(G):
    from ambivert_b200 import host as H
    al = G["aligns"]
    qs, refs = [a["q"] for a in al], [a["ref"] for a in al]
    for tag, local in (("nw", False), ("sw", True)):
        got = H.batched_pairwise(qs, refs, local)
        for k, a in enumerate(al):
            assert list(got[k]) == a[tag], (tag, k)


def test_bad_base_pair_is_reported():
    """two different non-ACGTN bases facing each other: encode_ambiguous raises KeyError in the reference"""
    from ambivert_b200 import batch as B
    from ambivert_b200.host import AmpliconData, reverse_complement
    fwd = "ACGTACGTACGTACGTACGTAGGCTTACGATCGATCGGATCRATCGATTAGC"
    rev = "ACGTACGTACGTACGTACGTAGGCTTACGATCGATCGGATCYATCGATTAGC"     # R vs Y at one column
    sc, status, off, merged = B.merge_pairs([fwd], [rev], 10)
    assert status[0] == -2
    amp = AmpliconData()
    amp.data["k"] = [(("r", fwd, ""), ("r", reverse_complement(rev), ""))]
    with pytest.raises(KeyError):
        amp.merge_overlaps(threshold=0, minimum_overlap=10)
    # identical ambiguity codes and N are fine: same base -> itself, anything vs N -> N
    sc, status, off, merged = B.merge_pairs([fwd], [fwd.replace("R", "N")], 10)
    assert status[0] == 1 and merged.tobytes().decode() == fwd.replace("R", "N")


def test_chunk_pipeline_equals_one_piece():
    """large one-to-one jobs run as a pipeline of chunks on alternating streams (uploads and downloads beside the kernels);
    forced here on small batches: scores, fragments, gapped rows and merged reads must equal the one-piece run, including
    pairs the packed traceback kernel cannot take (long sequences) and a fragment pool that is too small"""
    import random

    import numpy as np

    from ambivert_b200 import batch as B
    from util import mutate, rand_dna
    rng = random.Random(99)
    fwd, rev = [], []
    for _ in range(700):
        f = rand_dna(rng, rng.randint(20, 200))
        r = mutate(rng, f[rng.randint(0, len(f) - 10):] + rand_dna(rng, rng.randint(5, 80)), rng.randint(0, 3))
        fwd.append(f.encode()); rev.append(r.encode())
    for _ in range(6):                                   # too long for the packed traceback kernel
        f = rand_dna(rng, rng.randint(400, 700))
        fwd.append(f.encode()); rev.append(mutate(rng, f, 4).encode())
    order = list(range(len(fwd)))
    rng.shuffle(order)
    fwd = [fwd[i] for i in order]; rev = [rev[i] for i in order]

    def frag_lists(res):
        sc, st, cn, fr = res
        return [(int(sc[k]), [tuple(int(x) for x in fr[int(st[k]) + j]) for j in range(int(cn[k]))]) for k in range(len(sc))]

    def run_all():
        out = {}
        for mode in (B.LOCAL, B.GLOBAL):
            out["align", mode] = frag_lists(B.align_pairs(fwd, rev, mode))
            g = B.gapped_pairs(fwd, rev, mode)
            out["gapped", mode] = [x.tolist() if hasattr(x, "tolist") else x for x in g]
        out["merge"] = [x.tolist() for x in B.merge_pairs(fwd, rev, 10)]
        return out

    B.set_option("pairs_chunk", 0)
    try:
        want = run_all()
        for per in (37, 150, 353):
            B.set_option("pairs_chunk", per)
            got = run_all()
            for key in want:
                assert got[key] == want[key], (per, key)
        # a fragment pool that is too small: the same error and the same required size as the one-piece run
        B.set_option("pairs_chunk", 0)
        with pytest.raises(B.AlignError) as e0:
            B.align_pairs(fwd, rev, B.LOCAL, frag_cap=50)
        B.set_option("pairs_chunk", 100)
        with pytest.raises(B.AlignError) as e1:
            B.align_pairs(fwd, rev, B.LOCAL, frag_cap=50)
        assert str(e0.value) == str(e1.value)
    finally:
        B.set_option("pairs_chunk", 200000)


def test_align_pairs_into_pinned_result_buffers():
    """align_pairs(out=align_pairs_buffers(...)): same results as the plain call, returned as views of the page-locked buffers;
    a fragment pool that is too small raises ERR_FRAG_CAP with scores and counts valid"""
    import random
    from ambivert_b200 import batch as B
    rng = random.Random(31)
    a = ["".join(rng.choice("ACGT") for _ in range(rng.randint(20, 160))).encode() for _ in range(300)]
    b = [x[rng.randrange(len(x) // 2):] + "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 40))).encode() for x in a]
    want = B.align_pairs(a, b, B.LOCAL)
    bufs = B.align_pairs_buffers(len(a) + 5, 8 * len(a))
    got = B.align_pairs(a, b, B.LOCAL, out=bufs)
    assert (got[0] == want[0]).all() and (got[2] == want[2]).all() and got[3].shape == want[3].shape
    for k in range(len(a)):            # a pair's place in the fragment pool depends on the order the warps finish in
        assert B.fragments(k, got[1], got[2], got[3]) == B.fragments(k, want[1], want[2], want[3]), k
    assert got[0].ctypes.data == bufs[0].ctypes.data and got[3].ctypes.data == bufs[3].ctypes.data
    small = B.align_pairs_buffers(len(a), 10)
    with pytest.raises(B.AlignError) as e:
        B.align_pairs(a, b, B.LOCAL, out=small)
    assert e.value.code == B.ERR_FRAG_CAP
    assert (small[0][:len(a)] == want[0]).all() and (small[2][:len(a)] == want[2]).all()


def test_pipelined_jobs_reject_a_bad_late_pair_like_the_one_piece_run():
    """the pairs of the later chunks are validated after the first two chunks have been sent (StagedValidation, align_abi.cu): a bad
    pair anywhere gives the same error as the one-piece run, names the same pair, and nothing is written to the result buffers"""
    import random
    from ambivert_b200 import batch as B
    rng = random.Random(9)
    n = 3000
    a = np.frombuffer(bytes(rng.choice(b"ACGT") for _ in range(40 * n)), np.uint8).copy()
    off = np.arange(n + 1, dtype=np.int64) * 40
    for bad in (5, 2950):                                   # inside the first chunk / inside the last one
        boff = off.copy()
        boff[bad + 1:] -= 40                                # pair `bad` gets an empty second sequence
        b = a[:int(boff[-1])]
        msgs = []
        for per in (0, 250):
            B.set_option("pairs_chunk", per)
            try:
                bufs = B.align_pairs_buffers(n, 4 * n)
                for x in bufs:
                    x.fill(-7)
                with pytest.raises(B.AlignError) as e:
                    B.align_pairs((a, off), (b, boff), B.LOCAL, out=bufs)
                assert e.value.code == B.ERR_ARGS and ("sequence b %d " % bad) in str(e.value)
                assert all((x == -7).all() for x in bufs)
                with pytest.raises(B.AlignError) as e2:
                    B.merge_pairs((a, off), (b, boff), 10)
                assert e2.value.code == B.ERR_ARGS
                msgs.append((str(e.value), str(e2.value)))
            finally:
                B.set_option("pairs_chunk", 200000)
        assert msgs[0] == msgs[1]
