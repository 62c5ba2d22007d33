"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle and the golden
vectors.  Bit-exact: scores, fragment lists, chosen targets (all arithmetic is integer)."""
import ctypes as C
import random

import numpy as np
import pytest

import oracle as O
from util import as_result, case_scoring, gapped_strings, golden, mutate, rand_dna

pytestmark = pytest.mark.gpu

MODES = [O.OVERLAP, O.LOCAL, O.GLOBAL, O.GLOCAL]


@pytest.fixture(scope="module")
def B():
    from ambivert_b200 import batch
    assert batch.device_count() >= 1, "no CUDA device: the CUDA path cannot be tested"
    batch.set_option("path", batch.PATH_AUTO)
    return batch


@pytest.fixture(scope="module")
def A():
    import ambivert_b200.align as align
    return align


def legacy(A, mode, s1, s2, alpha=None, mp=None, M=None, go=-7, ge=-1):
    """one call through the reference's own ABI -> (score, frags) or None for NULL"""
    fn = {O.OVERLAP: A.align, O.LOCAL: A.local_align, O.GLOBAL: A.global_align, O.GLOCAL: A.glocal_align}[mode]
    alpha = A.DNA_MAP[0] if alpha is None else alpha
    mp = A.DNA_MAP[1] if mp is None else (C.c_ubyte * 256)(*[int(x) for x in mp])
    M = A.DNA_SCORE if M is None else (C.c_int * (alpha * alpha))(*[int(x) for x in M])
    al = fn(s1, len(s1), s2, len(s2), alpha, mp, M, go, ge)
    if not al:
        return None
    score = al[0].score
    out = []
    f = al[0].align_frag
    while f:
        f = f[0]
        out.append((f.type, f.sa_start, f.sb_start, f.hsp_len))
        f = f.next
    assert len(out) == al[0].frag_count
    A.alignment_free(al)
    return score, out


# ------------------------------------------------------------------------------------------------
def test_legacy_abi_reference_kats(A):
    """the known answers of the reference's own tests (test_ambivert.py:111-192), via the 8 legacy symbols"""
    n = 0
    for c in golden("kat_reference_tests.json"):
        for m in MODES:
            want = as_result(c["modes"][O.MODE_NAMES[m]])
            got = legacy(A, m, c["s1"].encode(), c["s2"].upper().encode())
            if isinstance(want, int):
                assert got is None, (c["name"], O.MODE_NAMES[m])
            else:
                assert got == want, (c["name"], O.MODE_NAMES[m])
                n += 1
        for tag, mode in (("sw", O.LOCAL), ("nw", O.GLOBAL)):
            exp = c["expect"].get(tag)
            if not exp:
                continue
            score, frags = legacy(A, mode, c["s1"].encode(), c["s2"].upper().encode())
            a1, a2 = gapped_strings(c["s1"], c["s2"], frags)
            if "aligned_s1" in exp:
                assert a1 == exp["aligned_s1"]
            if "aligned_s2" in exp:
                assert a2 == exp["aligned_s2"]
    assert n >= 60


def test_legacy_abi_raw_and_errors(A):
    M = A.DNA_SCORE
    enc = lambda s: O.encode(s)
    al = A.global_align_raw(enc("ACGT"), 4, enc("A"), 1, 5, M, -7, -1)
    assert al and al[0].score == -8 and al[0].frag_count == 2
    A.alignment_free(al)
    al = A.local_align_raw(enc("AAAA"), 4, enc("CCCC"), 4, 5, M, -7, -1)
    assert al and al[0].score == 0 and al[0].frag_count == 0 and not al[0].align_frag
    A.alignment_free(al)
    # what makes the reference return NULL (global_align.c:118-122)
    assert not A.global_align(b"ACGT", 4, b"ACGT", 4, 5, A.DNA_MAP[1], M, 1, -1)
    assert not A.global_align(b"ACGT", 4, b"ACGT", 4, 5, A.DNA_MAP[1], M, -7, 1)
    assert not A.global_align(b"", 0, b"ACGT", 4, 5, A.DNA_MAP[1], M, -7, -1)
    assert not A.local_align(b"ACGT", 4, b"", 0, 5, A.DNA_MAP[1], M, -7, -1)
    # undefined in the reference -> NULL here
    assert not A.glocal_align(b"ACGT", 4, b"A", 1, 5, A.DNA_MAP[1], M, -7, -1)
    assert not A.align(b"A", 1, b"ACGT", 4, 5, A.DNA_MAP[1], M, -7, -1)


def test_golden_fuzz_vectors_through_batched_abi(B):
    """outputs of the compiled reference (tests/golden/fuzz_ref_vectors.json), all four modes"""
    g = golden("fuzz_ref_vectors.json")
    groups = {}
    for i, c in enumerate(g["cases"]):
        key = (c["alphabet"], c["unknown"], str(c["matrix"]), c["go"], c["ge"])
        groups.setdefault(key, []).append(i)
    checked = 0
    for key, idxs in groups.items():
        c0 = g["cases"][idxs[0]]
        alpha, mp, M = case_scoring(c0)
        for m in MODES:
            use = [i for i in idxs if not isinstance(as_result(g["cases"][i]["modes"][O.MODE_NAMES[m]]), int)]
            if not use:
                continue
            a = [g["cases"][i]["a"].encode() for i in use]
            b = [g["cases"][i]["b"].encode() for i in use]
            sc, st, cn, fr = B.align_pairs(a, b, m, alpha, mp, M, c0["go"], c0["ge"])
            for k, i in enumerate(use):
                want = as_result(g["cases"][i]["modes"][O.MODE_NAMES[m]])
                assert (int(sc[k]), B.fragments(k, st, cn, fr)) == want, (i, O.MODE_NAMES[m])
                checked += 1
    assert checked > 1500


@pytest.mark.parametrize("mode", MODES)
def test_pairs_live_fuzz_vs_oracle(B, oracle_port, mode):
    """one-to-one alignments with traceback on fresh random inputs: ragged lengths 1..700, N/IUPAC,
    repeats (ties), mutated copies; several scoring schemes"""
    rng = random.Random(1000 + mode)
    for scheme in range(4):
        if scheme == 0:
            alpha, mp, M, go, ge, alphabet = 5, O.DNA_MAP, O.DNA_SCORE, -7, -1, "ACGT"
        elif scheme == 1:
            alpha, mp, M, go, ge, alphabet = 5, O.DNA_MAP, O.DNA_SCORE, -7, -1, "AAAC"          # repeats: many ties
        elif scheme == 2:
            alpha, mp, M, go, ge, alphabet = 5, O.DNA_MAP, O.dna_matrix(2, -3, 0), 0, -2, "ACGTNRacgt"   # go > ge
        else:
            alpha = 24
            _, mp = O.make_map("ARNDCQEGHILKMFPSTWYVBZX*", "X", True)
            M = np.array([rng.randint(-4, 4) if a != b else rng.randint(4, 11) for a in range(24) for b in range(24)], np.int32)
            go, ge, alphabet = -11, -1, "ARNDCQEGHILKMFPSTWYV"
        a_list, b_list = [], []
        for _ in range(160):
            la = rng.choice([1, 2, 3, 31, 32, 33, 64, rng.randint(4, 90), rng.randint(100, 320), rng.randint(100, 320)])
            a = rand_dna(rng, la, alphabet)
            if rng.random() < 0.6:
                b = mutate(rng, a, rng.randint(0, 8), alphabet)
                if rng.random() < 0.3:
                    b = b[rng.randint(0, len(b) // 2):]
                if rng.random() < 0.3:
                    b = rand_dna(rng, rng.randint(1, 40), alphabet) + b
            else:
                b = rand_dna(rng, rng.choice([1, 2, 3, rng.randint(4, 320)]), alphabet)
            if (mode == O.GLOCAL and len(b) < 2) or (mode == O.OVERLAP and len(a) < 2):
                continue
            a_list.append(a.encode()); b_list.append(b.encode())
        a_list.append(rand_dna(rng, 700, alphabet).encode()); b_list.append(rand_dna(rng, 650, alphabet).encode())
        sc, st, cn, fr = B.align_pairs(a_list, b_list, mode, alpha, mp, M, go, ge)
        for k in range(len(a_list)):
            want = oracle_port.align(mode, a_list[k], b_list[k], alpha, mp, M, go, ge)
            assert (int(sc[k]), B.fragments(k, st, cn, fr)) == want, (scheme, k, a_list[k], b_list[k])


def _oracle_score_matrix(port, mode, q, t, alpha, M, go, ge, threads=8):
    qa = [x for x in q for _ in t]
    tb = [y for _ in q for y in t]
    _, sc = port.timed_pairs(mode, qa, tb, alpha, M, go, ge, threads)
    return sc.reshape(len(q), len(t))


def _amplicon_set(rng, n_t, n_q, lo, hi, n_rate=0.01):
    targets = [rand_dna(rng, rng.randint(lo, hi)) for _ in range(n_t)]
    queries = []
    for _ in range(n_q):
        if rng.random() < 0.9:
            s = mutate(rng, targets[rng.randrange(n_t)], rng.randint(0, 4))
        else:
            s = rand_dna(rng, rng.randint(lo, hi))
        if rng.random() < n_rate * 10:
            p = rng.randrange(len(s)); s = s[:p] + rng.choice("NRYS") + s[p + 1:]
        queries.append(s)
    return queries, targets


@pytest.mark.parametrize("mode", [O.GLOBAL, O.GLOCAL, O.LOCAL, O.OVERLAP])
def test_score_matrix_packed_and_wf32_vs_oracle(B, oracle_port, mode):
    """every (query, target) score of a many-to-many job: packed 16x2 kernel and 32-bit kernel, both
    against the oracle; query lengths straddle every K bucket boundary (overlap: the packed form of round 2, align.c:68-243)"""
    rng = random.Random(77 + mode)
    queries, targets = _amplicon_set(rng, 37, 60, 180, 262)
    # ... and, for the streaming kernels, the 31K boundaries between the border-lane and the all-32-lanes instantiation of a strip width
    for L in (1, 2, 31, 62, 63, 64, 65, 93, 94, 127, 128, 129, 155, 156, 186, 187, 191, 192, 193, 217, 218, 223, 224, 225, 248, 249, 256, 257, 279, 280, 288, 289,
              300, 310, 311, 320, 321, 372, 373, 383, 384, 385, 496, 497, 500, 512):
        if L == 1 and mode == O.OVERLAP:
            continue                                         # undefined in the reference (align.c:183): the library rejects the job
        queries.append(rand_dna(rng, L))
    targets += [rand_dna(rng, 2), rand_dna(rng, 3), rand_dna(rng, 85), "A" * 85]
    q = [O.encode(s) for s in queries]
    t = [O.encode(s) for s in targets]
    want = _oracle_score_matrix(oracle_port, mode, q, t, 5, O.DNA_SCORE, -7, -1)
    for path in (B.PATH_PACKED, B.PATH_WF32, B.PATH_AUTO):
        B.set_option("path", path)
        try:
            got = B.score_matrix([s.encode() for s in queries], [s.encode() for s in targets], mode)
        finally:
            B.set_option("path", B.PATH_AUTO)
        bad = np.argwhere(got != want)
        assert bad.size == 0, (path, bad[:5], got[tuple(bad[0])], want[tuple(bad[0])], len(queries[bad[0][0]]), len(targets[bad[0][1]]))


def test_score_matrix_overlap_and_long_queries_take_wf32(B, oracle_port):
    rng = random.Random(5)
    queries = [rand_dna(rng, rng.randint(2, 300)) for _ in range(20)] + [rand_dna(rng, 900), rand_dna(rng, 1500)]
    targets = [rand_dna(rng, rng.randint(2, 300)) for _ in range(9)] + [rand_dna(rng, 1200)]
    q = [O.encode(s) for s in queries]
    t = [O.encode(s) for s in targets]
    for mode in (O.OVERLAP, O.GLOBAL):
        want = _oracle_score_matrix(oracle_port, mode, q, t, 5, O.DNA_SCORE, -7, -1)
        got = B.score_matrix([s.encode() for s in queries], [s.encode() for s in targets], mode)
        assert (got == want).all()


def test_best_hits_selection_rule(B, oracle_port):
    """first maximum in target order, strict '>' against the seed (ambivert.py:478-485, :504-511)"""
    rng = random.Random(7)
    targets = [rand_dna(rng, rng.randint(200, 260)) for _ in range(45)]
    targets.insert(30, targets[4])          # duplicates: the lower index must win
    targets.insert(11, targets[40])
    queries = [mutate(rng, targets[rng.randrange(len(targets))], rng.randint(0, 3)) for _ in range(70)]
    queries[0] = targets[4]
    queries[1] = targets[11]
    queries.append("A" * 230)               # nothing scores above 0 -> no hit with seed 0
    q = [O.encode(s) for s in queries]
    t = [O.encode(s) for s in targets]
    for mode, seed in ((O.GLOBAL, -1000000), (O.GLOBAL, 0), (O.GLOCAL, 0), (O.LOCAL, 0), (O.GLOBAL, 150)):
        sm = _oracle_score_matrix(oracle_port, mode, q, t, 5, O.DNA_SCORE, -7, -1)
        want_s = np.empty(len(q), np.int32); want_i = np.empty(len(q), np.int32)
        for i in range(len(q)):
            bs, bi = seed, -1
            for j in range(len(t)):
                if sm[i, j] > bs:
                    bs, bi = sm[i, j], j
            want_s[i], want_i[i] = bs, bi
        for path in (B.PATH_AUTO, B.PATH_WF32):
            B.set_option("path", path)
            try:
                bs, bi = B.best_hits([s.encode() for s in queries], [s.encode() for s in targets], mode, seed)
            finally:
                B.set_option("path", B.PATH_AUTO)
            assert (bs == want_s).all() and (bi == want_i).all(), (mode, seed, path)
    bs, bi = B.best_hits([queries[0].encode(), queries[1].encode()], [s.encode() for s in targets])
    assert list(bi) == [4, 11]


def test_small_alphabets_and_other_penalties_on_the_packed_path(B, oracle_port):
    rng = random.Random(21)
    for alpha, alphabet in ((2, "AC"), (4, "ACGT"), (5, "ACGTN"), (7, "ACGTNRY")):
        _, mp = O.make_map(alphabet, alphabet[-1], True)
        M = np.array([rng.randint(-5, 5) for _ in range(alpha * alpha)], np.int32)
        go, ge = -rng.randint(0, 9), -rng.randint(0, 3)
        queries = [rand_dna(rng, rng.randint(20, 200), alphabet) for _ in range(25)]
        targets = [mutate(rng, queries[rng.randrange(25)], 3, alphabet) + "AC" for _ in range(11)]
        q = [O.encode(s, mp) for s in queries]
        t = [O.encode(s, mp) for s in targets]
        for mode in (O.GLOBAL, O.GLOCAL, O.LOCAL):
            want = _oracle_score_matrix(oracle_port, mode, q, t, alpha, M, go, ge)
            B.set_option("path", B.PATH_PACKED)
            try:
                got = B.score_matrix([s.encode() for s in queries], [s.encode() for s in targets], mode, alpha, mp, M, go, ge)
            finally:
                B.set_option("path", B.PATH_AUTO)
            assert (got == want).all(), (alpha, mode, go, ge)


def test_raw_codes_input(B, oracle_port):
    """map256=NULL: the buffers already hold symbol indices (the *_raw entry points)"""
    rng = random.Random(3)
    q = [bytes(rng.randrange(5) for _ in range(rng.randint(50, 120))) for _ in range(12)]
    t = [bytes(rng.randrange(5) for _ in range(rng.randint(50, 120))) for _ in range(7)]
    got = B.score_matrix(q, t, O.GLOBAL, 5, None, O.DNA_SCORE)
    assert (got == _oracle_score_matrix(oracle_port, O.GLOBAL, q, t, 5, O.DNA_SCORE, -7, -1)).all()
    sc, st, cn, fr = B.align_pairs(q[:7], t, O.LOCAL, 5, None, O.DNA_SCORE)
    for k in range(7):
        assert (int(sc[k]), B.fragments(k, st, cn, fr)) == oracle_port.align_raw(O.LOCAL, q[k], t[k], 5, O.DNA_SCORE, -7, -1)


def test_empty_and_degenerate_batches(B):
    bs, bi = B.best_hits([], [b"ACGT"])
    assert bs.size == 0 and bi.size == 0
    bs, bi = B.best_hits([b"ACGT", b"GG"], [], seed=-5)
    assert list(bs) == [-5, -5] and list(bi) == [-1, -1]
    # one query, one target (a single, unpaired target row) on the packed and the 32-bit path
    for path in (B.PATH_AUTO, B.PATH_WF32):
        B.set_option("path", path)
        try:
            bs, bi = B.best_hits([b"ACGTACGTTTGACCA"], [b"ACGTTCGTTTGACA"])
            gs, gi = B.best_hits([b"ACGTACGTTTGACCA"], [b"ACGTTCGTTTGACA"], O.GLOCAL, 0)
        finally:
            B.set_option("path", B.PATH_AUTO)
        assert (int(bs[0]), int(bi[0])) == (2, 0) and (int(gs[0]), int(gi[0])) == (4, 0)     # oracle: global 2, glocal 4
    sc, st, cn, fr = B.align_pairs([], [], O.LOCAL)
    assert sc.size == 0 and fr.shape[0] == 0
    sc, st, cn, fr = B.align_pairs([b"A"], [b"A"], O.GLOBAL)
    assert int(sc[0]) == 1 and B.fragments(0, st, cn, fr) == [(O.MATCH, 0, 0, 1)]
    # fragment pool too small: error code + the needed size, scores still valid
    with pytest.raises(B.AlignError) as e:
        B.align_pairs([b"ACGTACGT", b"AAAATTTT"], [b"ACGTTACGT", b"AAAACTTTT"], O.GLOBAL, frag_cap=1)
    assert e.value.code == B.ERR_FRAG_CAP


def test_job_api_resident_inputs_and_relaunch(B, oracle_port):
    rng = random.Random(12)
    queries, targets = _amplicon_set(rng, 20, 50, 200, 250)
    want = B.best_hits([s.encode() for s in queries], [s.encode() for s in targets])
    with B.BestHitsJob([s.encode() for s in queries], [s.encode() for s in targets]) as job:
        assert job.cells == sum(map(len, queries)) * sum(map(len, targets))
        for _ in range(3):                      # relaunch on resident inputs gives the same answer
            job.launch()
            bs, bi = job.fetch()
            assert (bs == want[0]).all() and (bi == want[1]).all()
        assert job.kernel_launches() >= 1


def test_full_size_properties_packed_path(B):
    """config-4-shaped data (2,000 targets) on a slice of queries: properties that need no oracle"""
    rng = np.random.default_rng(0xC4)
    n_t, n_q = 2000, 1500
    tl = rng.integers(200, 261, n_t)
    targets = [bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), int(L))) for L in tl]
    src = rng.integers(0, n_t, n_q)
    queries = []
    for i in range(n_q):
        s = bytearray(targets[src[i]])
        for _ in range(int(rng.poisson(1.5))):
            s[int(rng.integers(0, len(s)))] = b"ACGT"[int(rng.integers(0, 4))]
        queries.append(bytes(s))
    queries[:5] = [targets[j] for j in (0, 17, 999, 1998, 1999)]       # exact copies
    bs, bi = B.best_hits(queries, targets)
    # an exact copy scores its own length against itself and nothing can score higher
    assert list(bi[:5]) == [0, 17, 999, 1998, 1999]
    assert [int(x) for x in bs[:5]] == [len(targets[j]) for j in (0, 17, 999, 1998, 1999)]
    # substitutions only: the source target scores >= len - 5*subs, so the best hit does too, and it is the source
    assert (bi == src)[5:].mean() > 0.999
    # the packed and the 32-bit kernels agree on every query
    B.set_option("path", B.PATH_WF32)
    try:
        ws, wi = B.best_hits(queries[:300], targets)
    finally:
        B.set_option("path", B.PATH_AUTO)
    assert (ws == bs[:300]).all() and (wi == bi[:300]).all()
    # sharding independence: any split of the queries gives the same per-query answers
    s1, i1 = B.best_hits(queries[:700], targets)
    s2, i2 = B.best_hits(queries[700:], targets)
    assert (np.concatenate([s1, s2]) == bs).all() and (np.concatenate([i1, i2]) == bi).all()


def test_target_slicing_is_invisible(B, oracle_port):
    """work items = (query, slice of the target pairs): any slicing must give the same best hits
    (combined with a 64-bit atomicMax: higher score, then lower target index)"""
    rng = random.Random(31)
    targets = [rand_dna(rng, rng.randint(180, 260)) for _ in range(150)]
    targets[120] = targets[7]                  # duplicate far apart: lands in another slice, index 7 must still win
    targets[140] = targets[7]
    queries = [mutate(rng, targets[rng.randrange(len(targets))], rng.randint(0, 3)) for _ in range(40)]
    queries[0] = targets[7]
    ref = None
    for mode, seed in ((O.GLOBAL, -1000000), (O.GLOCAL, 0), (O.LOCAL, 0)):
        ref = None
        for split in (1, 2, 3, 5, 8, 0):
            B.set_option("split", split)
            try:
                got = B.best_hits([s.encode() for s in queries], [s.encode() for s in targets], mode, seed)
            finally:
                B.set_option("split", 0)
            if ref is None:
                ref = got
                q = [O.encode(s) for s in queries[:6]]
                t = [O.encode(s) for s in targets]
                ws, wi = oracle_port.best_hits(mode, q, t, 5, O.DNA_SCORE, -7, -1, seed)
                assert (got[0][:6] == ws).all() and (got[1][:6] == wi).all()
            assert (got[0] == ref[0]).all() and (got[1] == ref[1]).all(), (mode, split)
        if mode == O.GLOBAL:
            assert ref[1][0] == 7


def test_legacy_abi_is_thread_safe(A, oracle_port):
    """the reference calls the library from multiprocessing.dummy worker threads with the GIL released
    (ambivert.py:474-476): concurrent legacy calls must each get their own correct answer"""
    from multiprocessing.dummy import Pool
    rng = random.Random(8)
    refs = [rand_dna(rng, rng.randint(80, 240)) for _ in range(12)]
    jobs = []
    for i in range(96):
        t = refs[rng.randrange(len(refs))]
        jobs.append((mutate(rng, t, rng.randint(0, 4)), t, rng.choice([O.GLOBAL, O.LOCAL])))

    def run(job):
        q, t, mode = job
        return legacy(A, mode, q.encode(), t.encode())
    pool = Pool(8)
    got = pool.map(run, jobs)
    pool.close()
    for (q, t, mode), g in zip(jobs, got):
        assert g == oracle_port.align(mode, q.encode(), t.encode())


def test_packed_vs_wf32_random_sweep(B):
    """many small random jobs: the packed kernels (streaming global/glocal, first-generation local) must
    give the same full score matrix and best hits as the 32-bit kernel (itself checked against the oracle
    above), whatever the lengths, the alphabet, the penalties and the slicing of the targets"""
    rng = random.Random(20261020)
    for it in range(120):
        alpha = rng.choice([2, 4, 5, 5, 5, 7])
        alphabet = "ACGTNRY"[:alpha]
        _, mp = O.make_map(alphabet, alphabet[-1], True)
        if alpha == 5 and rng.random() < 0.5:
            M, go, ge = O.DNA_SCORE, -7, -1
        else:
            M = np.array([rng.randint(-6, 6) for _ in range(alpha * alpha)], np.int32)
            go, ge = -rng.randint(0, 12), -rng.randint(0, 4)
        n_q, n_t = rng.randint(1, 24), rng.randint(1, 60)
        lmax = rng.choice([8, 40, 130, 300, 520])
        targets = [rand_dna(rng, rng.randint(2, lmax), alphabet) for _ in range(n_t)]
        if rng.random() < 0.3:
            targets = [targets[0][:rng.randint(2, len(targets[0]))] if rng.random() < 0.5 else t for t in targets]
        queries = []
        for _ in range(n_q):
            if rng.random() < 0.5:
                queries.append(mutate(rng, targets[rng.randrange(n_t)], rng.randint(0, 4), alphabet))
            else:
                queries.append(rand_dna(rng, rng.randint(1, min(512, lmax + 60)), alphabet))
        mode = rng.choice([O.GLOBAL, O.GLOBAL, O.GLOCAL, O.LOCAL])
        seed = rng.choice([-1000000, 0, 5])
        q = [s.encode() for s in queries]
        t = [s.encode() for s in targets]
        B.set_option("path", B.PATH_WF32)
        try:
            want = B.score_matrix(q, t, mode, alpha, mp, M, go, ge)
            want_hits = B.best_hits(q, t, mode, seed, alpha, mp, M, go, ge)
        finally:
            B.set_option("path", B.PATH_AUTO)
        B.set_option("split", rng.choice([0, 0, 1, 2, 3, 7]))
        try:
            got = B.score_matrix(q, t, mode, alpha, mp, M, go, ge)
            got_hits = B.best_hits(q, t, mode, seed, alpha, mp, M, go, ge)
        finally:
            B.set_option("split", 0)
        bad = np.argwhere(got != want)
        assert bad.size == 0, (it, mode, alpha, go, ge, bad[:3], len(queries[bad[0][0]]), len(targets[bad[0][1]]))
        assert (got_hits[0] == want_hits[0]).all() and (got_hits[1] == want_hits[1]).all(), (it, mode, seed)


def test_streaming_kernels_many_pairs_vs_wf32(B):
    """the streaming kernels on jobs where every warp sweeps many target pairs: global climbs several staircases of
    levels (a pair enters by injection, every n-th one by an explicit reset; dp_core.cuh mm16g_levels), glocal parks the
    last rows of its targets (>= 32 rows) in shared memory and folds them once per pair; ragged lengths, every strip
    width including the odd ones, the bundled DNA scoring and random matrices / penalties"""
    rng = random.Random(31337)
    for it in range(10):
        dna = it < 6
        alpha = 5 if dna else rng.choice([4, 5, 7])
        alphabet = "ACGTNRY"[:alpha]
        _, mp = O.make_map(alphabet, alphabet[-1], True)
        if dna:
            M, go, ge = O.DNA_SCORE, -7, -1
        else:
            M = np.array([rng.randint(-5, 5) for _ in range(alpha * alpha)], np.int32)
            ge = -rng.randint(0, 3)
            go = ge - rng.randint(0, 8)
        lmax = rng.choice([40, 64, 120, 300])
        n_t = rng.choice([91, 400, 900]) if lmax <= 120 else rng.choice([91, 400])
        targets = [rand_dna(rng, rng.randint(32, lmax), alphabet) for _ in range(n_t)]
        queries = []
        for lo, hi in ((1, 64), (65, 96), (97, 128), (129, 160), (161, 192), (200, 224), (225, 256), (257, 288), (289, 320)):
            L = rng.randint(lo, hi)
            queries.append(mutate(rng, rand_dna(rng, L, alphabet), 2, alphabet) if rng.random() < 0.5 else
                           (targets[rng.randrange(n_t)] * 9)[:L])
        q = [s.encode() for s in queries]
        t = [s.encode() for s in targets]
        for mode in (O.GLOBAL, O.GLOCAL):
            B.set_option("path", B.PATH_WF32)
            try:
                want = B.score_matrix(q, t, mode, alpha, mp, M, go, ge)
            finally:
                B.set_option("path", B.PATH_AUTO)
            got = B.score_matrix(q, t, mode, alpha, mp, M, go, ge)
            bad = np.argwhere(got != want)
            assert bad.size == 0, (it, mode, alpha, go, ge, bad[:3], got[tuple(bad[0])], want[tuple(bad[0])],
                                   len(queries[bad[0][0]]), len(targets[bad[0][1]]))
            if dna:     # the 16-bit bounds hold: the streaming kernels served these queries (odd strip widths exist only there)
                with B.BestHitsJob(q, t, mode, 0, alpha, mp, M, go, ge) as job:
                    job.launch()
                    bs, bi = job.fetch()
                    ks = {r["K"] for r in job.profile()}
                assert ks & {3, 5, 7, 9}, (it, mode, ks)
                col = np.arange(n_t)[None, :]
                best = want.max(axis=1)
                first = np.where(want == best[:, None], col, n_t).min(axis=1)
                hit = best > 0
                assert (bs == np.where(hit, best, 0)).all() and (bi == np.where(hit, first, -1)).all(), (it, mode)


def test_glocal_short_targets_beside_the_streaming_kernel(B):
    """glocal: the streaming kernel takes the target pairs of >= 32 rows, the first-generation kernel the short ones at
    the end of the length-sorted list; both fold into one best key per query (first maximum in ORIGINAL target order)"""
    rng = random.Random(808)
    for n_long, n_short in ((40, 7), (41, 6), (41, 1), (1, 9), (0, 5), (33, 0)):
        targets = [rand_dna(rng, rng.randint(32, 260)) for _ in range(n_long)] + [rand_dna(rng, rng.randint(2, 31)) for _ in range(n_short)]
        targets += targets[:3]                               # duplicates: equal scores, the lower original index must win
        rng.shuffle(targets)
        queries = [rand_dna(rng, L) for L in (10, 40, 70, 100, 150, 200, 230, 280)]
        queries += [mutate(rng, (t * 8)[:rng.randint(20, 250)], 2) for t in targets[:12]]
        q = [s.encode() for s in queries]
        t = [s.encode() for s in targets]
        B.set_option("path", B.PATH_WF32)
        try:
            want = B.score_matrix(q, t, O.GLOCAL)
            want_hits = B.best_hits(q, t, O.GLOCAL, 0)
        finally:
            B.set_option("path", B.PATH_AUTO)
        got = B.score_matrix(q, t, O.GLOCAL)
        assert (got == want).all(), (n_long, n_short, np.argwhere(got != want)[:3])
        with B.BestHitsJob(q, t, O.GLOCAL, 0) as job:
            job.launch()
            bs, bi = job.fetch()
            ks = {r["K"] for r in job.profile()}
            launches = job.kernel_launches()
        assert (bs == want_hits[0]).all() and (bi == want_hits[1]).all(), (n_long, n_short)
        if n_long >= 2:
            assert ks & {3, 5, 7, 9}, ks                     # odd strip widths exist only in the streaming kernel
        if n_long >= 2 and n_short:
            assert launches >= 2 * len(ks) + 1               # per strip width: streaming launch + short-target launch; + finalize


@pytest.mark.parametrize("mode", MODES)
def test_legacy_single_calls_fuzz(A, oracle_port, mode):
    """the lean one-launch path behind the reference's own symbols: the column blocks of the one pair are
    pipelined over the warps of a CTA, so lengths around the 32-column block and 8-block round boundaries matter"""
    rng = random.Random(500 + mode)
    for it in range(60):
        la = rng.choice([1, 2, 31, 32, 33, 64, 255, 256, 257, 289, rng.randint(3, 700), rng.randint(3, 300)])
        lb = rng.choice([1, 2, 15, 16, 17, 32, 33, rng.randint(3, 700), rng.randint(3, 300)])
        a = rand_dna(rng, la, "ACGTN")
        b = mutate(rng, a, rng.randint(0, 6)) if rng.random() < 0.5 else rand_dna(rng, lb, "ACGT")
        if (mode == O.GLOCAL and len(b) < 2) or (mode == O.OVERLAP and len(a) < 2):
            continue
        go, ge = rng.choice([(-7, -1), (-7, -1), (0, -2), (-3, 0)])
        want = oracle_port.align(mode, a.encode(), b.encode(), go=go, ge=ge)
        assert legacy(A, mode, a.encode(), b.encode(), go=go, ge=ge) == want, (it, la, len(b), go, ge)


@pytest.mark.parametrize("mode", [O.LOCAL, O.GLOBAL])
def test_packed_traceback_kernel_vs_32bit_kernel_sweep(B, oracle_port, mode):
    """tb16 (two pairs per warp in 16-bit halves, DPX max with predicates) against the 32-bit wavefront kernel on
    6000 ragged pairs: identical scores and fragment lists; a sample also against the oracle.  Lengths cross every
    strip bucket (K = 2..10), the partner of a pair has a different shape, odd counts leave a pair without partner."""
    rng = random.Random(77 + mode)
    a_list, b_list = [], []
    for i in range(6001):
        la = rng.choice([1, 2, 31, 33, 64, 65, 128, 129, 150, 160, 161, 192, 224, 225, 256, 257, 320, rng.randint(1, 330),
                         62, 63, 124, 125, 155, 156, 186, 187, 217, 218, 248, 249, 310, 311])     # 31*K and 31*K + 1: the local-mode capacity
        a = rand_dna(rng, la, "ACGTN" if rng.random() < 0.1 else "ACGT")
        x = rng.random()
        if x < 0.5:
            b = mutate(rng, a, rng.randint(0, 10))
        elif x < 0.7:                         # overlap-merge shape: suffix of a + new sequence
            cut = rng.randrange(len(a))
            b = a[cut:] + rand_dna(rng, rng.randint(1, 100))
        elif x < 0.8:
            b = rand_dna(rng, rng.randint(1, 6), "AC") * rng.randint(1, 40)      # repeats: ties everywhere
            a = rand_dna(rng, rng.randint(1, 6), "AC") * rng.randint(1, 40)
        else:
            b = rand_dna(rng, rng.choice([1, 2, 3, rng.randint(4, 400), 1024, 1030]) if rng.random() < 0.9 else 1)
        a_list.append(a.encode()); b_list.append(b.encode())
    got = B.align_pairs(a_list, b_list, mode)
    B.set_option("tb16", 0)
    try:
        want = B.align_pairs(a_list, b_list, mode)
    finally:
        B.set_option("tb16", 1)
    assert (got[0] == want[0]).all() and (got[2] == want[2]).all()
    for k in range(len(a_list)):
        assert B.fragments(k, got[1], got[2], got[3]) == B.fragments(k, want[1], want[2], want[3]), (k, a_list[k], b_list[k])
    for k in range(0, len(a_list), 23):
        assert (int(got[0][k]), B.fragments(k, got[1], got[2], got[3])) == oracle_port.align(mode, a_list[k], b_list[k]), k
