"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/ambivert_align.h declares, keeps the reference's struct layout, and fails loudly (no CPU
fallback) when no CUDA device is present.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "ambivert_align.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined", "push", "pop")))


def test_header_declares_reference_abi():
    names = declared_functions()
    for n in ["align", "align_raw", "local_align", "local_align_raw", "global_align", "global_align_raw",
              "glocal_align", "glocal_align_raw", "alignment_free", "alignment_new", "align_frag_free", "align_frag_count"]:
        assert n in names
    for n in ["abv_best_hits", "abv_score_matrix", "abv_align_pairs", "abv_best_hits_prepare", "abv_best_hits_launch",
              "abv_best_hits_fetch", "abv_job_free", "abv_int_peak", "abv_get_timing"]:
        assert n in names


def test_library_exports_every_declared_symbol():
    import ambivert_b200.align as A
    assert "align_c." in A.so_filename          # what the reference loader looks for (align_ctypes.py:31)
    for n in declared_functions():
        assert hasattr(A.align_c, n), "missing export: " + n


def test_struct_layout_matches_reference():
    import ambivert_b200.align as A
    assert C.sizeof(A.AlignFrag) == 24 and C.sizeof(A.Alignment) == 16     # align.h:31-41 on LP64
    assert [f[0] for f in A.AlignFrag._fields_] == ["next", "type", "sa_start", "sb_start", "hsp_len"]
    assert [f[0] for f in A.Alignment._fields_] == ["align_frag", "frag_count", "score"]
    assert (A.A_GAP, A.B_GAP, A.MATCH) == (0, 1, 2)


def test_python_names_of_reference_module():
    import ambivert_b200.align as A
    import oracle as O
    for n in ["A_GAP", "B_GAP", "MATCH", "AlignFrag", "Alignment", "align", "align_raw", "local_align", "local_align_raw",
              "global_align", "global_align_raw", "glocal_align", "glocal_align_raw", "alignment_free", "format_alignment",
              "make_map", "make_DNA_scoring_matrix", "DNA_MAP", "DNA_SCORE", "PROT_MAP", "BLOSUM45", "BLOSUM50", "BLOSUM62",
              "BLOSUM80", "BLOSUM90"]:
        assert hasattr(A, n), n
    assert A.DNA_MAP[0] == 5 and bytes(A.DNA_MAP[1]) == O.DNA_MAP.tobytes()
    assert list(A.DNA_SCORE) == [int(x) for x in O.DNA_SCORE]
    assert A.PROT_MAP[0] == 24 and len(A.BLOSUM62) == 576 and A.BLOSUM62[0] == 4 and A.BLOSUM62[575] == 1
    # make_map semantics (align_ctypes.py:140-158)
    n, m = A.make_map("ACGT", None, False)
    assert n == 4 and m[ord("a")] == 4 and m[ord("C")] == 1
    n, m = A.make_map("ACGT", 300, True)
    assert n == 255 and m[ord("x")] == 255 and m[ord("g")] == 2


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ambivert_b200.align as A
    from ambivert_b200 import batch as B
    assert B.device_count() == 0
    al = A.global_align(b"ACGT", 4, b"ACGT", 4, 5, A.DNA_MAP[1], A.DNA_SCORE, -7, -1)
    assert not al                                   # NULL, like every reference failure
    with pytest.raises(B.AlignError) as e:
        B.best_hits([b"ACGT"], [b"ACGT"])
    assert e.value.code == B.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    with pytest.raises(B.AlignError):
        B.align_pairs([b"ACGT"], [b"ACGT"], B.LOCAL)
    for call in (lambda: B.call_variants([b"ACGT"], [b"ACGA"]), lambda: B.cluster_pairs([b"ACGT"], [b"ACGA"]),
                 lambda: B.merge_pairs([b"ACGT"], [b"ACGT"])):
        with pytest.raises(B.AlignError) as e:      # the steps either side of the path have no CPU implementation either
            call()
        assert e.value.code == B.ERR_NO_DEVICE
    from ambivert_b200 import call_variants as CV
    with pytest.raises(B.AlignError):
        list(CV.caller("AAAA", "AAGA"))


def test_argument_validation_needs_no_gpu():
    from ambivert_b200 import batch as B
    with pytest.raises(B.AlignError) as e:
        B.align_pairs([b"ACGT"], [b"ACGT"], B.GLOBAL, gap_open=1)
    assert e.value.code == B.ERR_ARGS
    with pytest.raises(B.AlignError) as e:
        B.align_pairs([b"ACGT"], [b"A"], B.GLOCAL)
    assert e.value.code == B.ERR_UNDEFINED
    with pytest.raises(B.AlignError) as e:
        B.best_hits([b"A"], [b"ACGT"], mode=B.OVERLAP)
    assert e.value.code == B.ERR_UNDEFINED
    with pytest.raises(B.AlignError) as e:
        B.best_hits([b""], [b"ACGT"])
    assert e.value.code == B.ERR_ARGS
    # symbols outside the alphabet (the reference reads past its score matrix): raw codes >= alpha_len, or a byte whose
    # map entry is >= alpha_len (make_map(s, None): unknown == alpha_len) -- rejected, not silently clamped
    import numpy as np
    with pytest.raises(B.AlignError) as e:
        B.best_hits([bytes([0, 1, 5])], [bytes([0, 1, 2])], map256=None)
    assert e.value.code == B.ERR_ARGS and "outside the alphabet" in str(e.value)
    import ambivert_b200.align as A
    n, m = A.make_map("ACGT", None, True)
    with pytest.raises(B.AlignError) as e:
        B.align_pairs([b"ACGT"], [b"ACNT"], B.GLOBAL, alpha_len=n, map256=bytes(m), matrix=np.ones(16, np.int32))
    assert e.value.code == B.ERR_ARGS
    with pytest.raises(B.AlignError) as e:              # the same map is fine while no unknown byte occurs: fails later, for want of a GPU
        B.align_pairs([b"ACGT"], [b"ACGT"], B.GLOBAL, alpha_len=n, map256=bytes(m), matrix=np.ones(16, np.int32))
    import torch
    if not torch.cuda.is_available():
        assert e.value.code == B.ERR_NO_DEVICE


def test_chunk_bounds_of_the_pair_pipeline():
    """host logic of the chunk pipeline of the one-to-one jobs (align_abi.cu, pairs_chunk_bounds): the chunks tile [0, n) in
    order, there are at most 16, small jobs stay in one piece, and the first and last chunks are the small ones"""
    import random

    import ambivert_b200.align as A
    lib = A.align_c
    lib.abv_pairs_chunk_bounds.restype = C.c_int
    lib.abv_set_option.argtypes = [C.c_char_p, C.c_int]
    buf = (C.c_int * 64)()
    rng = random.Random(3)
    try:
        for per in (0, 1, 37, 1000, 200000):
            assert lib.abv_set_option(b"pairs_chunk", per) == 0
            for n in [0, 1, 2, per, max(2 * per - 1, 0), 2 * per, 2 * per + 1, 5 * per, 1000000, 40 * max(per, 1)] + [rng.randint(0, 3000000) for _ in range(50)]:
                k = lib.abv_pairs_chunk_bounds(n, buf, 64)
                b = list(buf[:k])
                assert 2 <= k <= 17 and b[0] == 0 and b[-1] == n, (per, n, b)
                assert all(x <= y for x, y in zip(b, b[1:])), (per, n, b)
                if per == 0 or n < 2 * per:
                    assert k == 2                                   # one piece
                else:
                    sizes = [y - x for x, y in zip(b, b[1:])]
                    assert k > 2 and min(sizes) > 0, (per, n, b)
                    if k - 1 < 16:
                        assert sizes[0] <= max(sizes) and sizes[-1] <= max(sizes)
                        assert max(sizes) <= max(per, 1) + max(per // 8, 1), (per, n, sizes)
    finally:
        lib.abv_set_option(b"pairs_chunk", 200000)
