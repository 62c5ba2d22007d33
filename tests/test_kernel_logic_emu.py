"""CPU tests of the DEVICE cell logic (ambivert_b200/csrc/dp_core.cuh).

tests/emu/emu.cpp includes the very header the sm_100a kernels are built from and drives its
__host__ __device__ functions lane by lane, mirroring the kernels' warp loops.  This pins the
recurrences, tie rules, direction nibbles, traceback walk and the packed 16-bit arithmetic against the
oracle without a GPU.  It is a test of the kernel source, not a product path.
"""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

import oracle as O
from util import as_result, case_scoring, golden

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
MODES = [O.OVERLAP, O.LOCAL, O.GLOBAL, O.GLOCAL]
KS = [2, 4, 6, 8, 10, 12, 16]


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU_DIR, "libemu.so")
    src = os.path.join(EMU_DIR, "emu.cpp")
    hdr = os.path.join(os.path.dirname(HERE), "ambivert_b200", "csrc", "dp_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-o", so, src])
    lib = C.CDLL(so)
    lib.emu_wf32.restype = C.c_int
    lib.emu_mm16.restype = C.c_int
    return lib


def emu_align(lib, mode, sa, sb, alpha, M, go, ge):
    M = np.ascontiguousarray(M, dtype=np.int32)
    cap = len(sa) + len(sb) + 2
    fr = np.zeros(cap * 4, np.int32)
    sc = C.c_int(0)
    n = lib.emu_wf32(C.c_int(mode), sa, C.c_int(len(sa)), sb, C.c_int(len(sb)), C.c_int(alpha),
                     M.ctypes.data_as(C.c_void_p), C.c_int(go), C.c_int(ge), C.byref(sc),
                     fr.ctypes.data_as(C.c_void_p), C.c_int(cap))
    assert 0 <= n <= cap
    return sc.value, [tuple(int(x) for x in fr[4 * i:4 * i + 4]) for i in range(n)]


def emu_pair_scores(lib, mode, K, q, t1, t2, alpha, M, go, ge):
    M = np.ascontiguousarray(M, dtype=np.int32)
    s1, s2 = C.c_int(0), C.c_int(0)
    rc = lib.emu_mm16(C.c_int(mode), C.c_int(K), q, C.c_int(len(q)), t1, C.c_int(len(t1)), t2, C.c_int(len(t2)),
                      C.c_int(alpha), M.ctypes.data_as(C.c_void_p), C.c_int(go), C.c_int(ge), C.byref(s1), C.byref(s2))
    assert rc == 0
    return s1.value, s2.value


def undefined(mode, sa, sb):
    return (mode == O.GLOCAL and len(sb) < 2) or (mode == O.OVERLAP and len(sa) < 2)


def test_wf32_logic_golden_vectors(emu, oracle_port):
    g = golden("fuzz_ref_vectors.json")
    n = 0
    for c in g["cases"]:
        alpha, mp, M = case_scoring(c)
        sa, sb = O.encode(c["a"], mp), O.encode(c["b"], mp)
        for m in MODES:
            want = as_result(c["modes"][O.MODE_NAMES[m]])
            if isinstance(want, int):
                continue
            assert emu_align(emu, m, sa, sb, alpha, M, c["go"], c["ge"]) == want, (c["kind"], O.MODE_NAMES[m], c["a"], c["b"])
            n += 1
    assert n > 1500
    for c in golden("kat_reference_tests.json"):
        for m in MODES:
            want = as_result(c["modes"][O.MODE_NAMES[m]])
            if isinstance(want, int):
                continue
            got = emu_align(emu, m, O.encode(c["s1"]), O.encode(c["s2"].upper()), 5, O.DNA_SCORE, -7, -1)
            assert got == want, (c["name"], O.MODE_NAMES[m])


def test_wf32_logic_fuzz_vs_oracle(emu, oracle_port):
    rng = random.Random(4242)
    chk = oracle_port
    for it in range(150):
        m = rng.choice(MODES)
        alpha = rng.choice([2, 4, 5, 5, 5, 7, 24])
        la = rng.choice([1, 2, 3, 31, 32, 33, 64, 65, rng.randint(1, 100), rng.randint(1, 100)])
        sa = bytes(rng.randrange(alpha) for _ in range(la))
        if rng.random() < 0.55:
            sb = bytearray(sa)
            for _ in range(rng.randint(0, 6)):
                if len(sb) > 2 and rng.random() < 0.5:
                    p = rng.randrange(len(sb)); del sb[p:p + rng.randint(1, 6)]
                else:
                    p = rng.randrange(len(sb) + 1); sb[p:p] = bytes(rng.randrange(alpha) for _ in range(rng.randint(1, 6)))
            sb = bytes(sb) or b"\x00"
        else:
            sb = bytes(rng.randrange(alpha) for _ in range(rng.choice([1, 2, 3, rng.randint(1, 100)])))
        if undefined(m, sa, sb):
            continue
        if alpha == 5 and rng.random() < 0.6:
            M = O.DNA_SCORE
        else:
            M = np.array([rng.randint(-5, 5) for _ in range(alpha * alpha)], np.int32)
        go, ge = (-7, -1) if rng.random() < 0.5 else (-rng.randint(0, 9), -rng.randint(0, 3))
        assert emu_align(emu, m, sa, sb, alpha, M, go, ge) == chk.align_raw(m, sa, sb, alpha, M, go, ge), (it, m, sa, sb, go, ge)


def test_packed_logic_vs_oracle(emu, oracle_port):
    """the packed 16x2 rows (global, local, glocal, overlap): two targets per word, every K, ragged lengths"""
    rng = random.Random(99)
    chk = oracle_port
    for it in range(900):
        m = rng.choice([O.GLOBAL, O.LOCAL, O.GLOCAL, O.OVERLAP])
        K = rng.choice(KS)
        lq = rng.randint(max(1, 32 * K - 70), 32 * K) if rng.random() < 0.7 else rng.randint(1, 32 * K)
        if m == O.OVERLAP:
            lq = max(lq, 2)                                  # a first sequence of one symbol is undefined in the reference (align.c:183)
        alpha = 5
        q = bytes(rng.choice([0, 1, 2, 3, 0, 1, 2, 3, 4]) for _ in range(lq))

        def target():
            if rng.random() < 0.6:
                t = bytearray(q)
                for _ in range(rng.randint(0, 5)):
                    if len(t) > 4 and rng.random() < 0.5:
                        p = rng.randrange(len(t)); del t[p:p + rng.randint(1, 5)]
                    else:
                        p = rng.randrange(len(t) + 1); t[p:p] = bytes(rng.randrange(4) for _ in range(rng.randint(1, 5)))
                if rng.random() < 0.3:
                    t = t[rng.randint(0, len(t) // 3):]
                t = bytes(t)
            else:
                t = bytes(rng.randrange(5) for _ in range(rng.randint(2, 300)))
            return t if len(t) >= 2 else b"\x00\x01"
        t1 = target()
        t2 = target() if rng.random() < 0.85 else b""
        if rng.random() < 0.7:
            M, go, ge = O.DNA_SCORE, -7, -1
        else:
            M = np.array([rng.randint(-4, 4) for _ in range(25)], np.int32)
            go, ge = -rng.randint(0, 8), -rng.randint(0, 3)
            if m == O.OVERLAP and go > ge:
                go = ge                                      # the packed overlap form is exact for go <= ge (dp_core.cuh, mm16_overlap_cands)
        s1, s2 = emu_pair_scores(emu, m, K, q, t1, t2, alpha, M, go, ge)
        assert s1 == chk.align_raw(m, q, t1, alpha, M, go, ge)[0], (it, m, K, lq, len(t1), len(t2))
        if t2:
            assert s2 == chk.align_raw(m, q, t2, alpha, M, go, ge)[0], (it, m, K, lq, len(t1), len(t2))


@pytest.mark.parametrize("mode", [O.GLOBAL, O.GLOCAL])
def test_streaming_logic_vs_oracle(emu, oracle_port, mode):
    """mm16g_kernel's pipeline: pairs streaming back to back through the lanes, drift-normalised biased
    operands, pending corner captures, tiny targets (< 32 rows), single (unpaired) last target; for
    glocal the dependency-free first row and the last-row maxima collected lane by lane"""
    rng = random.Random(2024 + mode)
    chk = oracle_port
    lmin = 32 if mode == O.GLOCAL else 1      # glocal streams only targets of >= 32 rows (deferred last rows), shorter ones take mm16_kernel
    for it in range(150):
        # global: pairs climb a staircase of levels and every n-th is reset explicitly (dp_core.cuh, mm16g_levels);
        # short staircases make a handful of pairs wrap around, 1 = every pair reset explicitly
        emu.emu_set_level_cap(C.c_int(rng.choice([1, 2, 3, 5, 1 << 20])))
        K = rng.choice(KS + [7, 7, 3, 5, 9])
        # global: the border-lane instantiation (lane 31 emits column -1, up to 31K columns: dp_core.cuh, mm16g_border) or the one with all 32 lanes
        bcap = emu.emu_mm16g_border_capacity(C.c_int(K), C.c_int(1 if mode == O.GLOCAL else 0))
        border = 1 if (bcap and rng.random() < 0.6) else 0
        cap = bcap if border else 32 * K
        lq = rng.randint(max(1, cap - 70), cap) if rng.random() < 0.7 else rng.randint(1, cap)
        q = bytes(rng.choice([0, 1, 2, 3, 0, 1, 2, 3, 4]) for _ in range(lq))
        n_pairs = rng.choice([1, 2, 3, 4, 7, 12])
        lens = sorted([max(lmin, rng.choice([1, 2, 5, 31, 32, 33, rng.randint(lmin, 300), rng.randint(150, 300)])) for _ in range(2 * n_pairs)], reverse=True)
        if mode == O.GLOBAL and rng.random() < 0.15:
            lens = sorted([rng.randint(1, 31) for _ in lens], reverse=True)   # only short targets: 32-step pairs, separator row at index 32
        if rng.random() < 0.3:
            lens[1] = lens[0]                                # equal lengths: both halves end on the same row
        if rng.random() < 0.5:
            lens[-1] = 0                                     # odd number of targets: the last pair is single
        targets = []
        for L in lens:
            if L and rng.random() < 0.5 and L <= lq:
                st = rng.randint(0, lq - L)
                t = bytearray(q[st:st + L])
                for _ in range(rng.randint(0, 3)):
                    t[rng.randrange(L)] = rng.randrange(5)
                targets.append(bytes(t))
            else:
                targets.append(bytes(rng.randrange(5) for _ in range(L)))
        if rng.random() < 0.7:
            M, go, ge = O.DNA_SCORE, -7, -1
        else:
            M = np.array([rng.randint(-4, 4) for _ in range(25)], np.int32)
            go, ge = -rng.randint(0, 8), -rng.randint(0, 3)
            if go > ge and border:
                go = ge              # an opening cheaper than an extension stays off the border-lane instantiation (align_abi.cu, job_prepare)
        M = np.ascontiguousarray(M, dtype=np.int32)
        t1 = [targets[2 * i] for i in range(n_pairs)]
        t2 = [targets[2 * i + 1] for i in range(n_pairs)]
        arr = lambda xs: (C.c_char_p * len(xs))(*xs)
        l1 = np.array([len(x) for x in t1], np.int32)
        l2 = np.array([len(x) for x in t2], np.int32)
        s1 = np.full(n_pairs, -99999, np.int32)
        s2 = np.full(n_pairs, -99999, np.int32)
        rc = emu.emu_mm16g(C.c_int(mode), C.c_int(K), C.c_int(border), q, C.c_int(lq), C.c_int(n_pairs), arr(t1), l1.ctypes.data_as(C.c_void_p), arr(t2),
                           l2.ctypes.data_as(C.c_void_p), C.c_int(5), M.ctypes.data_as(C.c_void_p), C.c_int(go), C.c_int(ge),
                           s1.ctypes.data_as(C.c_void_p), s2.ctypes.data_as(C.c_void_p))
        assert rc == 0, (it, rc)
        for i in range(n_pairs):
            assert s1[i] == chk.align_raw(mode, q, t1[i], 5, M, go, ge)[0], (it, K, lq, i, list(l1), list(l2), go, ge)
            if len(t2[i]):
                assert s2[i] == chk.align_raw(mode, q, t2[i], 5, M, go, ge)[0], (it, K, lq, i, list(l1), list(l2), go, ge)


TB16_KS = [2, 4, 5, 6, 7, 8, 10]


def emu_tb16(lib, mode, K, pair0, pair1, alpha, M, go, ge):
    """-> [(score, frags) of pair 0, (score, frags) of pair 1]; frags in sequence order"""
    M = np.ascontiguousarray(M, dtype=np.int32)
    pad_score = -(int(np.abs(M).max()) + 1)
    cap = max(len(pair0[0]) + len(pair0[1]), len(pair1[0]) + len(pair1[1])) + 2
    f0, f1 = np.zeros(cap * 4, np.int32), np.zeros(cap * 4, np.int32)
    sc, nf = (C.c_int * 2)(), (C.c_int * 2)()
    lib.emu_tb16.restype = C.c_int
    rc = lib.emu_tb16(C.c_int(mode), C.c_int(K), pair0[0], C.c_int(len(pair0[0])), pair0[1], C.c_int(len(pair0[1])),
                      pair1[0], C.c_int(len(pair1[0])), pair1[1], C.c_int(len(pair1[1])), C.c_int(alpha), M.ctypes.data_as(C.c_void_p),
                      C.c_int(go), C.c_int(ge), C.c_int(pad_score), sc, nf, f0.ctypes.data_as(C.c_void_p), f1.ctypes.data_as(C.c_void_p), C.c_int(cap))
    assert rc == 0
    out = []
    for h, f in enumerate((f0, f1)):
        n = nf[h]
        assert 0 <= n <= cap
        out.append((sc[h], [tuple(int(x) for x in f[4 * i:4 * i + 4]) for i in range(n)][::-1]))
    return out


def test_tb16_logic_vs_oracle(emu, oracle_port):
    """the packed traceback kernel's row function, plane layout and end-cell rules on the CPU: two pairs of different
    shapes per word, every strip width, local and global, ties (repeats), N, other penalties and matrices"""
    rng = random.Random(1616)
    chk = oracle_port

    def pair(cap, alpha, alphabet_n):
        la = rng.randint(max(1, cap - 40), cap) if rng.random() < 0.5 else rng.randint(1, cap)
        if rng.random() < 0.15:
            unit = bytes(rng.randrange(2) for _ in range(rng.randint(1, 4)))
            sa = (unit * (la // len(unit) + 1))[:la]
        else:
            sa = bytes(rng.randrange(alphabet_n) for _ in range(la))
        x = rng.random()
        if x < 0.5:
            sb = bytearray(sa)
            for _ in range(rng.randint(0, 6)):
                if len(sb) > 2 and rng.random() < 0.5:
                    p = rng.randrange(len(sb)); del sb[p:p + rng.randint(1, 6)]
                else:
                    p = rng.randrange(len(sb) + 1); sb[p:p] = bytes(rng.randrange(alphabet_n) for _ in range(rng.randint(1, 6)))
            sb = bytes(sb) or b"\x00"
        elif x < 0.7:
            cut = rng.randrange(len(sa))
            sb = sa[cut:] + bytes(rng.randrange(alphabet_n) for _ in range(rng.randint(1, 60)))
        else:
            sb = bytes(rng.randrange(alphabet_n) for _ in range(rng.choice([1, 2, 3, rng.randint(1, 200)])))
        return sa, sb

    for it in range(400):
        mode = rng.choice([O.LOCAL, O.GLOBAL])
        K = rng.choice(TB16_KS)
        if rng.random() < 0.6:
            alpha, M, go, ge = 5, O.DNA_SCORE, -7, -1
        else:
            alpha = rng.choice([2, 4, 5, 7])
            M = np.array([rng.randint(-5, 5) for _ in range(alpha * alpha)], np.int32)
            go, ge = -rng.randint(1, 9), -rng.randint(0, 3)
        cap = (31 if mode == O.LOCAL else 32) * K          # local keeps lane 31 free of columns (dp_core.cuh::tb16_max_la)
        p0, p1 = pair(cap, alpha, alpha), pair(cap, alpha, alpha)
        got = emu_tb16(emu, mode, K, p0, p1, alpha, M, go, ge)
        for h, (sa, sb) in enumerate((p0, p1)):
            assert got[h] == chk.align_raw(mode, sa, sb, alpha, M, go, ge), (it, mode, K, h, sa, sb, go, ge)


def test_level_plan_invariants(emu):
    """mm16g_levels: every level of the staircase keeps a pair's values inside the biased 16-bit range, and a step is
    large enough for the injected constant to dominate everything the previous pair left behind"""
    rng = random.Random(5)
    out = (C.c_int * 3)()
    seen_many = 0
    for _ in range(4000):
        K = rng.choice([2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16])
        P = 32 * K
        rows = rng.choice([32, 48, 96, 272, 528, 1040])
        ge = -rng.randint(0, 4)
        go = -rng.randint(0, 12)
        maxpos = rng.randint(0, 6)
        maxabs = max(1, maxpos, rng.randint(0, 8), -go, -ge)
        emu.emu_levels(C.c_int(P), C.c_int(rows), C.c_int(go), C.c_int(ge), C.c_int(maxabs), C.c_int(maxpos), out)
        n, lvl0, step = out[0], out[1], out[2]
        W = (P + rows + 12) * (maxabs + abs(ge))
        assert n >= 1
        if n == 1:
            assert lvl0 == 0 and step == 0          # every pair reset explicitly, values centred as before
            continue
        seen_many += 1
        assert W <= 16000 and go <= ge
        U = (min(P, rows + 1) + 3) * (maxpos + 2 * abs(ge))     # d diagonal steps add at most d * (max S + 2|g|) in drift form
        top = lvl0 + (n - 1) * step
        assert lvl0 - W >= -16000 and top + U <= 16000
        assert step >= U + abs(go - ge) + abs(go + ge)
    assert seen_many > 1000
