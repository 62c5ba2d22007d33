"""oracle -- CPU parity checker for the CUDA alignment path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; nothing under ambivert_b200/ does.

Two back ends, same Python call shapes:
  * `port`  -- liboracle.so, the C restatement in align_oracle.c (always buildable: gcc only);
  * `ref`   -- _ref/libalign_ref.so, the UNMODIFIED reference sources lib/align/*.c compiled where
               they lie under /root/reference (present only if `make ref` ran in the build container;
               the .so travels to the GPU box, the sources do not).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libalign_ref.so")
REF_O3_SO = os.path.join(HERE, "_ref", "libalign_ref_O3.so")
REF_SRC = "/root/reference/lib/align"
REF_PKG_CPU = os.path.join(HERE, "_ref", "pkg_cpu")      # the unmodified reference package + its own C extension
REF_PKG_GPU = os.path.join(HERE, "_ref", "pkg_gpu")      # the unmodified reference package, extension slot empty

OVERLAP, LOCAL, GLOBAL, GLOCAL = 0, 1, 2, 3
MODE_NAMES = {OVERLAP: "overlap", LOCAL: "local", GLOBAL: "global", GLOCAL: "glocal"}
A_GAP, B_GAP, MATCH = 0, 1, 2
ERR_ARGS, ERR_UNDEFINED = -1, -2


def build(ref=True, quiet=True):
    """Compile liboracle.so and, when the reference sources are present, _ref/libalign_ref.so."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=out)
    if ref and os.path.isdir(REF_SRC):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=out)


# ---------------------------------------------------------------------------------------------
# scoring constants (restating ambivert/align/align_ctypes.py:140-171 for the oracle's own use)
def make_map(alphabet, unknown, case_insensitive=True):
    unk = alphabet.find(unknown) if isinstance(unknown, str) else unknown
    if unk is None or unk < 0:
        unk = len(alphabet)
    m = np.full(256, unk, dtype=np.uint8)
    for i, ch in enumerate(alphabet):
        m[ord(ch)] = i
        if case_insensitive:
            m[ord(ch.lower())] = i
    return max(len(alphabet), unk), m


def dna_matrix(match=1, mismatch=-4, nmatch=0):
    m = np.full((5, 5), mismatch, dtype=np.int32)
    np.fill_diagonal(m, match)
    m[4, :] = nmatch
    m[:, 4] = nmatch
    return m.reshape(-1).copy()


DNA_ALPHA, DNA_MAP = make_map("ACGTN", "N", True)
DNA_SCORE = dna_matrix()
GAP_OPEN, GAP_EXTEND = -7, -1


def encode(seq, map256=DNA_MAP):
    if isinstance(seq, str):
        seq = seq.encode("latin-1")
    return map256[np.frombuffer(seq, dtype=np.uint8)].tobytes()


# ---------------------------------------------------------------------------------------------
class _Frag(C.Structure):
    _fields_ = [("type", C.c_int), ("sa_start", C.c_int), ("sb_start", C.c_int), ("hsp_len", C.c_int)]


class _RefFrag(C.Structure):
    pass


_RefFrag._fields_ = [("next", C.POINTER(_RefFrag)), ("type", C.c_int), ("sa_start", C.c_int),
                     ("sb_start", C.c_int), ("hsp_len", C.c_int)]


class _RefAlignment(C.Structure):
    _fields_ = [("align_frag", C.POINTER(_RefFrag)), ("frag_count", C.c_int), ("score", C.c_int)]


_RAW_ARGS = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int]
_CHR_ARGS = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_ubyte), C.POINTER(C.c_int), C.c_int, C.c_int]


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def pack(seqs):
    """list of bytes -> (concatenated uint8 array, int64 offsets[n+1])"""
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    if seqs:
        off[1:] = np.cumsum([len(s) for s in seqs])
    buf = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, np.uint8)
    return buf, off


class Port:
    """The C restatement (align_oracle.c)."""
    kind = "port"

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        self.lib = L = C.CDLL(PORT_SO)
        L.orc_align_raw.argtypes = [C.c_int] + _RAW_ARGS + [C.POINTER(C.c_int), C.POINTER(_Frag), C.c_int]
        L.orc_align_raw.restype = C.c_int
        L.orc_align.argtypes = [C.c_int] + _CHR_ARGS + [C.POINTER(C.c_int), C.POINTER(_Frag), C.c_int]
        L.orc_align.restype = C.c_int
        L.orc_best_hits.restype = C.c_int
        L.harness_best_hits.restype = C.c_double
        L.harness_pairs.restype = C.c_double

    def align_raw(self, mode, sa, sb, alpha, matrix, go, ge):
        """-> (score, [(type, sa_start, sb_start, hsp_len), ...]) or a negative error code"""
        m = np.ascontiguousarray(matrix, dtype=np.int32)
        cap = len(sa) + len(sb) + 2
        fr = (_Frag * cap)()
        sc = C.c_int(0)
        n = self.lib.orc_align_raw(mode, sa, len(sa), sb, len(sb), alpha, _iptr(m), go, ge, C.byref(sc), fr, cap)
        if n < 0:
            return n
        return sc.value, [(f.type, f.sa_start, f.sb_start, f.hsp_len) for f in fr[:n]]

    def align(self, mode, seqa, seqb, alpha=DNA_ALPHA, map256=DNA_MAP, matrix=DNA_SCORE, go=GAP_OPEN, ge=GAP_EXTEND):
        m = np.ascontiguousarray(matrix, dtype=np.int32)
        mp = np.ascontiguousarray(map256, dtype=np.uint8)
        cap = len(seqa) + len(seqb) + 2
        fr = (_Frag * cap)()
        sc = C.c_int(0)
        n = self.lib.orc_align(mode, seqa, len(seqa), seqb, len(seqb), alpha,
                               mp.ctypes.data_as(C.POINTER(C.c_ubyte)), _iptr(m), go, ge, C.byref(sc), fr, cap)
        if n < 0:
            return n
        return sc.value, [(f.type, f.sa_start, f.sb_start, f.hsp_len) for f in fr[:n]]

    def best_hits(self, mode, queries, targets, alpha, matrix, go, ge, seed):
        """queries/targets: lists of code bytes -> (best_score[n_q], best_idx[n_q]) int32 arrays"""
        q, qo = pack(queries)
        t, to = pack(targets)
        m = np.ascontiguousarray(matrix, dtype=np.int32)
        bs = np.zeros(len(queries), np.int32)
        bi = np.zeros(len(queries), np.int32)
        rc = self.lib.orc_best_hits(C.c_int(mode), q.ctypes.data_as(C.c_void_p), qo.ctypes.data_as(C.c_void_p), C.c_int(len(queries)),
                                    t.ctypes.data_as(C.c_void_p), to.ctypes.data_as(C.c_void_p), C.c_int(len(targets)),
                                    C.c_int(alpha), _iptr(m), C.c_int(go), C.c_int(ge), C.c_int(seed), _iptr(bs), _iptr(bi))
        if rc < 0:
            raise ValueError("orc_best_hits failed: %d" % rc)
        return bs, bi

    # -- timing harness ------------------------------------------------------------------------
    def _fn(self, mode):
        name = {OVERLAP: "harness_port_overlap", LOCAL: "harness_port_local",
                GLOBAL: "harness_port_global", GLOCAL: "harness_port_glocal"}[mode]
        return C.cast(getattr(self.lib, name), C.c_void_p), C.cast(self.lib.harness_port_free, C.c_void_p)

    def timed_best_hits(self, mode, queries, targets, alpha, matrix, go, ge, seed, threads, fns=None):
        """Full-work (jump matrix + traceback per pair) many-to-many job on `threads` pthreads.
        -> (seconds, best_score, best_idx)"""
        fn, ffn = fns or self._fn(mode)
        q, qo = pack(queries)
        t, to = pack(targets)
        m = np.ascontiguousarray(matrix, dtype=np.int32)
        bs = np.zeros(len(queries), np.int32)
        bi = np.zeros(len(queries), np.int32)
        sec = self.lib.harness_best_hits(fn, ffn, q.ctypes.data_as(C.c_void_p), qo.ctypes.data_as(C.c_void_p), C.c_int(len(queries)),
                                         t.ctypes.data_as(C.c_void_p), to.ctypes.data_as(C.c_void_p), C.c_int(len(targets)),
                                         C.c_int(alpha), _iptr(m), C.c_int(go), C.c_int(ge), C.c_int(seed),
                                         _iptr(bs), _iptr(bi), C.c_int(threads))
        return sec, bs, bi

    def timed_pairs(self, mode, seqs_a, seqs_b, alpha, matrix, go, ge, threads, fns=None):
        fn, ffn = fns or self._fn(mode)
        a, ao = pack(seqs_a)
        b, bo = pack(seqs_b)
        m = np.ascontiguousarray(matrix, dtype=np.int32)
        sc = np.zeros(len(seqs_a), np.int32)
        sec = self.lib.harness_pairs(fn, ffn, a.ctypes.data_as(C.c_void_p), ao.ctypes.data_as(C.c_void_p),
                                     b.ctypes.data_as(C.c_void_p), bo.ctypes.data_as(C.c_void_p), C.c_int(len(seqs_a)),
                                     C.c_int(alpha), _iptr(m), C.c_int(go), C.c_int(ge), _iptr(sc), C.c_int(threads))
        return sec, sc


class Ref:
    """The unmodified reference C extension, compiled from /root/reference/lib/align/*.c."""
    kind = "reference"
    RAW = {OVERLAP: "align_raw", LOCAL: "local_align_raw", GLOBAL: "global_align_raw", GLOCAL: "glocal_align_raw"}
    CHR = {OVERLAP: "align", LOCAL: "local_align", GLOBAL: "global_align", GLOCAL: "glocal_align"}

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.lib = L = C.CDLL(path)
        for n in self.RAW.values():
            getattr(L, n).argtypes = _RAW_ARGS
            getattr(L, n).restype = C.POINTER(_RefAlignment)
        for n in self.CHR.values():
            getattr(L, n).argtypes = _CHR_ARGS
            getattr(L, n).restype = C.POINTER(_RefAlignment)
        L.alignment_free.argtypes = [C.POINTER(_RefAlignment)]
        L.alignment_free.restype = None

    def _unpack(self, al):
        if not al:
            return ERR_ARGS
        score = al[0].score
        out = []
        f = al[0].align_frag
        while f:
            f = f[0]
            out.append((f.type, f.sa_start, f.sb_start, f.hsp_len))
            f = f.next
        assert len(out) == al[0].frag_count
        self.lib.alignment_free(al)
        return score, out

    def align_raw(self, mode, sa, sb, alpha, matrix, go, ge):
        if (mode == GLOCAL and len(sb) < 2) or (mode == OVERLAP and len(sa) < 2):
            return ERR_UNDEFINED      # undefined behaviour upstream: never executed
        m = np.ascontiguousarray(matrix, dtype=np.int32)
        return self._unpack(getattr(self.lib, self.RAW[mode])(sa, len(sa), sb, len(sb), alpha, _iptr(m), go, ge))

    def align(self, mode, seqa, seqb, alpha=DNA_ALPHA, map256=DNA_MAP, matrix=DNA_SCORE, go=GAP_OPEN, ge=GAP_EXTEND):
        if (mode == GLOCAL and len(seqb) < 2) or (mode == OVERLAP and len(seqa) < 2):
            return ERR_UNDEFINED
        m = np.ascontiguousarray(matrix, dtype=np.int32)
        mp = np.ascontiguousarray(map256, dtype=np.uint8)
        return self._unpack(getattr(self.lib, self.CHR[mode])(seqa, len(seqa), seqb, len(seqb), alpha,
                                                               mp.ctypes.data_as(C.POINTER(C.c_ubyte)), _iptr(m), go, ge))

    def fns(self, mode):
        """(aligner, free) function addresses for Port.timed_* (the harness lives in liboracle.so)."""
        return (C.cast(getattr(self.lib, self.RAW[mode]), C.c_void_p), C.cast(self.lib.alignment_free, C.c_void_p))


def install_dropin(so_path):
    """Put a shared object into the empty extension slot of the vendored reference package (pkg_gpu) under a
    name its loader accepts (align_ctypes.py:30-32: the first file whose name contains 'align_c.').
    -> directory to put on PYTHONPATH, or None when the package was not vendored (no /root/reference at build time)"""
    import shutil
    slot = os.path.join(REF_PKG_GPU, "ambivert", "align")
    if not os.path.isdir(slot):
        return None
    for f in os.listdir(slot):
        if "align_c." in f:
            os.remove(os.path.join(slot, f))
    shutil.copy2(so_path, os.path.join(slot, "align_c." + os.path.basename(so_path).split("align_c.", 1)[-1]))
    return REF_PKG_GPU


def as_shipped_available():
    return os.path.exists(os.path.join(REF_PKG_CPU, "ambivert", "align", "align_c.ref.so"))


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        _port = Port()
    return _port


def ref():
    global _ref
    if _ref is None:
        _ref = Ref()
    return _ref


def best_available():
    """The strongest checker present: the real reference if it was compiled, else the restatement."""
    return ref() if Ref.available() else port()
