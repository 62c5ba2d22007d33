"""numpy-facing wrappers of the batched C ABI (include/ambivert_align.h, part 2).

Sequences go in as lists of `bytes`/`str` (or pre-packed (buffer, offsets) pairs) and results come
back as numpy arrays.  All computation happens in the CUDA library; a failing call raises
`AlignError` with the library's message -- nothing here computes an alignment on the CPU.
"""
import ctypes as C

import numpy as np

from .align.align_ctypes import DNA_MAP, DNA_SCORE, align_c as _lib

OVERLAP, LOCAL, GLOBAL, GLOCAL = 0, 1, 2, 3
MODE_BY_NAME = {"overlap": OVERLAP, "local": LOCAL, "global": GLOBAL, "glocal": GLOCAL}
A_GAP, B_GAP, MATCH = 0, 1, 2
OK, ERR_ARGS, ERR_UNDEFINED, ERR_CUDA, ERR_NOMEM, ERR_FRAG_CAP, ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5, -6
PATH_AUTO, PATH_WF32, PATH_PACKED = 0, 1, 2

# scoring used at every pipeline call site (ambivert.py:86-89, :127-130, :170-173)
GAP_OPEN, GAP_EXTEND = -7, -1
SEED_GLOBAL = -1000000   # ambivert.py:478
SEED_LOCAL = 0           # ambivert.py:504

_vp = C.c_void_p
_lib.abv_device_count.restype = C.c_int
_lib.abv_set_device.argtypes = [C.c_int]
_lib.abv_set_thread_device.argtypes = [C.c_int]
_lib.abv_get_device.restype = C.c_int
_lib.abv_last_error.restype = C.c_char_p
_lib.abv_version.restype = C.c_char_p
_lib.abv_set_option.argtypes = [C.c_char_p, C.c_int]
_MM = [C.c_int, _vp, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int]
_lib.abv_best_hits.argtypes = _MM + [C.c_int, _vp, _vp]
_lib.abv_score_matrix.argtypes = _MM + [_vp]
_lib.abv_best_hits_multi.argtypes = _MM + [C.c_int, _vp, C.c_int, _vp, _vp]
_lib.abv_shard_bounds.argtypes = [_vp, C.c_int, C.c_int, _vp]
_lib.abv_shard_bounds.restype = None
_lib.abv_best_hits_prepare.argtypes = _MM + [C.c_int]
_lib.abv_best_hits_prepare.restype = _vp
_lib.abv_best_hits_launch.argtypes = [_vp, _vp, _vp, _vp]
_lib.abv_best_hits_fetch.argtypes = [_vp, _vp, _vp]
_lib.abv_job_cells.argtypes = [_vp]
_lib.abv_job_cells.restype = C.c_double
_lib.abv_job_kernel_launches.argtypes = [_vp]
_lib.abv_job_profile.argtypes = [_vp, _vp, C.c_int]
_lib.abv_job_profile_history.argtypes = [_vp, C.c_int, _vp, C.c_int]
_lib.abv_job_free.argtypes = [_vp]
_lib.abv_job_free.restype = None
_lib.abv_align_pairs.argtypes = [C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int,
                                 _vp, _vp, _vp, _vp, C.c_longlong, _vp]
_lib.abv_gapped_pairs.argtypes = [C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int,
                                  _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_longlong, _vp]
_lib.abv_merge_pairs.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int,
                                 _vp, _vp, _vp, _vp, C.c_longlong, _vp]
_lib.abv_cluster_pairs.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]
_lib.abv_call_variants.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.c_longlong, _vp]
_lib.abv_get_timing.argtypes = [_vp, C.c_int]
_lib.abv_pinned_alloc.argtypes = [C.c_ulonglong]
_lib.abv_pinned_alloc.restype = _vp
_lib.abv_pinned_free.argtypes = [_vp]
_lib.abv_pinned_free.restype = None
_lib.abv_int_peak.argtypes = [C.c_int, C.c_int, _vp, _vp]


class AlignError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("ambivert_b200 error %d: %s" % (code, message))
        self.code = code


def _check(rc):
    if rc != OK:
        raise AlignError(rc, (_lib.abv_last_error() or b"").decode("utf-8", "replace"))


def device_count():
    return int(_lib.abv_device_count())


def set_device(ordinal):
    _check(_lib.abv_set_device(int(ordinal)))


def set_thread_device(ordinal):
    """device of the calling thread's subsequent calls (-1: the process default); every device has its own context"""
    _check(_lib.abv_set_thread_device(int(ordinal)))


def version():
    return _lib.abv_version().decode()


def set_option(key, value):
    _check(_lib.abv_set_option(key.encode(), int(value)))


class _Pinned(object):
    """owner of one page-locked allocation; numpy arrays made from it keep it alive"""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = _lib.abv_pinned_alloc(max(self.nbytes, 16))
        if not self.ptr:
            raise AlignError(ERR_NOMEM, "cannot allocate %d bytes of page-locked memory" % self.nbytes)
        self.__array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def __del__(self, _free=_lib.abv_pinned_free):       # bound at definition: module globals may be gone at interpreter exit
        if getattr(self, "ptr", None):
            _free(self.ptr)
            self.ptr = None


def pinned_empty(count, dtype=np.uint8):
    """uninitialised numpy array in page-locked host memory (abv_pinned_alloc): copies to / from it are asynchronous and
    run at full bus speed.  Use it for the big inputs (packed sequences) and for `out=` buffers of merge_pairs/gapped_pairs."""
    dt = np.dtype(dtype)
    return np.asarray(_Pinned(int(count) * dt.itemsize)).view(dt)[:int(count)]


def pack(seqs):
    """list of bytes/str -> (uint8 buffer, int64 offsets[n+1]); (buffer, offsets) passes through"""
    if isinstance(seqs, tuple) and len(seqs) == 2 and isinstance(seqs[1], np.ndarray):
        buf, off = seqs
        return np.ascontiguousarray(buf, dtype=np.uint8), np.ascontiguousarray(off, dtype=np.int64)
    bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        np.cumsum([len(b) for b in bs], out=off[1:])
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return buf, off


def pack_text(seqs):
    """list of str -> (uint8 buffer, int64 offsets[n+1]) without one bytes object per sequence: one join, one
    encode.  Non-ASCII text raises UnicodeEncodeError, as bytes(seq, 'ascii') does at the reference's call sites
    (ambivert.py:84, :125, :168)."""
    n = len(seqs)
    off = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum(np.fromiter(map(len, seqs), dtype=np.int64, count=n), out=off[1:])
    buf = np.frombuffer("".join(seqs).encode("ascii"), dtype=np.uint8)
    return buf, off


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _scoring(alpha_len, map256, matrix):
    if map256 is None:
        mp = None
    else:
        mp = np.ascontiguousarray(np.frombuffer(bytes(bytearray(map256)), dtype=np.uint8)) if not isinstance(map256, np.ndarray) \
            else np.ascontiguousarray(map256, dtype=np.uint8)
        assert mp.size == 256
    m = np.ascontiguousarray(np.array(list(matrix), dtype=np.int32) if not isinstance(matrix, np.ndarray) else matrix, dtype=np.int32)
    assert m.size == alpha_len * alpha_len, "score matrix must be alpha_len x alpha_len"
    return mp, m


def timing():
    """timings of the last batched call: dict(h2d_ms, kernel_ms, d2h_ms, cells, launches, h2d_bytes, d2h_bytes, dp_ms)"""
    out = np.zeros(8, np.float64)
    _lib.abv_get_timing(out.ctypes.data_as(_vp), 8)
    keys = ["h2d_ms", "kernel_ms", "d2h_ms", "cells", "launches", "h2d_bytes", "d2h_bytes", "dp_ms"]
    return dict(zip(keys, out.tolist()))


def best_hits(queries, targets, mode=GLOBAL, seed=SEED_GLOBAL, alpha_len=DNA_MAP[0], map256=DNA_MAP[1],
              matrix=DNA_SCORE, gap_open=GAP_OPEN, gap_extend=GAP_EXTEND):
    """For every query the best-scoring target under the reference's selection rule
    (ambivert.py:478-485): -> (best_score int32[n_q], best_idx int32[n_q]); idx -1 = nothing beat `seed`."""
    q, qo = pack(queries)
    t, to = pack(targets)
    mp, m = _scoring(alpha_len, map256, matrix)
    n_q, n_t = len(qo) - 1, len(to) - 1
    bs = np.empty(n_q, np.int32)
    bi = np.empty(n_q, np.int32)
    _check(_lib.abv_best_hits(mode, _ptr(q), _ptr(qo), n_q, _ptr(t), _ptr(to), n_t, alpha_len,
                              _ptr(mp), _ptr(m), gap_open, gap_extend, seed, _ptr(bs), _ptr(bi)))
    return bs, bi


def shard_bounds(q_off, n_shards):
    """the contiguous query shards abv_best_hits_multi uses, balanced by summed length: int32[n_shards + 1]"""
    qo = np.ascontiguousarray(q_off, dtype=np.int64)
    out = np.zeros(int(n_shards) + 1, np.int32)
    _lib.abv_shard_bounds(_ptr(qo), len(qo) - 1, int(n_shards), _ptr(out))
    return out


def best_hits_multi(queries, targets, gpus, mode=GLOBAL, seed=SEED_GLOBAL, alpha_len=DNA_MAP[0], map256=DNA_MAP[1],
                    matrix=DNA_SCORE, gap_open=GAP_OPEN, gap_extend=GAP_EXTEND):
    """best_hits on several GPUs of the box from this one process (abv_best_hits_multi): `gpus` = a count (devices
    0..gpus-1) or a list of device ordinals (an ordinal may repeat).  Same result as best_hits."""
    q, qo = pack(queries)
    t, to = pack(targets)
    mp, m = _scoring(alpha_len, map256, matrix)
    n_q, n_t = len(qo) - 1, len(to) - 1
    dev = np.arange(int(gpus), dtype=np.int32) if isinstance(gpus, (int, np.integer)) else np.ascontiguousarray(list(gpus), dtype=np.int32)
    bs = np.empty(n_q, np.int32)
    bi = np.empty(n_q, np.int32)
    _check(_lib.abv_best_hits_multi(mode, _ptr(q), _ptr(qo), n_q, _ptr(t), _ptr(to), n_t, alpha_len,
                                    _ptr(mp), _ptr(m), gap_open, gap_extend, seed, _ptr(dev), len(dev), _ptr(bs), _ptr(bi)))
    return bs, bi


def score_matrix(queries, targets, mode=GLOBAL, alpha_len=DNA_MAP[0], map256=DNA_MAP[1], matrix=DNA_SCORE,
                 gap_open=GAP_OPEN, gap_extend=GAP_EXTEND):
    """all n_q x n_t scores (int32)"""
    q, qo = pack(queries)
    t, to = pack(targets)
    mp, m = _scoring(alpha_len, map256, matrix)
    n_q, n_t = len(qo) - 1, len(to) - 1
    out = np.empty((n_q, n_t), np.int32)
    _check(_lib.abv_score_matrix(mode, _ptr(q), _ptr(qo), n_q, _ptr(t), _ptr(to), n_t, alpha_len,
                                 _ptr(mp), _ptr(m), gap_open, gap_extend, _ptr(out)))
    return out


def _lanes(k_field):
    """abv_job_profile reports K + 0.5 for a bucket that runs the border-lane instantiation of the streaming global kernel
    (31 lanes own query columns: queries of up to 31K columns) and plain K where all 32 lanes do"""
    return 31 if k_field != int(k_field) else 32


class BestHitsJob:
    """A many-to-many job whose inputs stay resident in HBM (abv_best_hits_prepare/launch/fetch)."""

    def __init__(self, queries, targets, mode=GLOBAL, seed=SEED_GLOBAL, alpha_len=DNA_MAP[0], map256=DNA_MAP[1],
                 matrix=DNA_SCORE, gap_open=GAP_OPEN, gap_extend=GAP_EXTEND):
        q, qo = pack(queries)
        t, to = pack(targets)
        mp, m = _scoring(alpha_len, map256, matrix)
        self.n_q, self.n_t = len(qo) - 1, len(to) - 1
        self._h = _lib.abv_best_hits_prepare(mode, _ptr(q), _ptr(qo), self.n_q, _ptr(t), _ptr(to), self.n_t, alpha_len,
                                             _ptr(mp), _ptr(m), gap_open, gap_extend, seed)
        if not self._h:
            raise AlignError(ERR_ARGS, (_lib.abv_last_error() or b"").decode("utf-8", "replace"))
        self.cells = float(_lib.abv_job_cells(self._h))

    def launch(self, stream=None, d_best_score=None, d_best_idx=None):
        """enqueue the kernels; stream = raw cudaStream_t (int) or None; outputs = device pointers (int) or None"""
        _check(_lib.abv_best_hits_launch(self._h, _vp(stream) if stream else None,
                                         _vp(d_best_score) if d_best_score else None,
                                         _vp(d_best_idx) if d_best_idx else None))

    def kernel_launches(self):
        return int(_lib.abv_job_kernel_launches(self._h))

    def profile(self):
        """per packed-kernel launch of the last launch(): list of dict(K, lanes, queries, cells, ms); lanes = 31: the border-lane
        instantiation of the streaming global kernel (queries of up to 31K columns), 32: every lane owns query columns"""
        out = np.zeros((16, 4), np.float64)
        n = _lib.abv_job_profile(self._h, out.ctypes.data_as(_vp), 16)
        return [dict(K=int(r[0]), lanes=_lanes(r[0]), queries=int(r[1]), cells=float(r[2]), ms=float(r[3])) for r in out[:n]]

    def profile_history(self, last_launches):
        """the same for the last `last_launches` (<= 32) calls of launch(), timed back to back: dicts with `launch` added"""
        out = np.zeros((16 * 32, 5), np.float64)
        n = _lib.abv_job_profile_history(self._h, int(last_launches), out.ctypes.data_as(_vp), out.shape[0])
        return [dict(K=int(r[0]), lanes=_lanes(r[0]), queries=int(r[1]), cells=float(r[2]), ms=float(r[3]), launch=int(r[4])) for r in out[:n]]

    def fetch(self):
        bs = np.empty(self.n_q, np.int32)
        bi = np.empty(self.n_q, np.int32)
        _check(_lib.abv_best_hits_fetch(self._h, _ptr(bs), _ptr(bi)))
        return bs, bi

    def close(self):
        if self._h:
            _lib.abv_job_free(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def align_pairs_buffers(n, frag_cap):
    """page-locked result buffers for align_pairs(out=...): (scores int32[n], frag_start int64[n], frag_count int32[n],
    frags int32[frag_cap, 4]).  Results land in them at bus speed and without holding the calling thread; reuse them from call to call."""
    return (pinned_empty(n, np.int32), pinned_empty(n, np.int64), pinned_empty(n, np.int32), pinned_empty(4 * int(frag_cap), np.int32).reshape(-1, 4))


def align_pairs(seqs_a, seqs_b, mode, alpha_len=DNA_MAP[0], map256=DNA_MAP[1], matrix=DNA_SCORE,
                gap_open=GAP_OPEN, gap_extend=GAP_EXTEND, frag_cap=None, out=None):
    """One-to-one alignments with traceback.
    -> (scores int32[n], frag_start int64[n], frag_count int32[n], frags int32[total, 4])
    frags rows are (type, sa_start, sb_start, hsp_len); pair k owns rows frag_start[k] .. +frag_count[k].
    out: optional result buffers (align_pairs_buffers: page-locked); the returned arrays are views of them and the size of
    out[3] is the fragment pool (AlignError ERR_FRAG_CAP if it is too small)."""
    a, ao = pack(seqs_a)
    b, bo = pack(seqs_b)
    n = len(ao) - 1
    assert len(bo) - 1 == n, "align_pairs needs as many b sequences as a sequences"
    mp, m = _scoring(alpha_len, map256, matrix)
    if out is not None:
        scores, start, count, frags = out
        assert scores.dtype == np.int32 and start.dtype == np.int64 and count.dtype == np.int32 and frags.dtype == np.int32
        assert scores.size >= n and start.size >= n and count.size >= n and frags.ndim == 2 and frags.shape[1] == 4
        assert all(x.flags.c_contiguous for x in out)
        total = C.c_longlong(0)
        _check(_lib.abv_align_pairs(mode, _ptr(a), _ptr(ao), _ptr(b), _ptr(bo), n, alpha_len, _ptr(mp), _ptr(m), gap_open, gap_extend,
                                    _ptr(scores), _ptr(start), _ptr(count), _ptr(frags), frags.shape[0], C.byref(total)))
        return scores[:n], start[:n], count[:n], frags[:total.value]
    scores = np.empty(n, np.int32)
    start = np.empty(n, np.int64)
    count = np.empty(n, np.int32)
    cap = int(frag_cap) if frag_cap is not None else max(1024, 24 * n)
    for _ in range(2):
        frags = np.empty((cap, 4), np.int32)
        total = C.c_longlong(0)
        rc = _lib.abv_align_pairs(mode, _ptr(a), _ptr(ao), _ptr(b), _ptr(bo), n, alpha_len,
                                  _ptr(mp), _ptr(m), gap_open, gap_extend,
                                  _ptr(scores), _ptr(start), _ptr(count), _ptr(frags), cap, C.byref(total))
        if rc == ERR_FRAG_CAP and frag_cap is None:
            cap = int(total.value)      # retry once with the exact size the library reported
            continue
        _check(rc)
        return scores, start, count, frags[:total.value]
    _check(rc)


def gapped_pairs(seqs_a, seqs_b, mode, alpha_len=DNA_MAP[0], map256=DNA_MAP[1], matrix=DNA_SCORE,
                 gap_open=GAP_OPEN, gap_extend=GAP_EXTEND):
    """One-to-one alignments returned as gapped rows built on the device (ambivert.py:95-111, :136-152).
    -> (scores, status, start_a, start_b, out_off int64[n+1], rows_a uint8[total], rows_b uint8[total])"""
    a, ao = pack(seqs_a)
    b, bo = pack(seqs_b)
    n = len(ao) - 1
    assert len(bo) - 1 == n
    mp, m = _scoring(alpha_len, map256, matrix)
    scores, status = np.empty(n, np.int32), np.empty(n, np.int32)
    sa, sb = np.empty(n, np.int32), np.empty(n, np.int32)
    off = np.zeros(n + 1, np.int64)
    cap = int(a.size + b.size) + 16          # a gapped row is never longer than both sequences together
    ra, rb = np.empty(cap, np.uint8), np.empty(cap, np.uint8)
    total = C.c_longlong(0)
    _check(_lib.abv_gapped_pairs(mode, _ptr(a), _ptr(ao), _ptr(b), _ptr(bo), n, alpha_len, _ptr(mp), _ptr(m), gap_open, gap_extend,
                                 _ptr(scores), _ptr(status), _ptr(sa), _ptr(sb), _ptr(off), _ptr(ra), _ptr(rb), cap, C.byref(total)))
    return scores, status, sa, sb, off, ra[:total.value], rb[:total.value]


def merge_pairs(fwds, revs, minimum_overlap=10, alpha_len=DNA_MAP[0], map256=DNA_MAP[1], matrix=DNA_SCORE,
                gap_open=GAP_OPEN, gap_extend=GAP_EXTEND, out=None):
    """Local alignment + merged read of every (forward read, reverse-complemented reverse read) pair, built on
    the device (ambivert.py:320-327).  -> (scores, status, merged_off int64[n+1], merged uint8[total]);
    status: 1 merged, 0 unmergable (overlap below minimum_overlap), -1 empty alignment, -2 bad base pair.
    out: optional uint8 buffer for the merged reads (e.g. pinned_empty(...)), at least len(fwd bytes) + len(rev bytes) + 16"""
    a, ao = pack(fwds)
    b, bo = pack(revs)
    n = len(ao) - 1
    assert len(bo) - 1 == n
    mp, m = _scoring(alpha_len, map256, matrix)
    scores, status = np.empty(n, np.int32), np.empty(n, np.int32)
    off = np.zeros(n + 1, np.int64)
    cap = int(a.size + b.size) + 16          # a merged read is never longer than both reads together
    if out is not None:
        merged = out
        if merged.dtype != np.uint8 or not merged.flags.c_contiguous or merged.size < cap:
            raise ValueError("out must be a contiguous uint8 array of at least %d bytes" % cap)
        cap = int(merged.size)
    else:
        merged = np.empty(cap, np.uint8)
    total = C.c_longlong(0)
    _check(_lib.abv_merge_pairs(_ptr(a), _ptr(ao), _ptr(b), _ptr(bo), n, alpha_len, _ptr(mp), _ptr(m), gap_open, gap_extend,
                                int(minimum_overlap), _ptr(scores), _ptr(status), _ptr(off), _ptr(merged), cap, C.byref(total)))
    return scores, status, off, merged[:total.value]


def cluster_pairs(fwds, revs, trim5=None, trim3=None):
    """Cluster read pairs by md5(f[trim5:trim3] + r[trim5:trim3]) on the device (AmpliconData.add_reads,
    ambivert.py:244-262).  trim3 is the slice stop as the reference stores it (None or a negative number).
    -> (cluster_of int32[n], keys: list of hexdigests in order of first occurrence, count int32[k], first int32[k])"""
    a, ao = pack(fwds)
    b, bo = pack(revs)
    n = len(ao) - 1
    assert len(bo) - 1 == n
    cluster_of = np.empty(n, np.int32)
    digest = np.empty((max(n, 1), 16), np.uint8)
    count = np.empty(max(n, 1), np.int32)
    first = np.empty(max(n, 1), np.int32)
    k = C.c_int(0)
    _check(_lib.abv_cluster_pairs(_ptr(a), _ptr(ao), _ptr(b), _ptr(bo), n, 0 if trim5 is None else 1, 0 if trim5 is None else int(trim5),
                                  0 if trim3 is None else 1, 0 if trim3 is None else int(trim3),
                                  _ptr(cluster_of), C.byref(k), _ptr(digest), _ptr(count), _ptr(first)))
    k = k.value
    hexes = digest[:k].tobytes().hex()
    keys = [hexes[32 * i:32 * i + 32] for i in range(k)]
    return cluster_of, keys, count[:k], first[:k]


VAR_X, VAR_D, VAR_I = 0, 1, 2
VAR_ERR_LENGTH, VAR_ERR_INDEX = -1, -2


def call_variants(samples, refs, softmask=True, gap="-"):
    """Variant scan of aligned row pairs on the device (caller(), ambivert/call_variants.py:32-128).
    -> (status int32[n], rec_start int64[n], rec_count int32[n], recs int32[total, 5])
    a record is (kind, start, pos, len, run_start), see abv_call_variants in include/ambivert_align.h"""
    a, ao = pack(samples)
    b, bo = pack(refs)
    n = len(ao) - 1
    assert len(bo) - 1 == n
    status = np.zeros(n, np.int32)
    start = np.zeros(n, np.int64)
    count = np.zeros(n, np.int32)
    total = C.c_longlong(0)
    cap = max(64, n // 2)
    g = gap.encode("latin-1") if isinstance(gap, str) else bytes(gap)
    if len(g) != 1:
        raise ValueError("gap must be one character")
    for _ in range(2):
        recs = np.empty((cap, 5), np.int32)
        rc = _lib.abv_call_variants(_ptr(a), _ptr(ao), _ptr(b), _ptr(bo), n, 1 if softmask else 0, g[0],
                                    _ptr(status), _ptr(start), _ptr(count), _ptr(recs), cap, C.byref(total))
        if rc != ERR_FRAG_CAP:
            break
        cap = total.value
    _check(rc)
    return status, start, count, recs[:total.value]


def fragments(k, start, count, frags):
    """pair k's fragment list as tuples, the shape the oracle returns"""
    s = int(start[k])
    return [tuple(int(x) for x in frags[s + i]) for i in range(int(count[k]))]


def int_peak(which, iters=20000):
    """integer-pipe microbenchmark -> (lane-ops per second, nominal SM clock MHz)"""
    ops = C.c_double(0)
    mhz = C.c_double(0)
    _check(_lib.abv_int_peak(int(which), int(iters), C.byref(ops), C.byref(mhz)))
    return ops.value, mhz.value
