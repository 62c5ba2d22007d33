"""Drop-in for the reference's ambivert/align/align_ctypes.py, bound to the CUDA library.

Exposes the names the reference module exposes (align_ctypes.py:38-313): the fragment-type
constants, the AlignFrag/Alignment ctypes structures, the eight aligners + alignment_free with the
reference's argtypes/restype, format_alignment, make_map, make_DNA_scoring_matrix, DNA_MAP,
DNA_SCORE, PROT_MAP and BLOSUM45/50/62/80/90.  The shared object is located the way the reference
locates its extension (align_ctypes.py:30-35): the first file beside this module whose name contains
'align_c.'.  If it is missing the import fails -- there is nothing to fall back to.
"""
import ctypes as _C
import json as _json
import os as _os
from ctypes import POINTER, Structure, c_char_p, c_int, c_ubyte

_HERE = _os.path.dirname(_os.path.abspath(__file__))


def _load():
    override = _os.environ.get("AMBIVERT_B200_LIB")       # tuning builds of the SAME library (tools/sweep_builds.sh); no fallback
    if override:
        return _C.cdll.LoadLibrary(override), _os.path.basename(override)
    names = sorted(f for f in _os.listdir(_HERE) if 'align_c.' in f)
    if not names:
        raise OSError("ambivert_b200: CUDA alignment library align_c.*.so not found in %s (contents: %s); "
                      "build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
                      "`make -C ambivert_b200/csrc`. There is no CPU fallback." % (_HERE, _os.listdir(_HERE)))
    return _C.cdll.LoadLibrary(_os.path.join(_HERE, names[0])), names[0]


align_c, so_filename = _load()

A_GAP = 0
B_GAP = 1
MATCH = 2


class AlignFrag(Structure):
    pass


AlignFrag._fields_ = [("next", POINTER(AlignFrag)), ("type", c_int), ("sa_start", c_int),
                      ("sb_start", c_int), ("hsp_len", c_int)]


class Alignment(Structure):
    pass


Alignment._fields_ = [("align_frag", POINTER(AlignFrag)), ("frag_count", c_int), ("score", c_int)]

_RAW = [c_char_p, c_int, c_char_p, c_int, c_int, POINTER(c_int), c_int, c_int]
_CHR = [c_char_p, c_int, c_char_p, c_int, c_int, c_ubyte * 256, POINTER(c_int), c_int, c_int]


def _bind(name, argtypes):
    fn = getattr(align_c, name)
    fn.argtypes = argtypes
    fn.restype = POINTER(Alignment)
    return fn


align_raw = _bind("align_raw", _RAW)
align = _bind("align", _CHR)
local_align_raw = _bind("local_align_raw", _RAW)
local_align = _bind("local_align", _CHR)
global_align_raw = _bind("global_align_raw", _RAW)
global_align = _bind("global_align", _CHR)
glocal_align_raw = _bind("glocal_align_raw", _RAW)
glocal_align = _bind("glocal_align", _CHR)

alignment_free = align_c.alignment_free
alignment_free.argtypes = [POINTER(Alignment)]
alignment_free.restype = None


def format_alignment(alignment, s1, s2, sequence_map, scoring_matrix, match_character=None):
    """Three display rows (sequence 1, match line, sequence 2) of an Alignment.
    Same output as the reference's format_alignment (align_ctypes.py:103-137)."""
    s1, s2 = str(s1), str(s2)
    alpha, table = sequence_map
    top, mid, bot = [], [], []
    node = alignment[0].align_frag
    while node:
        f = node[0]
        n = f.hsp_len
        if f.type == MATCH:
            x, y = s1[f.sa_start:f.sa_start + n], s2[f.sb_start:f.sb_start + n]
            top.append(x)
            bot.append(y)
            for a, b in zip(x, y):
                if a == b:
                    mid.append(match_character if match_character else a)
                elif scoring_matrix[table[ord(a)] * alpha + table[ord(b)]] > 0:
                    mid.append('+')
                else:
                    mid.append(' ')
        elif f.type == A_GAP:
            top.append('-' * n)
            mid.append(' ' * n)
            bot.append(s2[f.sb_start:f.sb_start + n])
        elif f.type == B_GAP:
            top.append(s1[f.sa_start:f.sa_start + n])
            mid.append(' ' * n)
            bot.append('-' * n)
        node = f.next
    return ''.join(top), ''.join(mid), ''.join(bot)


def make_map(s, unknown, case_insensitive):
    """(alphabet size, 256-entry byte -> symbol index table); align_ctypes.py:140-158."""
    if unknown is None:
        unknown = len(s)
    elif isinstance(unknown, str):
        pos = s.find(unknown)
        unknown = pos if pos != -1 else len(s)
    else:
        unknown = min(max(0, unknown), 255)
    table = (c_ubyte * 256)(*([unknown] * 256))
    for i, ch in enumerate(s):
        table[ord(ch)] = i
        if case_insensitive:
            table[ord(ch.lower())] = i
    return max(len(s), unknown), table


def make_DNA_scoring_matrix(match, mismatch, nmatch):
    """5x5 ACGTN matrix: `match` on the ACGT diagonal, `mismatch` elsewhere among ACGT, `nmatch` for
    anything against N (align_ctypes.py:160-167)."""
    rows = [[nmatch if 4 in (a, b) else (match if a == b else mismatch) for b in range(5)] for a in range(5)]
    return (c_int * 25)(*[v for row in rows for v in row])


DNA_MAP = make_map('ACGTN', 'N', True)
DNA_SCORE = make_DNA_scoring_matrix(match=1, mismatch=-4, nmatch=0)

with open(_os.path.join(_HERE, "matrices.json")) as _f:
    _m = _json.load(_f)
PROT_MAP = make_map(_m["alphabet"], 'X', True)
BLOSUM45 = (c_int * (24 * 24))(*_m["BLOSUM45"])
BLOSUM50 = (c_int * (24 * 24))(*_m["BLOSUM50"])
BLOSUM62 = (c_int * (24 * 24))(*_m["BLOSUM62"])
BLOSUM80 = (c_int * (24 * 24))(*_m["BLOSUM80"])
BLOSUM90 = (c_int * (24 * 24))(*_m["BLOSUM90"])
del _f, _m
