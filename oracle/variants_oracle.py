"""oracle/variants_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by ambivert_b200/).

CPU restatement of the reference's variant scan, caller() in ambivert/call_variants.py:32-128, for SURVEY.md
section 8f row f4.  Pinned by tests/test_variants_cpu.py against tests/golden/variants.json.gz (outputs of the
unmodified reference Python, including the known answers of its own tests, tests/test_call_variants.py:41-85).

Unlike the device code (dp_core.cuh::variants_walk, which keeps a length and a first site per pending event), this
restatement accumulates the bases themselves as the reference does, so agreement between the two also checks the
device's claim that the accumulated bases are always a contiguous slice of a row.
"""


class Pending(object):
    """the three accumulating strings of call_variants.py:47-49 and the order in which they are reported"""

    def __init__(self):
        self.bases = {"X": "", "D": "", "I": ""}

    def flush_first(self, order, site, pos, pos_shift):
        """report the first non-empty event in `order` (the reference's if/elif chains) and clear it"""
        for kind in order:
            b = self.bases[kind]
            if b:
                self.bases[kind] = ""
                where = pos if kind == "I" else pos - len(b) + pos_shift.get(kind, 0)
                return [(kind, site - len(b), where, b)]
        return []


def scan(sample, ref, gap="-", softmask=True):
    """-> list of (type, start, pos, bases); raises RuntimeError / IndexError where the reference does"""
    if len(sample) != len(ref):
        raise RuntimeError("Sequence lengths are not equal")                   # call_variants.py:54-55
    out = []
    pend = Pending()
    pos = 0
    in_leading_primer = True
    site = None
    for site in range(len(sample)):
        r, v = ref[site], sample[site]
        r_is_gap = r == gap
        r_is_masked = r == r.lower()
        if softmask and in_leading_primer:
            if r_is_masked and not r_is_gap:                                   # :58-61
                pos += 1
                continue
            if r_is_gap:                                                       # :62-69
                k = site
                while ref[k] == gap:                                           # IndexError when the run reaches the end
                    k += 1
                if ref[k] == ref[k].lower():
                    continue
            else:
                in_leading_primer = False                                      # :73-75
        elif softmask and not r_is_gap and r_is_masked:                        # :70-72
            break
        else:
            in_leading_primer = False
        if v.upper() == r.upper() and not r_is_gap and v != gap:               # :77-88
            pos += 1
            out += pend.flush_first("XDI", site, pos, {})
        elif r_is_gap:                                                         # :89-96
            pend.bases["I"] += v
            out += pend.flush_first("XD", site, pos, {"D": 1})
        elif v == gap:                                                         # :97-105
            pend.bases["D"] += r
            pos += 1
            out += pend.flush_first("XI", site, pos, {})
        else:                                                                  # :106-117
            pos += 1
            out += pend.flush_first("X", site, pos, {})
            pend.bases["X"] += v
            out += pend.flush_first("ID", site, pos, {})
    pos += 1                                                                   # :119-128
    for kind in "XDI":
        b = pend.bases[kind]
        if b:
            out.append((kind, site - len(b) + 1, pos if kind == "I" else pos - len(b), b))
    return out
