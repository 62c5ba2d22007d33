#!/usr/bin/env python3
"""tools/dump_matrices.py -- write ambivert_b200/align/matrices.json.

The protein substitution matrices (NCBI BLOSUM45/50/62/80/90 over 'ARNDCQEGHILKMFPSTWYVBZX*') are
plain data that `ambivert.align` exposes as module constants (align_ctypes.py:173-313).  This script
reads the values out of the reference module (run in the build container, where /root/reference and
oracle/_ref exist) and stores them as JSON so that the drop-in package exposes identical constants.
"""
import json
import os
import shutil
import sys
import tempfile
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tmp = tempfile.mkdtemp(prefix="abv_ref_")
    try:
        shutil.copytree("/root/reference/ambivert", os.path.join(tmp, "ambivert"))
        shutil.copy(os.path.join(ROOT, "oracle", "_ref", "libalign_ref.so"), os.path.join(tmp, "ambivert", "align", "align_c.so"))
        sys.path.insert(0, tmp)
        warnings.simplefilter("ignore")
        from ambivert import align as A
        out = {"alphabet": "ARNDCQEGHILKMFPSTWYVBZX*"}
        for name in ("BLOSUM45", "BLOSUM50", "BLOSUM62", "BLOSUM80", "BLOSUM90"):
            out[name] = [int(x) for x in getattr(A, name)]
            assert len(out[name]) == 24 * 24
        with open(os.path.join(ROOT, "ambivert_b200", "align", "matrices.json"), "w") as f:
            json.dump(out, f, separators=(",", ":"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
