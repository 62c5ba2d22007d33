"""GPU parity at medium size against results of the REAL reference C code
(tests/golden/panel_results.npz, made by oracle/make_golden_panel.py in the build container):
a config-3-shaped panel (300 targets x 4,000 amplicons: best hit per amplicon, global and glocal) and
20,000 config-2-shaped read pairs (local alignment score + every fragment)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "panel_results.npz")


@pytest.fixture(scope="module")
def G():
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/panel_results.npz not generated")
    return np.load(GOLD)


@pytest.fixture(scope="module")
def gen():
    import make_golden_panel as M
    return M


def test_panel_best_hits_match_reference_c(G, gen):
    from ambivert_b200 import batch as B
    queries, targets = gen.panel_inputs()
    assert (gen.digest(queries) + gen.digest(targets)).encode() == G["panel_sha"].tobytes(), "synthetic inputs drifted from the golden run"
    for name, mode, seed in (("global", B.GLOBAL, B.SEED_GLOBAL), ("glocal", B.GLOCAL, 0)):
        for path in (B.PATH_AUTO, B.PATH_WF32):
            if path == B.PATH_WF32 and name == "glocal":
                continue
            B.set_option("path", path)
            try:
                bs, bi = B.best_hits(queries, targets, mode, seed)
            finally:
                B.set_option("path", B.PATH_AUTO)
            assert (bs == G["panel_%s_score" % name]).all(), (name, path)
            assert (bi == G["panel_%s_idx" % name]).all(), (name, path)


def test_read_pair_local_alignments_match_reference_c(G, gen):
    from ambivert_b200 import batch as B
    fwd, rev = gen.pair_inputs()
    assert (gen.digest(fwd) + gen.digest(rev)).encode() == G["pairs_sha"].tobytes(), "synthetic inputs drifted from the golden run"
    sc, st, cn, fr = B.align_pairs(fwd, rev, B.LOCAL)
    assert (sc == G["pairs_score"]).all()
    assert (cn == G["pairs_count"]).all()
    # the library's fragment pool is unordered across pairs: gather each pair's run in pair order
    order = np.concatenate([np.arange(s, s + c) for s, c in zip(st.tolist(), cn.tolist())]) if len(st) else np.zeros(0, np.int64)
    assert (fr[order] == G["pairs_frags"]).all()
