"""Parity on the HEADLINE workload itself (BASELINE.json configs[3]: 100,000 amplicons x 2,000 targets, the data
bench.py runs): the GPU scores the full job, the compiled reference C (oracle/_ref, else the oracle port) scores a
random sample of the queries against all 2,000 targets on the host cores; best score and chosen target must be
identical for every sampled query, in global and in glocal mode."""
import os

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def test_headline_workload_sample_matches_reference_c(oracle_port):
    from ambivert_b200 import batch as B, synth
    targets = synth.panel_targets(2000)
    queries, source = synth.panel_queries(targets, 100000)
    targets = [t.upper() for t in targets]
    rng = np.random.default_rng(5)
    sample = np.sort(rng.choice(len(queries), 160, replace=False))
    # make sure the special cases of the generator are in the sample: unrelated queries and IUPAC bases
    unrelated = np.flatnonzero(source < 0)[:8]
    iupac = np.array([i for i, q in enumerate(queries[:40000]) if any(c in q for c in b"NRYSWKM")][:8], dtype=np.int64)
    sample = np.unique(np.concatenate([sample, unrelated, iupac]))
    fns = {m: (O.ref().fns(m) if O.Ref.available() else None) for m in (O.GLOBAL, O.GLOCAL)}
    t_codes = [O.encode(t) for t in targets]
    q_codes = [O.encode(queries[i]) for i in sample]
    threads = os.cpu_count() or 1
    for mode, seed in ((O.GLOBAL, -1000000), (O.GLOCAL, 0)):
        bs, bi = B.best_hits(queries, targets, mode, seed)            # the full 100k x 2k job on the GPU
        _, ws, wi = oracle_port.timed_best_hits(mode, q_codes, t_codes, 5, O.DNA_SCORE, -7, -1, seed, threads, fns=fns[mode])
        assert (bs[sample] == ws).all() and (bi[sample] == wi).all(), mode
        if mode == O.GLOBAL:
            related = source >= 0
            assert (bi[related] == source[related]).mean() > 0.999      # mutated copies find the target they came from
