"""Multi-GPU inside the library (abv_best_hits_multi, include/ambivert_align.h): the hashtable build sharded over the GPUs of
one box from ONE process, behind the Python API (host.match_to_reference(gpus=N)) -- the successor of the reference's
--threads (ambivert.py:474-476, :1092-1095).  The driver's GPU box has one device, so the shards are also queued on device 0
(an ordinal may repeat); with more devices visible (gpurun --gpus N) every device takes part."""
import hashlib
import threading

import numpy as np
import pytest

import oracle as O
from ambivert_b200 import batch as B, host as H, synth
from ambivert_b200.dist import shard_bounds_by_work


def test_shard_bounds_of_the_library_match_the_torchrun_sharding():
    """abv_shard_bounds (host logic, no device) cuts the queries exactly like dist.shard_bounds_by_work"""
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 7, 1000, 4321):
        lens = rng.integers(1, 300, n)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64) + 17          # offsets need not start at 0
        for shards in (1, 2, 3, 8, 16):
            got = B.shard_bounds(off, shards)
            want = shard_bounds_by_work(lens, shards)
            assert got.tolist() == want.tolist(), (n, shards)


def _panel(n_t=120, n_q=900):
    targets = synth.panel_targets(n_t)
    queries, _ = synth.panel_queries(targets, n_q)
    return queries, [t.upper() for t in targets]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [B.GLOBAL, B.GLOCAL, B.LOCAL])
def test_multi_device_call_equals_the_single_call(mode, oracle_port):
    queries, targets = _panel()
    seed = B.SEED_GLOBAL if mode != B.LOCAL else B.SEED_LOCAL
    one = B.best_hits(queries, targets, mode, seed)
    ndev = B.device_count()
    for gpus in ([0, 0, 0], list(range(ndev)) + [0], 1, [0] * 7, ndev):
        got = B.best_hits_multi(queries, targets, gpus, mode, seed)
        assert (got[0] == one[0]).all() and (got[1] == one[1]).all(), gpus
    # and the oracle on a sample drawn from all shards of the 3-way cut
    qo = B.pack(queries)[1]
    bounds = B.shard_bounds(qo, 3)
    pick = sorted(set(int(x) for r in range(3) for x in np.linspace(bounds[r], bounds[r + 1] - 1, 6)))
    want = oracle_port.best_hits(mode, [O.encode(queries[i]) for i in pick], [O.encode(t) for t in targets], 5, O.DNA_SCORE, -7, -1, seed)
    assert (want[0] == one[0][pick]).all() and (want[1] == one[1][pick]).all()


@pytest.mark.gpu
def test_more_shards_than_queries_and_errors():
    queries, targets = _panel(10, 3)
    one = B.best_hits(queries, targets)
    got = B.best_hits_multi(queries, targets, [0] * 8)
    assert (got[0] == one[0]).all() and (got[1] == one[1]).all()
    assert B.best_hits_multi([], targets, [0, 0])[0].size == 0
    with pytest.raises(B.AlignError) as e:
        B.best_hits_multi(queries, targets, [0, 63])             # no such device: the shard's error comes back
    assert e.value.code in (B.ERR_ARGS, B.ERR_NO_DEVICE, B.ERR_CUDA)
    with pytest.raises(B.AlignError):
        B.best_hits_multi(queries + [b""], targets, [0, 0])


@pytest.mark.gpu
def test_match_to_reference_gpus_builds_the_same_hash_table():
    queries, targets = _panel(60, 400)

    def table(**kw):
        amp = H.AmpliconData()
        for i, q in enumerate(queries):
            amp.merged["k%05d" % i] = q.decode()
        for j, t in enumerate(targets):
            amp.reference_sequences[("t%03d" % j, "chr1", str(j), str(j + 1), "+")] = t.decode().lower() if j % 2 else t.decode()
        amp.match_to_reference(**kw)
        return amp.reference
    base = table()
    ndev = B.device_count()
    assert len(base) > 300
    for gpus in ([0, 0], [0, 0, 0, 0, 0], ndev, list(range(ndev))):
        got = table(gpus=gpus)
        assert list(got.items()) == list(base.items())           # same entries in the same insertion order
    assert list(table(global_align=False, gpus=[0, 0]).items()) == list(table(global_align=False).items())


@pytest.mark.gpu
def test_host_threads_on_their_own_device_contexts():
    """abv_set_thread_device: concurrent host threads, each with its own device context (here all ordinals the box has,
    at least device 0 twice through the default and through an explicit choice)"""
    queries, targets = _panel(40, 200)
    want = B.best_hits(queries, targets)
    out, errs = {}, []

    def work(tag, dev):
        try:
            B.set_thread_device(dev)
            out[tag] = B.best_hits(queries, targets)
        except Exception as e:        # noqa: BLE001
            errs.append(e)
    devs = [-1, 0] + list(range(B.device_count()))
    th = [threading.Thread(target=work, args=(i, d)) for i, d in enumerate(devs)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for tag in out:
        assert (out[tag][0] == want[0]).all() and (out[tag][1] == want[1]).all()
