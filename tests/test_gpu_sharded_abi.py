"""pbx_sharded_*: the single-process multi-device form of the path through the C ABI, without torch.
On a 1-GPU box the device list repeats cuda:0 (several shards on one GPU); with more GPUs every device gets a shard."""
import ctypes

import numpy as np
import pytest

from oracle import oracle
from pixelbox_b200 import _native as nat
from pixelbox_b200 import synth
from pixelbox_b200.corpus import Corpus, MultiDeviceCorpus

pytestmark = pytest.mark.gpu


def devices(n):
    have = max(1, nat.lib().pbx_device_count())
    return [i % have for i in range(n)]


def same(res, want):
    o_ids, o_dist, o_dot, o_n2 = want
    return (list(res.ids) == list(o_ids) and np.array_equal(res.dist.view(np.uint32), o_dist.view(np.uint32))
            and np.array_equal(res.dot, o_dot) and np.array_equal(res.norm2, o_n2))


@pytest.mark.parametrize("n_shards", [1, 2, 3])
def test_sharded_abi_equals_oracle_with_ties_across_shards(n_shards):
    rng = np.random.default_rng(31 + n_shards)
    n, d = 50_000, 256
    cent = rng.integers(0, 256, size=(40, d))
    rows = np.clip(cent[rng.integers(0, 40, n)] + rng.integers(-2, 3, size=(n, d)), 0, 255).astype(np.uint8)
    rows[100:350] = rows[100]
    rows[n - 350:] = rows[100]                           # the same plateau in the first and in the last shard
    ids = rng.permutation(np.arange(1, n + 1)).astype(np.int64) * 3
    queries = np.concatenate([rows[[100, 20_000]], rng.integers(0, 256, size=(30, d), dtype=np.uint8)])
    with MultiDeviceCorpus(d, devices(n_shards)) as mc:
        mc.load(ids, rows)
        assert len(mc) == n
        for k, md, qs in ((100, 1e3, queries), (10, 1e3, queries[:1]), (100, 0.02, queries[:3]), (300, 1e7, queries[:2])):
            got = mc.search(qs, k, md)
            for qi, q in enumerate(qs):
                assert same(got[qi], oracle.topk(rows, ids, q, k, md, threads=4)), f"shards={n_shards} k={k} md={md} q={qi}"
        # appended rows land on the emptiest shard and are found
        extra = np.repeat(queries[2:3], 3, axis=0)
        mc.append(np.array([10**9 + 2, 10**9 + 1, 10**9 + 3], np.int64), extra)
        got = mc.search(queries[2], 5)[0]
        assert list(got.ids[:3]) == [10**9 + 1, 10**9 + 2, 10**9 + 3]


def test_sharded_abi_synthetic_equals_single_corpus():
    d, per, k = 256, 150_000, 100
    devs = devices(2)
    queries = synth.synth_queries(5, 40, d, per * len(devs), 42)
    with MultiDeviceCorpus(d, devs, capacity_hint=per * len(devs)) as mc, Corpus(d, capacity_hint=per * len(devs)) as c:
        mc.fill_synthetic(per, 42)
        c.fill_synthetic(per * len(devs), 42, 0)
        a, b = mc.search(queries, k), c.search(queries, k)
        for x, y in zip(a, b):
            assert list(x.ids) == list(y.ids) and np.array_equal(x.dist.view(np.uint32), y.dist.view(np.uint32))
        one = mc.search(queries[7], k)[0]
        assert list(one.ids) == list(b[7].ids)


def test_sharded_abi_argument_errors():
    L = nat.lib()
    h = ctypes.c_void_p(0)
    assert L.pbx_sharded_create(64, 0, None, 0, ctypes.byref(h)) == -1
    with MultiDeviceCorpus(16, devices(2)) as mc:
        assert mc.search(np.zeros((2, 16), np.uint8), 10)[0].ids.size == 0          # empty shards
        with pytest.raises(nat.PbxError):
            mc.search(np.zeros((1, 15), np.uint8), 10)
