"""world_size-2/3 tests of the sharded path's host logic on CPU (gloo): row partition, the
all-gather of pbx_hit records and the merge under (dist, image_id).  Each rank's local shard is
played by the oracle (test infrastructure standing in for the GPU shard, which cannot run here);
the exchange and the merge are the product's."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle
from pixelbox_b200 import _native as nat
from pixelbox_b200 import build as pbx_build
from pixelbox_b200.shard import ShardedCorpus, shard_rows


class OracleShard:
    """Same surface as Corpus for what ShardedCorpus needs on the host path."""

    def __init__(self, dim):
        self.dim = dim
        self.ids = np.zeros(0, np.int64)
        self.rows = np.zeros((0, dim), np.uint8)

    def load(self, ids, rows):
        self.ids, self.rows = np.asarray(ids, np.int64).copy(), np.asarray(rows, np.uint8).copy()

    def __len__(self):
        return len(self.ids)

    def search_hits(self, queries, k, max_dist):
        nq = len(queries)
        hits = np.zeros((nq, k), nat.HIT_DTYPE)
        hits["image_id"] = np.iinfo(np.int64).max
        hits["dist"] = np.inf
        cnt = np.zeros(nq, np.uint32)
        for qi, q in enumerate(queries):
            if len(self.ids) == 0:
                continue
            o_ids, o_dist, o_dot, o_n2 = oracle.topk(self.rows, self.ids, q, k, max_dist)
            c = len(o_ids)
            cnt[qi] = c
            hits["image_id"][qi, :c], hits["dist"][qi, :c], hits["dot"][qi, :c], hits["norm2"][qi, :c] = o_ids, o_dist, o_dot, o_n2
        return hits, cnt


def _table(seed, n, d):
    rng = np.random.default_rng(seed)
    cent = rng.integers(0, 256, size=(7, d))
    rows = np.clip(cent[rng.integers(0, 7, n)] + rng.integers(-1, 2, size=(n, d)), 0, 255).astype(np.uint8)
    if n > 90:
        rows[40:90] = rows[40]                   # ties that straddle shard boundaries
    ids = np.arange(1, n + 1, dtype=np.int64) * 7
    queries = np.stack([rows[min(40, n - 1)], rows[n - 1], rng.integers(0, 256, d, dtype=np.uint8)])
    return ids, rows, queries


def _worker(rank, world, port, n, d, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids, rows, queries = _table(5, n, d)
        sc = ShardedCorpus(d, local=OracleShard(d))
        sc.load_table(ids, rows)
        first, count = shard_rows(n, rank, world)
        assert len(sc.local) == count and (count == 0 or sc.local.ids[0] == ids[first])
        assert sc.total_rows() == n
        ok = True
        for k, md in ((10, 1e3), (64, 1e3), (64, 0.03), (100, 1e7)):
            res = sc.search(queries, k, md)
            for qi, q in enumerate(queries):
                o_ids, o_dist, o_dot, o_n2 = oracle.topk(rows, ids, q, k, md)
                ok &= list(res[qi].ids) == list(o_ids)
                ok &= np.array_equal(res[qi].dist.view(np.uint32), o_dist.view(np.uint32))
                ok &= np.array_equal(res[qi].dot, o_dot) and np.array_equal(res[qi].norm2, o_n2)
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
            f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,n", [(2, 2001), (3, 500), (2, 1)])
def test_sharded_search_matches_global_oracle(tmp_path, world, n):
    pbx_build.build()
    oracle.build()
    mp.spawn(_worker, args=(world, _free_port(), n, 24, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"rank{r}.txt").read_text() == "ok"


def test_shard_rows_partition_is_exact():
    for n in (0, 1, 7, 100, 1_000_000_007):
        for w in (1, 2, 3, 8):
            parts = [shard_rows(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
