"""Developer repro: one small batched search (tensor-core path)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelbox_b200.corpus import Corpus
from oracle import oracle
d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 32
rng = np.random.default_rng(1)
corpus = rng.integers(0, 256, size=(n, d), dtype=np.uint8)
ids = np.arange(1, n + 1, dtype=np.int64)
q = rng.integers(0, 256, size=(nq, d), dtype=np.uint8)
with Corpus(d) as c:
    c.load(ids, corpus)
    res = c.search(q, 100, 1e3)
    ok = 0
    for qi in range(min(nq, 8)):
        o = oracle.topk(corpus, ids, q[qi], 100, 1e3, threads=4)
        ok += int(list(res[qi].ids) == list(o[0]))
    print("batched_queries", c.stats().batched_queries, "match", ok, "of", min(nq, 8))
