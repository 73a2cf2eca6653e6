import numpy as np, sys
sys.path.insert(0, "/root/repo")
from pixelbox_b200.corpus import Corpus
from oracle import oracle
d = int(sys.argv[1]); n = int(sys.argv[2]); nq = int(sys.argv[3]); k = int(sys.argv[4])
rng = np.random.default_rng(1000 + d)
corpus = rng.integers(0, 256, size=(n, d), dtype=np.uint8)
ids = np.arange(1, n + 1, dtype=np.int64)
q = rng.integers(0, 256, size=(nq, d), dtype=np.uint8)
with Corpus(d) as c:
    c.load(ids, corpus)
    got = c.search(q, k)
    print("batched", c.stats().batched_queries)
    for i in range(nq):
        o = oracle.topk(corpus, ids, q[i], k, 1e3)
        assert list(got[i].ids) == list(o[0]), i
print("ok", d, n, nq, k)
