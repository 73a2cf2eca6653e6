"""Reduced-size driver for compute-sanitizer (tools/sanitize.sh): every kernel family of the path once, small enough to
finish under instrumentation.  Cases: single queries (certified), forced exact passes (device-side tail launches),
back-to-back async queries on one stream, a batched call with ties (tensor-core kernel + batched finalize + exact
passes), append concurrent with search, pair distances / quantizer, a 2-shard pbx_sharded corpus on this GPU."""
import sys
import threading

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from pixelbox_b200.corpus import Corpus, MultiDeviceCorpus, cosine_distance_pairs, quantize  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "all"
rng = np.random.default_rng(3)
n, d = 24_000, 256
cent = rng.integers(0, 256, size=(30, d))
rows = np.clip(cent[rng.integers(0, 30, n)] + rng.integers(-2, 3, size=(n, d)), 0, 255).astype(np.uint8)
rows[50:200] = rows[50]
ids = np.arange(1, n + 1, dtype=np.int64)
queries = np.concatenate([rows[[50, 9000]], rng.integers(0, 256, size=(38, d), dtype=np.uint8)])


def check(got, q, k, md=1e3, r=rows, i=ids):
    o = oracle.topk(r, i, q, k, md, threads=4)
    assert list(got.ids) == list(o[0]) and np.array_equal(got.dist.view(np.uint32), o[1].view(np.uint32))


if case in ("all", "single"):
    with Corpus(d) as c:
        c.load(ids, rows)
        c.set_batch_min(0xFFFFFFFF)
        for qi in range(3):
            check(c.search(queries[qi], 100)[0], queries[qi], 100)
        c.set_candidate_slack(1)                                   # forces the exact pass (tail launch from the finalize kernel)
        for qi in range(3):
            check(c.search(queries[qi], 50)[0], queries[qi], 50)
        assert c.stats().exact_passes > 0
        got = c.search(queries[:6], 20)                            # six queries back to back on one stream
        for qi in range(6):
            check(got[qi], queries[qi], 20)
    print("case single ok")
if case in ("all", "batched"):
    with Corpus(d) as c:
        c.load(ids, rows)
        got = c.search(queries, 100)
        assert c.stats().batched_queries == len(queries)
        for qi in range(0, len(queries), 5):
            check(got[qi], queries[qi], 100)
        assert c.stats().exact_passes > 0                          # the plateau of identical rows
    print("case batched ok")
if case in ("all", "append"):
    with Corpus(d, capacity_hint=16) as c:
        stop = threading.Event()

        def writer():
            for b in range(0, n, 1500):
                c.append(ids[b:b + 1500], rows[b:b + 1500])
            stop.set()

        t = threading.Thread(target=writer)
        t.start()
        while not stop.is_set():
            r = c.search(queries[1], 10)[0]
            m = len(c)
            assert len(r.ids) <= 10 and (len(r.ids) == 0 or r.ids.max() <= n)
        t.join()
        check(c.search(queries[1], 10)[0], queries[1], 10)
    print("case append ok")
if case in ("all", "misc"):
    a, b = rows[:300], rows[300:600]
    dist, dot, na, nb = cosine_distance_pairs(a, b)
    for j in (0, 17, 299):
        assert np.float32(oracle.cosine_distance(a[j], b[j])).view(np.uint32) == dist[j].view(np.uint32)
    assert list(quantize(np.array([-1.0, 1.0, 0.0, 0.1], np.float32))) == [0x00, 0xFF, 0x80, 0x8C]
    print("case misc ok")
if case in ("all", "sharded"):
    with MultiDeviceCorpus(d, [0, 0]) as mc:
        mc.load(ids, rows)
        got = mc.search(queries[:4], 100)
        for qi in range(4):
            check(got[qi], queries[qi], 100)
        got = mc.search(queries, 20)
        for qi in range(0, len(queries), 7):
            check(got[qi], queries[qi], 20)
    print("case sharded ok")
print("sanitize_case done:", case)
