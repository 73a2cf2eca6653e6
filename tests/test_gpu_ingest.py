"""The ingest side of the path: per-image appends as the writer loop makes them
(src/engine.rs:186-203, :251-256), coalesced inside the library; rows and embeddings that are already on the device;
growth of the shard by mapping memory (no copy) while searches run."""
import os
import subprocess
import sys
import threading
import time

import numpy as np
import pytest

from oracle import oracle
from pixelbox_b200 import _native as nat
from pixelbox_b200.corpus import Corpus, quantize, quantize_device

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def same(res, want):
    return list(res.ids) == list(want[0]) and np.array_equal(res.dist.view(np.uint32), want[1].view(np.uint32))


def test_single_row_appends_are_visible_at_once_and_equal_a_bulk_load():
    rng = np.random.default_rng(2)
    n, d = 5000, 64
    rows = rng.integers(0, 256, size=(n, d), dtype=np.uint8)
    ids = np.arange(1, n + 1, dtype=np.int64) * 2
    with Corpus(d, capacity_hint=8) as c:
        for i in range(n):
            c.append(ids[i:i + 1], rows[i:i + 1])
            if i in (0, 7, 1023, 1024, 3000):
                assert len(c) == i + 1                                   # pending rows are part of the table
                got = c.search(rows[i], 5)[0]                            # ... and a search sees the row just appended
                assert got.ids[0] == ids[i]
        assert len(c) == n
        c.flush()
        r_ids, r_rows = c.read_rows(0, n)
        assert np.array_equal(r_ids, ids) and np.array_equal(r_rows, rows)
        q = rng.integers(0, 256, size=d, dtype=np.uint8)
        assert same(c.search(q, 50)[0], oracle.topk(rows, ids, q, 50, 1e3))
        st = c.stats()
        assert st.rows == n and st.capacity_rows >= n


def test_append_throughput_concurrent_with_searches():
    """200k single-row appends from one thread while another searches: the appends are coalesced (no per-row
    synchronisation) and growth maps memory instead of copying the shard, so neither side stalls the other."""
    rng = np.random.default_rng(4)
    n, d = 200_000, 256
    rows = rng.integers(0, 256, size=(n, d), dtype=np.uint8)
    ids = np.arange(1, n + 1, dtype=np.int64)
    L = nat.lib()
    with Corpus(d, capacity_hint=1024) as c:
        stop = threading.Event()
        lat = []

        def reader():
            q = rows[5]
            while not stop.is_set():
                t0 = time.perf_counter()
                r = c.search(q, 10)[0]
                lat.append(time.perf_counter() - t0)
                assert len(r.ids) == 0 or r.ids.max() <= n

        t = threading.Thread(target=reader)
        t.start()
        h = c.handle
        p_ids, p_rows = ids.ctypes.data, rows.ctypes.data
        import ctypes
        t0 = time.perf_counter()
        for i in range(n):
            rc = L.pbx_corpus_append(h, ctypes.c_void_p(p_ids + 8 * i), ctypes.c_void_p(p_rows + d * i), 1)
            assert rc == 0
        c.flush()
        dt = time.perf_counter() - t0
        stop.set()
        t.join()
        assert len(c) == n
        rate = n / dt
        lat = np.array(lat[2:]) if len(lat) > 4 else np.array(lat)
        print(f"single-row appends: {rate:.0f} rows/s; {len(lat)} concurrent searches, median {np.median(lat) * 1e3:.3f} ms, max {lat.max() * 1e3:.3f} ms")
        assert rate > 100_000, rate
        assert lat.max() < 0.05, lat.max()              # no search waited for a shard copy
        q = rows[77]
        assert same(c.search(q, 20)[0], oracle.topk(rows, ids, q, 20, 1e3, threads=4))


def test_device_resident_ingest_quantize_then_append():
    import torch
    rng = np.random.default_rng(6)
    n, d = 3000, 128
    emb = np.tanh(rng.normal(size=(n, d))).astype(np.float32)
    emb[0, :4] = [-1.0, 1.0, 0.0, 0.1]                                   # README.md:54
    d_emb = torch.from_numpy(emb).cuda()
    d_rows = torch.empty((n, d), dtype=torch.uint8, device="cuda")
    d_ids = torch.arange(10, 10 + n, dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    quantize_device(d_emb.data_ptr(), n * d, d_rows.data_ptr(), 0, s)
    host_rows = quantize(emb)
    assert list(host_rows[0, :4]) == [0x00, 0xFF, 0x80, 0x8C]
    with Corpus(d) as c:
        c.append(np.array([1, 2], np.int64), host_rows[:2])              # two coalesced rows first: order of arrival is kept
        c.append_device(d_ids.data_ptr(), d_rows.data_ptr(), n, s)
        assert len(c) == n + 2
        r_ids, r_rows = c.read_rows(2, n)
        assert np.array_equal(r_rows, host_rows) and np.array_equal(r_ids, np.arange(10, 10 + n))
        ids = np.concatenate([[1, 2], np.arange(10, 10 + n)]).astype(np.int64)
        allrows = np.concatenate([host_rows[:2], host_rows])
        assert same(c.search(host_rows[5], 10)[0], oracle.topk(allrows, ids, host_rows[5], 10, 1e3))


def test_growth_maps_memory_and_the_cudamalloc_fallback_still_works():
    with Corpus(32, capacity_hint=16) as c:
        assert c.stats().reserved == 1, "the shard should live in reserved address ranges (CUDA VMM)"
    code = ("import numpy as np\n"
            "from pixelbox_b200.corpus import Corpus\n"
            "rng = np.random.default_rng(1)\n"
            "rows = rng.integers(0, 256, size=(40000, 32), dtype=np.uint8)\n"
            "ids = np.arange(1, 40001, dtype=np.int64)\n"
            "with Corpus(32, capacity_hint=16) as c:\n"
            "    assert c.stats().reserved == 0\n"
            "    for b in range(0, 40000, 5000):\n"
            "        c.append(ids[b:b + 5000], rows[b:b + 5000])\n"
            "    r = c.search(rows[123], 3)[0]\n"
            "    assert r.ids[0] == 124 and len(c) == 40000\n"
            "print('fallback ok')\n")
    env = dict(os.environ, PBX_NO_VMM="1", PYTHONPATH=ROOT)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0 and "fallback ok" in res.stdout, res.stdout + res.stderr
