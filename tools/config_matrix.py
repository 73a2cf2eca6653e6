"""BASELINE configs[4] as seen by ONE GPU of the 8-GPU layout: a 12.5M-row shard (100M rows / 8) for
d in {64, 256, 1024} and k in {10, 100, 1000}: single-query device-resident time, scan-kernel time, and the
1024-query batch where the tensor-core path applies.  Prints one JSON line per cell.
    python tools/config_matrix.py [rows_per_gpu]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from pixelbox_b200 import _native as nat  # noqa: E402
from pixelbox_b200 import synth  # noqa: E402
from pixelbox_b200.corpus import Corpus  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
torch.cuda.init()
stream = torch.cuda.Stream()


def timed(fn, iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(iters):
        fn(i)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for dim in (64, 256, 1024):
    c = Corpus(dim, capacity_hint=rows)
    c.fill_synthetic(rows, 42, 0)
    nq = 64
    queries = synth.synth_queries(7, nq, dim, rows, 42)
    dq = torch.from_numpy(queries).cuda()
    bq = synth.synth_queries(9, 1024, dim, rows, 42)
    dbq = torch.from_numpy(bq).cuda()
    for k in (10, 100, 1000):
        dh = torch.zeros(1024 * k * 24, dtype=torch.uint8, device="cuda")
        dc = torch.zeros(1024, dtype=torch.int32, device="cuda")

        def one(i):
            q = i % nq
            c.search_device(dq.data_ptr() + q * dim, 1, k, 1e3, dh.data_ptr() + q * k * 24, dc.data_ptr() + 4 * q, stream.cuda_stream)
        timed(one, 10)
        ms = timed(one, 100)
        c.set_profiling(True)
        scan = []
        for i in range(10):
            c.search(queries[i % nq], k)
            scan.append(c.stats().last_scan_ms)
        c.set_profiling(False)
        scan_ms = float(np.median(scan))
        # parity of one query against the oracle on the returned rows (regenerated) -- full-corpus parity is the tests' job
        hits = dh.cpu().numpy().view(nat.HIT_DTYPE).reshape(1024, k)[0]
        ids = hits["image_id"]
        back = np.concatenate([synth.synth_rows(42, int(i) - 1, 1, dim) for i in ids[:50]])
        o = oracle.topk(back, ids[:50], queries[0], 50, 1e3)
        ok = list(o[0]) == list(ids[:50]) and np.array_equal(o[1].view(np.uint32), hits["dist"][:50].view(np.uint32))
        cell = {"rows": rows, "dim": dim, "k": k, "ms_per_query": round(ms, 4), "queries_per_sec": round(1e3 / ms, 1),
                "corpus_GB_per_s": round(rows * dim / ms / 1e6, 1), "scan_ms": round(scan_ms, 4),
                "scan_GB_per_s": round(rows * dim / scan_ms / 1e6, 1), "exact_passes": c.stats().exact_passes,
                "returned_rows_match_oracle": bool(ok)}
        before = c.stats().batched_queries

        def batch(i):
            c.search_device(dbq.data_ptr(), 1024, k, 1e3, dh.data_ptr(), dc.data_ptr(), stream.cuda_stream)
        timed(batch, 1)
        bms = timed(batch, 3)
        if c.stats().batched_queries > before:
            cell["batch1024_ms"] = round(bms, 3)
            cell["batch1024_queries_per_sec"] = round(1024e3 / bms)
            cell["batch1024_int8_tops"] = round(2.0 * rows * 1024 * dim / bms / 1e9, 1)
        else:
            cell["batch1024_ms"] = round(bms, 3)
            cell["batch1024_path"] = "1024 single-query scans (tensor-core path needs dim % 128 == 0 and keep * 8 <= 4096)"
        print(json.dumps(cell), flush=True)
    c.close()
