"""Developer tool: times the device-resident single-query path on a synthetic corpus.
    python tools/quick_time.py [rows] [dim] [k] [iters] [ctas_per_sm,ctas_per_sm,...]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelbox_b200 import synth  # noqa: E402
from pixelbox_b200.corpus import Corpus  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 256
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 100
cps_list = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0]

torch.cuda.init()
c = Corpus(dim, capacity_hint=rows)
c.fill_synthetic(rows, 42, 0)
nq = 64
queries = synth.synth_queries(7, nq, dim, rows, 42)
dq = torch.from_numpy(queries).cuda()
dh = torch.zeros(nq * k * 24, dtype=torch.uint8, device="cuda")
dc = torch.zeros(nq, dtype=torch.int32, device="cuda")
s = torch.cuda.Stream()


def one(i):
    q = i % nq
    c.search_device(dq.data_ptr() + q * dim, 1, k, 1e3, dh.data_ptr() + q * k * 24, dc.data_ptr() + 4 * q, s.cuda_stream)


for cps in cps_list:
    c.set_scan_ctas_per_sm(cps)
    for i in range(10):
        one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for i in range(iters):
        one(i)
    e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    gbs = rows * dim / ms / 1e6
    scan = []
    c.set_profiling(True)
    for i in range(20):
        c.search(queries[i % nq], k)
        scan.append(c.stats().last_scan_ms)
    c.set_profiling(False)
    scan_ms = float(np.median(scan))
    st = c.stats()
    print(f"rows={rows} dim={dim} k={k} ctas/sm={cps} grid={st.scan_grid} ms/query={ms:.4f} qps={1000/ms:.1f} "
          f"corpus_GB/s={gbs:.1f} scan_ms={scan_ms:.4f} scan_GB/s={rows*dim/scan_ms/1e6:.1f} exact_passes={st.exact_passes}", flush=True)
