"""Developer tool: times the batched (tensor-core) path on a synthetic corpus, device-resident queries.
    python tools/batch_time.py [rows] [dim] [nq] [k] [iters]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from pixelbox_b200 import _native as nat  # noqa: E402
from pixelbox_b200 import synth  # noqa: E402
from pixelbox_b200.corpus import Corpus  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
k = int(sys.argv[4]) if len(sys.argv) > 4 else 100
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5

torch.cuda.init()
c = Corpus(dim, capacity_hint=rows)
c.fill_synthetic(rows, 42, 0)
queries = synth.synth_queries(43, nq, dim, rows, 42)
dq = torch.from_numpy(queries).cuda()
dh = torch.zeros(nq * k * 24, dtype=torch.uint8, device="cuda")
dc = torch.zeros(nq, dtype=torch.int32, device="cuda")
s = torch.cuda.Stream()
for _ in range(2):
    c.search_device(dq.data_ptr(), nq, k, 1e3, dh.data_ptr(), dc.data_ptr(), s.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(iters):
    c.search_device(dq.data_ptr(), nq, k, 1e3, dh.data_ptr(), dc.data_ptr(), s.cuda_stream)
e1.record(s)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
st = c.stats()
ops = 2.0 * rows * nq * dim
print(f"rows={rows} dim={dim} nq={nq} k={k} ms/batch={ms:.3f} qps={nq / ms * 1e3:.0f} int8_TOPS={ops / ms / 1e9:.1f} "
      f"corpus_GB/s_per_pass={rows * dim / ms / 1e6:.1f} batched={st.batched_queries} exact_passes={st.exact_passes}", flush=True)
# spot check against the oracle on regenerated rows
hits = dh.cpu().numpy().view(nat.HIT_DTYPE).reshape(nq, k)
ok = True
for qi in (0, nq // 2, nq - 1):
    ids = hits[qi]["image_id"]
    rows_back = np.concatenate([synth.synth_rows(42, int(i) - 1, 1, dim) for i in ids])
    o = oracle.topk(rows_back, ids, queries[qi], k, 1e3)
    ok &= list(o[0]) == list(ids) and np.array_equal(o[1].view(np.uint32), hits[qi]["dist"].view(np.uint32))
print("returned rows re-verify against the oracle:", "ok" if ok else "MISMATCH")
