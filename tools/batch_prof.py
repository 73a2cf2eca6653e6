"""Developer tool: role counters of batch_mma_kernel from a -DPBX_BATCH_PROF build.
    PBX_NVCC_EXTRA="-DPBX_BATCH_PROF" PBX_SO_OUT=pixelbox_b200/lib/exp/lib_bprof.so python -m pixelbox_b200.build --force
    PBX_SO=pixelbox_b200/lib/exp/lib_bprof.so python tools/batch_prof.py [rows] [dim] [nq]
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelbox_b200 import _native as nat  # noqa: E402
from pixelbox_b200 import synth  # noqa: E402
from pixelbox_b200.corpus import Corpus  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
k = 100
torch.cuda.init()
L = nat.lib()
c = Corpus(dim, capacity_hint=rows)
c.fill_synthetic(rows, 42, 0)
q = synth.synth_queries(43, nq, dim, rows, 42)
dq = torch.from_numpy(q).cuda()
dh = torch.zeros(nq * k * 24, dtype=torch.uint8, device="cuda")
dc = torch.zeros(nq, dtype=torch.int32, device="cuda")
s = torch.cuda.Stream()
c.search_device(dq.data_ptr(), nq, k, 1e3, dh.data_ptr(), dc.data_ptr(), s.cuda_stream)
torch.cuda.synchronize()
buf = np.zeros((2048, 12), np.uint64)
L.pbx_debug_batch_profile(None, 1)
c.search_device(dq.data_ptr(), nq, k, 1e3, dh.data_ptr(), dc.data_ptr(), s.cuda_stream)
torch.cuda.synchronize()
L.pbx_debug_batch_profile(ctypes.c_void_p(buf.ctypes.data), 0)
commit = (buf[:148, 10] >> np.uint64(32)).astype(np.float64)      # MMA role: cycles in the per-stage commits (high half of column 10)
buf[:, 10] &= np.uint64(0xFFFFFFFF)
b = buf[:148].astype(np.float64)
names = ["prod wait a_empty", "prod wait m_empty", "mma wait acc_empty", "mma wait a_full", "mma total", "epi wait acc_full", "epi ld+arrive",
         "epi process", "epi wait m_full", "epi total", "stages", "mma issue (UTCIMMA)"]
lead = b[0::2]
print("per CTA means (cycles), seed + main pass of one search; leaders only for the mma rows")
for i, nme in enumerate(names):
    src = lead if nme.startswith("mma") else b
    print(f"  {nme:22s} {src[:, i].mean():14.0f}   per stage {src[:, i].mean() / max(1.0, b[:, 10].mean()):8.1f}")
print(f"  {'mma a_empty commits':22s} {commit[commit > 0].mean():14.0f}   per stage {commit[commit > 0].mean() / max(1.0, b[:, 10].mean()):8.1f}")
