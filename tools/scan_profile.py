"""Experiment tool: per-CTA counters of the fast-pass scan (needs a -DPBX_EXP_PROFILE build, PBX_SO=...)."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelbox_b200 import _native as nat  # noqa: E402
from pixelbox_b200 import synth  # noqa: E402
from pixelbox_b200.corpus import Corpus  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 256
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
L = nat.lib()
L.pbx_debug_scan_profile.argtypes = [ctypes.c_void_p, ctypes.c_int]
c = Corpus(dim, capacity_hint=rows)
c.fill_synthetic(rows, 42, 0)
c.set_profiling(True)
q = synth.synth_queries(7, 8, dim, rows, 42)
for i in range(3):
    c.search(q[i], k)
L.pbx_debug_scan_profile(None, 1)
nrun = 5
for i in range(nrun):
    c.search(q[i], k)
prof = np.zeros((2048, 8), np.uint64)
L.pbx_debug_scan_profile(prof.ctypes.data, 0)
g = c.stats().scan_grid
p = prof[:g].astype(np.float64) / nrun
names = ["pushes", "compactions", "rendezvous_wait_cyc", "compaction_cyc", "total_cyc", "cnt_before_filter", "cnt_after_filter", "chunks"]
print(f"rows={rows} dim={dim} k={k} grid={g} scan_ms={c.stats().last_scan_ms:.4f}")
for i, nm in enumerate(names):
    print(f"  {nm:22s} mean={p[:, i].mean():12.1f} min={p[:, i].min():12.1f} max={p[:, i].max():12.1f}")

fin = np.zeros(16, np.int64)
L.pbx_debug_fin_profile.argtypes = [ctypes.c_void_p]
L.pbx_debug_fin_profile(fin.ctypes.data)
labels = ["prologue", "hist suffix scan", "gather", "rank sort", "(kappa stats)", "stage rows", "replay", "dist+ids", "order+output"]
print("  finalize phases (cycles @~1.9 GHz):")
for i in range(8):
    print(f"    {labels[i]:18s} {int(fin[i + 1] - fin[i]):8d}")
print(f"    {'total':18s} {int(fin[8] - fin[0]):8d}")
