"""Developer tool: times pbx_merge_hits_device for [n_shards][nq][k] sorted lists on one GPU."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelbox_b200 import _native as nat  # noqa: E402

for n_shards, nq, k in ((2, 1, 100), (8, 1, 100), (8, 1, 1000), (8, 1024, 100), (64, 1, 2048)):
    rng = np.random.default_rng(1)
    g = np.zeros((n_shards, nq, k), nat.HIT_DTYPE)
    g["dist"] = np.sort(rng.random((n_shards, nq, k), dtype=np.float32), axis=2)
    g["image_id"] = rng.integers(1, 1 << 40, size=(n_shards, nq, k))
    d_g = torch.from_numpy(g.view(np.uint8).reshape(-1)).cuda()
    d_out = torch.zeros(nq * k * 24, dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(nq, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream or 1

    def run():
        nat.check(nat.lib().pbx_merge_hits_device(0, d_g.data_ptr(), None, n_shards, nq, k, d_out.data_ptr(), d_cnt.data_ptr(), s))
    for _ in range(5):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"merge n_shards={n_shards} nq={nq} k={k}: {e0.elapsed_time(e1) / 50 * 1000:.1f} us/launch", flush=True)
