"""The exchange step of the row-sharded path with every rank ON THE SAME GPU (so it runs on a 1-GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tools/exchange_check.py

One process per rank over gloo, every rank a GPU shard on cuda:0.  CUDA IPC maps the mailboxes across the
processes, so `exchange_merge_kernel` (peer stores + release/acquire flags + merge) runs exactly as it does
between GPUs; the fallback (all-gather of the records + `merge_hits_kernel`) is exercised beside it.  Rank 0
compares both with the oracle over the whole table: ties that straddle the shards, a batch large enough for
the tensor-core path, filters that leave fewer than k rows.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from pixelbox_b200 import _native as nat  # noqa: E402
from pixelbox_b200.shard import ShardedCorpus  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(0)
    dist.init_process_group("gloo")
    ok = True
    rng = np.random.default_rng(23)
    n, d = 60_000, 256
    cent = rng.integers(0, 256, size=(50, d))
    rows = np.clip(cent[rng.integers(0, 50, n)] + rng.integers(-3, 4, size=(n, d)), 0, 255).astype(np.uint8)
    rows[200:500] = rows[200]                    # one plateau of identical rows in shard 0 ...
    rows[n - 400:] = rows[200]                   # ... continued in the last shard: ties ordered by image_id across shards
    ids = rng.permutation(np.arange(1, n + 1)).astype(np.int64) * 5
    queries = np.concatenate([rows[[200, 31_000]], rng.integers(0, 256, size=(38, d), dtype=np.uint8)])
    for peer in (True, False):
        sc = ShardedCorpus(d, device=0, use_peer_exchange=peer)
        sc.load_table(ids, rows)
        used_peer = sc._exchange is not None
        if peer and not used_peer and rank == 0:
            print("note: CUDA IPC mapping unavailable, the peer path was not exercised")
        for k, md, qs in ((100, 1e3, queries), (10, 1e3, queries[:3]), (100, 0.01, queries[:3]), (500, 1e7, queries[:2]), (100, 1e3, queries[:1])):
            res = sc.search(qs, k, md)
            if rank == 0:
                for qi, q in enumerate(qs):
                    o_ids, o_dist, o_dot, o_n2 = oracle.topk(rows, ids, q, k, md, threads=4)
                    good = (list(res[qi].ids) == list(o_ids) and np.array_equal(res[qi].dist.view(np.uint32), o_dist.view(np.uint32))
                            and np.array_equal(res[qi].dot, o_dot) and np.array_equal(res[qi].norm2, o_n2))
                    if not good:
                        print(f"MISMATCH peer={used_peer} k={k} md={md} q={qi}")
                    ok &= good
        if rank == 0:
            print(f"exchange path {'peer-memory kernel' if used_peer else 'all-gather + merge kernel'}: {'ok' if ok else 'FAILED'}")
        sc.close()
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("exchange_check", "OK" if ok else "FAILED", f"world={world} (all ranks on cuda:0)")
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
