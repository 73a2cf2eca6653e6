"""Multi-GPU parity check, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/shard_check.py

Every rank loads its block of the same table into its GPU shard, the ranks answer the same queries
through ShardedCorpus (local search -> NCCL all-gather of pbx_hit records -> device merge) and rank 0
compares with the oracle over the whole table; then the same on the synthetic corpus with the
host-buffer API and the device-resident API.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from pixelbox_b200 import _native as nat  # noqa: E402
from pixelbox_b200 import synth  # noqa: E402
from pixelbox_b200.shard import ShardedCorpus  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    # ---- a table with ties that straddle shards ------------------------------------------------
    rng = np.random.default_rng(17)
    n, d = 40_000, 64
    cent = rng.integers(0, 256, size=(40, d))
    rows = np.clip(cent[rng.integers(0, 40, n)] + rng.integers(-2, 3, size=(n, d)), 0, 255).astype(np.uint8)
    rows[100:400] = rows[100]
    rows[n - 300:] = rows[100]
    ids = np.arange(1, n + 1, dtype=np.int64) * 3
    queries = np.stack([rows[100], rows[12345], rng.integers(0, 256, d, dtype=np.uint8)])
    sc = ShardedCorpus(d, device=local)
    sc.load_table(ids, rows)
    assert sc.total_rows() == n
    for k, md in ((10, 1e3), (100, 1e3), (500, 1e3), (100, 0.02), (100, 1e7)):
        res = sc.search(queries, k, md)
        if rank == 0:
            for qi, q in enumerate(queries):
                o_ids, o_dist, o_dot, o_n2 = oracle.topk(rows, ids, q, k, md, threads=4)
                good = (list(res[qi].ids) == list(o_ids) and np.array_equal(res[qi].dist.view(np.uint32), o_dist.view(np.uint32))
                        and np.array_equal(res[qi].dot, o_dot) and np.array_equal(res[qi].norm2, o_n2))
                if not good:
                    print(f"MISMATCH table k={k} md={md} q={qi}")
                ok &= good
    sc.close()
    # ---- synthetic shards, device-resident API ---------------------------------------------------
    per, d, k = 300_000, 256, 100
    sc = ShardedCorpus(d, capacity_hint=per, device=local)
    sc.fill_synthetic(per, 42)
    queries = synth.synth_queries(5, 4, d, per * world, 42)
    res = sc.search(queries, k)
    # a batch large enough for the tensor-core path on every shard, then the all-gather + merge of [nq][k] records
    bq = synth.synth_queries(6, 48, d, per * world, 42)
    bres = sc.search(bq, k)
    dq = torch.from_numpy(queries).cuda()
    d_hits, d_cnt = sc.search_device(dq, len(queries), k)
    torch.cuda.synchronize()
    hits = d_hits.cpu().numpy().view(nat.HIT_DTYPE).reshape(len(queries), k)
    if rank == 0:
        full = synth.synth_rows(42, 0, per * world, d)
        f_ids = np.arange(1, per * world + 1, dtype=np.int64)
        for qi, q in enumerate(queries):
            o_ids, o_dist, _, _ = oracle.topk(full, f_ids, q, k, 1e3, threads=oracle.max_threads())
            good = list(res[qi].ids) == list(o_ids) and np.array_equal(res[qi].dist.view(np.uint32), o_dist.view(np.uint32))
            good &= np.array_equal(hits[qi]["image_id"], o_ids) and np.array_equal(hits[qi]["dist"].view(np.uint32), o_dist.view(np.uint32))
            if not good:
                print(f"MISMATCH synthetic q={qi}")
            ok &= good
        for qi in range(0, len(bq), 5):
            o_ids, o_dist, _, _ = oracle.topk(full, f_ids, bq[qi], k, 1e3, threads=oracle.max_threads())
            good = list(bres[qi].ids) == list(o_ids) and np.array_equal(bres[qi].dist.view(np.uint32), o_dist.view(np.uint32))
            if not good:
                print(f"MISMATCH batched q={qi}")
            ok &= good
        ok &= sc.local.stats().batched_queries >= len(bq)
    sc.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("shard_check", "OK" if ok else "FAILED", f"world={world}")
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
