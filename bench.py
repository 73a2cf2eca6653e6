#!/usr/bin/env python
"""bench.py -- the similarity-search hot path on synthetic corpora (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is ONE single-query top-100 scan of the whole corpus (configs[1]: 10M x 256-byte rows
per GPU, k = 100).  With N > 1 the corpus is row-sharded (10M rows on every GPU, N x 10M in
total: weak scaling), every rank scans its shard and the k records per shard are exchanged by
one NCCL all-gather and merged on the device.  The 2.56 GB per GPU exceed the 126 MB L2, so no
L2 flush is needed between steps.

value  = corpus GB/s scanned by the whole job with the query already in HBM (queries/s beside it)
e2e    = the same through the host-buffer API: pinned host query -> H2D -> search -> D2H result
--impl reference = the CPU restatement of the reference's scan (oracle/, the reference itself
         needs a Rust toolchain this image does not have) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 42
NQ = 64          # distinct queries cycled through


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows-per-gpu", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true", help="skip the 1024-query tensor-core measurement")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm (oracle): bounded sample of the same workload
# ---------------------------------------------------------------------------------------------
def cpu_scan(rows: int, dim: int, k: int, queries: np.ndarray, threads: int, steps: int, warmup: int):
    """Times oracle.topk (the restatement of src/engine.rs:572-588 + :375-383 as a bare loop) over the
    first `rows` rows of the synthetic corpus.  Returns (seconds per step, checksum of ids)."""
    from oracle import oracle
    from pixelbox_b200 import synth
    corpus = np.empty((rows, dim), np.uint8)
    for r0 in range(0, rows, 1 << 20):                  # C generator of the oracle library, chunked
        r1 = min(rows, r0 + (1 << 20))
        corpus[r0:r1] = oracle.synth_rows(SEED, r0, r1 - r0, dim)
    assert np.array_equal(corpus[:8], synth.synth_rows(SEED, 0, 8, dim))
    ids = np.arange(1, rows + 1, dtype=np.int64)
    for i in range(warmup):
        oracle.topk(corpus, ids, queries[i % len(queries)], k, 1e3, threads=threads)
    t0 = time.perf_counter()
    chk = 0
    for i in range(steps):
        o_ids, _, _, _ = oracle.topk(corpus, ids, queries[i % len(queries)], k, 1e3, threads=threads)
        chk ^= int(o_ids.sum())
    return (time.perf_counter() - t0) / steps, chk


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle
    from pixelbox_b200 import synth
    threads = oracle.max_threads()
    dim, k = args.dim, args.k
    # ~0.77 us per row per thread at d=256: size the sample so a step takes ~0.25 s
    rows = args.cpu_rows or int(min(args.rows_per_gpu, max(100_000, 300_000 * threads * 256 // max(dim, 1))))
    steps = max(1, min(args.steps, 40))
    warmup = max(1, min(args.warmup, 3))
    queries = synth.synth_queries(7, NQ, dim, args.rows_per_gpu * args.gpus, SEED)
    sec, _ = cpu_scan(rows, dim, k, queries, threads, steps, warmup)
    gbs = rows * dim / sec / 1e9
    sample = (f"first {rows} of {args.rows_per_gpu * args.gpus} rows x {dim} B, {steps} single-query top-{k} scans, "
              f"oracle/pbx_oracle.c (C restatement of src/engine.rs:572-588 + :375-383; upstream is 1 thread, "
              f"this arm splits rows over {threads} pthreads)")
    line = {
        "impl": "reference", "metric": "corpus_GB_per_s_scanned", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "queries_per_sec_on_sample": 1.0 / sec,
        "queries_per_sec_full_corpus_extrapolated": gbs * 1e9 / (args.rows_per_gpu * args.gpus * dim),
        "config": workload_config(args),
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


TRAFFIC_CSV = "r1i_scan_kernel_ncu_full_summary.csv"      # ncu --set full of the scan kernel, last build of round 1


def ncu_traffic(rows: int, dim: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the scan kernel per launch, from the committed ncu --set full
    capture of this workload (profiles/); None for any other workload."""
    if rows != 10_000_000 or dim != 256:
        return None
    path = os.path.join(ROOT, "profiles", TRAFFIC_CSV)
    try:
        import csv
        with open(path) as f:
            r = list(csv.reader(f))
        hdr, units, first = r[0], r[1], r[2]
        rd = float(first[hdr.index("dram__bytes_read.sum")]) * (1e9 if units[hdr.index("dram__bytes_read.sum")].startswith("G") else 1e6)
        wr = float(first[hdr.index("dram__bytes_write.sum")]) * (1e9 if units[hdr.index("dram__bytes_write.sum")].startswith("G") else 1e6)
        return rd + wr
    except Exception:
        return None


def workload_config(args):
    return {"workload": f"{args.rows_per_gpu // 1_000_000}M x {args.dim}-byte corpus per GPU, single query top-{args.k} "
                        f"(BASELINE configs[1]); row-sharded x{args.gpus}",
            "rows_per_gpu": args.rows_per_gpu, "rows_total": args.rows_per_gpu * args.gpus, "dim": args.dim, "k": args.k,
            "max_dist": 1e3, "distinct_queries": NQ, "parallelism": f"row-shard x{args.gpus}",
            "l2": "corpus shard (2.56 GB) >> 126 MB L2, no flush between steps"}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pixelbox_b200 import _native as nat
    from pixelbox_b200 import synth
    from pixelbox_b200.corpus import Corpus
    from pixelbox_b200.shard import ShardedCorpus

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if rank == 0:
            sys.stderr.write(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run\n")
        if world == 1 and args.gpus > 1:
            return 2
    nat.lib()                                     # fails loudly if the CUDA library is missing
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: pixelbox_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dim, k, rows = args.dim, args.k, args.rows_per_gpu
    total_rows = rows * world
    queries = synth.synth_queries(7, NQ, dim, total_rows, SEED)

    if world > 1:
        sc = ShardedCorpus(dim, capacity_hint=rows, device=local_rank, use_peer_exchange=os.environ.get("PBX_NO_PEER_EXCHANGE") is None)
        sc.fill_synthetic(rows, SEED)
        corpus = sc.local
    else:
        sc = None
        corpus = Corpus(dim, capacity_hint=rows, device=local_rank)
        corpus.fill_synthetic(rows, SEED, 0)

    stream = torch.cuda.Stream()
    dq = torch.from_numpy(queries).cuda()
    d_hits = torch.zeros(NQ * k * 24, dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(NQ, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(i):
        q = i % NQ
        if sc is None:
            corpus.search_device(dq.data_ptr() + q * dim, 1, k, 1e3, d_hits.data_ptr() + q * k * 24, d_cnt.data_ptr() + 4 * q,
                                 stream.cuda_stream)
        else:
            sc.search_device(dq[q], 1, k, 1e3)

    def step_e2e(i):
        q = queries[i % NQ]
        if sc is None:
            return corpus.search(q, k, 1e3)[0]
        return sc.search(q, k, 1e3)[0]

    # ---- value: device-resident, back to back ---------------------------------------------------
    launches_per_step = 3 + (1 if world > 1 else 0)   # prep+seed, scan, finalize (+ merge); the exact pass is device-launched on demand
    with torch.cuda.stream(stream):
        for i in range(max(args.warmup, 3)):
            step_device(i)
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(args.steps):
            step_device(i)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        # ---- e2e: host buffers through the public API --------------------------------------------
        for i in range(3):
            step_e2e(i)
        barrier()
        t0 = time.perf_counter()
        last = None
        for i in range(args.steps):
            last = step_e2e(i)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        barrier()
        clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])

    # ---- dominant kernel: per-launch duration of the scan, CUDA events on its own stream ---------
    scan_ms = []
    corpus.set_profiling(True)
    for i in range(min(args.steps, 50)):
        corpus.search(queries[i % NQ], k, 1e3)
        scan_ms.append(corpus.stats().last_scan_ms)
    corpus.set_profiling(False)
    scan_ms = float(np.mean(scan_ms))
    st = corpus.stats()

    # ---- configs[2]: a batch of 1024 queries through the tensor-core path (reported beside the headline) ---------
    batched = None
    if dim % 128 == 0 and dim <= 1024 and not args.no_batched:
        nqb = 1024
        bq = synth.synth_queries(43, nqb, dim, total_rows, SEED)
        d_bq = torch.from_numpy(bq).cuda()
        d_bh = torch.zeros(nqb * k * 24, dtype=torch.uint8, device="cuda")
        d_bc = torch.zeros(nqb, dtype=torch.int32, device="cuda")

        def batch_step():
            if sc is None:
                corpus.search_device(d_bq.data_ptr(), nqb, k, 1e3, d_bh.data_ptr(), d_bc.data_ptr(), stream.cuda_stream)
                return d_bh
            return sc.search_device(d_bq, nqb, k, 1e3)[0]       # every shard contracts its rows against all queries, then the exchange

        with torch.cuda.stream(stream):
            for _ in range(2):
                batch_step()
            barrier()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            b0.record(stream)
            for _ in range(reps):
                out_b = batch_step()
            b1.record(stream)
            barrier()
        bt = torch.tensor([b0.elapsed_time(b1) / reps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(bt, op=dist.ReduceOp.MAX)
        bms = float(bt[0])
        bh = out_b.cpu().numpy().view(nat.HIT_DTYPE).reshape(nqb, k)
        from oracle import oracle as _orc
        bok = True
        if rank == 0:
            for qi in (0, 511, 1023):
                rb = np.concatenate([synth.synth_rows(SEED, int(i) - 1, 1, dim) for i in bh[qi]["image_id"]])
                o = _orc.topk(rb, bh[qi]["image_id"], bq[qi], k, 1e3)
                bok &= list(o[0]) == list(bh[qi]["image_id"]) and np.array_equal(o[1].view(np.uint32), bh[qi]["dist"].view(np.uint32))
        tops = 2.0 * total_rows * nqb * dim / (bms * 1e-3) / 1e12
        batched = {"workload": f"{rows // 1_000_000}M x {dim}-byte corpus per GPU x{world}, batch of {nqb} queries top-{k} (BASELINE configs[2])",
                   "ms_per_batch": bms, "queries_per_sec": nqb / (bms * 1e-3), "int8_tops": tops,
                   "frac_of_nominal_int8_peak": tops / (4500.0 * world), "peak_note": "nominal 4.5 POPS dense int8 per GPU (no measured int8 peak on file)",
                   "kernel": "batch_mma_kernel (tcgen05.mma.kind::i8, TMA, TMEM) + fused top-k epilogue",
                   "parity_check": "ok" if bok else "MISMATCH"}

    # ---- correctness of what was timed: the last e2e result against the oracle on its own rows ---
    check = "skipped"
    if rank == 0 and last is not None and len(last.ids):
        from oracle import oracle
        rows_back = np.concatenate([synth.synth_rows(SEED, int(i) - 1, 1, dim) for i in last.ids])
        o = oracle.topk(rows_back, last.ids, queries[(args.steps - 1) % NQ], k, 1e3)
        okay = list(o[0]) == list(last.ids) and np.array_equal(o[1].view(np.uint32), last.dist.view(np.uint32))
        check = "ok" if okay else "MISMATCH"

    if rank == 0:
        peak, peak_src = peaks()
        ms_step = ms / args.steps
        bytes_step = total_rows * dim
        value = bytes_step / (ms_step * 1e-3) / 1e9
        e2e_step = e2e_ms / args.steps
        achieved = rows * dim / (scan_ms * 1e-3) / 1e9
        line = {
            "metric": "corpus_GB_per_s_scanned", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 x s16 -> int32 (IDP.2A), f32 replay on candidates", "data": "synthetic",
            "queries_per_sec": 1e3 / ms_step,
            "config": workload_config(args),
            "e2e": {"value": bytes_step / (e2e_step * 1e-3) / 1e9, "unit": "GB/s", "queries_per_sec": 1e3 / e2e_step,
                    "ms_per_query": e2e_step, "h2d_bytes_per_step": dim, "d2h_bytes_per_step": k * 24 + 4,
                    "api": "pbx_search (C ABI, host buffers)" if world == 1 else "ShardedCorpus.search (host buffers; exchange as in 'exchange')"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "kernel": "scan_kernel<16,1,false,3>" if dim == 256 else "scan_kernel", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(rows, dim),
                         "traffic_source": f"profiles/{TRAFFIC_CSV} (ncu --set full, bytes per launch)",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": rows * dim, "launch_ms": scan_ms,
                         "share_of_step": scan_ms / ms_step},
            "clocks": clocks,
            "exact_passes": int(st.exact_passes), "scan_grid": int(st.scan_grid), "parity_check": check,
            "exchange": None if sc is None else ("one kernel over NVLink peer memory (post + signal + wait + merge)" if sc._exchange is not None
                                                 else "NCCL all-gather + merge kernel"),
        }
        if batched is not None:
            line["batched"] = batched
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle
            cpu_rows = args.cpu_rows or min(rows, 2_000_000 * 256 // dim)
            cpu_steps = 8
            sec, _ = cpu_scan(cpu_rows, dim, k, queries, 1, cpu_steps, 1)
            line["cpu_baseline"] = {
                "value": cpu_rows * dim / sec / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                "queries_per_sec_full_corpus_extrapolated": 1.0 / (sec * rows / cpu_rows),
                "host_cores_available": oracle.max_threads(),
                "sample": f"first {cpu_rows} of {rows} rows x {dim} B, {cpu_steps} single-query top-{k} scans, 1 thread "
                          f"(the reference scans inside one SQLite statement on one thread), oracle/pbx_oracle.c"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
