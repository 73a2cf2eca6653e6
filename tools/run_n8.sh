# 8-GPU check of the exchange step: parity, then the weak-scaling bench with the peer-memory exchange and with NCCL.
mkdir -p gpurun_out/r1e
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29611 tools/shard_check.py > gpurun_out/r1e/shard_check_n8.log 2>&1; echo "shard_check rc=$?"
timeout 200 $TR --nproc-per-node 8 --master-port 29628 bench.py --gpus 8 --steps 200 --warmup 20 --no-batched --no-cpu-baseline > gpurun_out/r1e/bench_n8_peer.json 2> gpurun_out/r1e/bench_n8_peer.err; echo "peer N=8 rc=$?"
PBX_NO_PEER_EXCHANGE=1 timeout 200 $TR --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --steps 200 --warmup 20 --no-batched --no-cpu-baseline > gpurun_out/r1e/bench_n8_nccl.json 2> gpurun_out/r1e/bench_n8_nccl.err; echo "nccl N=8 rc=$?"
tail -1 gpurun_out/r1e/shard_check_n8.log
python - <<'PY'
import json
for f in ("peer", "nccl"):
    d = json.loads(open(f"gpurun_out/r1e/bench_n8_{f}.json").read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_query"], d["exchange"])
PY
