# 8-GPU check: parity of the sharded paths, then the weak-scaling bench line (single query + the 1024-query batch).
mkdir -p gpurun_out/r1g
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29611 tools/shard_check.py > gpurun_out/r1g/shard_check_n8.log 2>&1; echo "shard_check rc=$?"; tail -1 gpurun_out/r1g/shard_check_n8.log
timeout 300 $TR --nproc-per-node 8 --master-port 29628 bench.py --gpus 8 --steps 200 --warmup 20 --no-cpu-baseline 2> gpurun_out/r1g/bench_n8.err | tail -1 > gpurun_out/r1g/bench_n8.json; echo "bench N=8 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r1g/bench_n8.json").read())
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_query"], d["parity_check"], d["exchange"])
print(d["batched"])
PY
