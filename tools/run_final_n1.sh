# Round-end single-GPU pass: full GPU test suite, smoke(), the default bench line, its ncu launch list, the config matrix.
# Outputs under gpurun_out/r1h/.
mkdir -p gpurun_out/r1h
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r1h/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r1h/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r1h/bench_n1.json 2> gpurun_out/r1h/bench_n1.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1h/launches_bench_steps20.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r1h/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 python tools/config_matrix.py > gpurun_out/r1h/config_matrix.jsonl 2> gpurun_out/r1h/config_matrix.err; echo "matrix rc=$?"
cat gpurun_out/r1h/bench_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_query'], 'scan frac', d['roofline']['frac'], 'batched', d['batched']['ms_per_batch'], d['parity_check'], d['batched']['parity_check'], d['clocks'])"
python - <<'PY'
import json
for l in open("gpurun_out/r1h/config_matrix.jsonl"):
    c = json.loads(l)
    print(c["dim"], c["k"], c["ms_per_query"], c["scan_ms"], c.get("batch1024_ms"), c.get("batch1024_int8_tops"), c["returned_rows_match_oracle"])
PY
