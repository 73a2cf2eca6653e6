# Round-end single-GPU pass: full GPU test suite, the default bench line, its ncu launch list, and one ncu --set full
# capture of the batched kernel's last (largest) round.  Outputs under gpurun_out/r1f/.
mkdir -p gpurun_out/r1f
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r1f/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r1f/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r1f/bench_n1.json 2> gpurun_out/r1f/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1f/bench_reference.json 2> gpurun_out/r1f/bench_reference.err; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1f/launches_bench_steps20.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r1f/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:batch_mma -s 5 -c 1 -f -o gpurun_out/r1f/batch_mma_full python tools/batch_time.py 10000000 256 1024 100 1 > gpurun_out/r1f/batch_full.log 2>&1; echo "ncu full rc=$?"
cat gpurun_out/r1f/bench_n1.json
