#!/bin/bash
# Round-2 evidence pass on one GPU: launch list of the default bench command, ncu --set full of the scan kernel and of
# the batched main pass.  Outputs under gpurun_out/r2prof/ (summarised into profiles/ by tools/ncu_summary.py + launch_summary.py).
out=gpurun_out/r2prof; mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bench_steps20.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-c1 > $out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
python tools/launch_summary.py $out/launches_bench_steps20.csv > $out/launches_summary.txt; cat $out/launches_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 5 -c 3 -o $out/scan_full python tools/quick_time.py 10000000 256 100 20 > $out/scan_full.log 2>&1; echo "ncu scan rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:batch_mma_kernel -s 2 -c 2 -o $out/batch_full python tools/batch_time.py 10000000 256 1024 100 1 > $out/batch_full.log 2>&1; echo "ncu batch rc=$?"
