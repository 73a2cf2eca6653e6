#!/bin/bash
# compute-sanitizer over the reduced-size cases of tools/sanitize_case.py; summaries go to gpurun_out/sanitizer/.
#   tools/sanitize.sh [tool ...]      tools: memcheck racecheck synccheck initcheck (default: memcheck racecheck synccheck)
out=gpurun_out/sanitizer; mkdir -p $out
tools=${@:-memcheck racecheck synccheck}
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in $tools; do
  for c in single batched append misc sharded; do
    log=$out/${tool}_${c}.log
    timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 77 python tools/sanitize_case.py $c > $log 2>&1
    rc=$?
    echo "$tool $c rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|case .* ok' $log | tr '\n' ' ')"
  done
done | tee $out/summary.txt
