#!/bin/bash
# Developer tool: per-launch times of one batched search (ncu launch list) for a given environment, e.g.
#   PBX_BATCH_CG=1 tools/batch_rounds.sh tag [rows] [dim] [nq]
tag=$1; rows=${2:-10000000}; dim=${3:-256}; nq=${4:-1024}
mkdir -p gpurun_out/rounds
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rounds/$tag.csv python tools/batch_time.py $rows $dim $nq 100 1 > gpurun_out/rounds/$tag.log 2>&1
echo "== $tag rc=$?"; grep "ms/batch\|oracle" gpurun_out/rounds/$tag.log
python - <<PY
import csv
lines=[l for l in open('gpurun_out/rounds/$tag.csv') if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if r.get("Metric Name")=="gpu__time_duration.sum" and 'batch' in r["Kernel Name"]]
rows=rows[len(rows)*2//3:]
print(" ".join(f"{r['Kernel Name'].split('::')[1][:14]}={float(r['Metric Value'].replace(',',''))/1e3:.0f}" for r in rows))
PY
