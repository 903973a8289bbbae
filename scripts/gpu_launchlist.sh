#!/bin/bash
# ncu launch list (per-kernel durations, cold-cache / serialised): scripts/gpu_launchlist.sh <tag>
TAG=$1
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-autoreset > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "ncu launches exit $?"
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_launches.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4].split("(")[0][:60]].append(float(r[-1]))
for k,v in agg.items(): print("%-62s n=%3d  mean %.1f us  last %.1f us" % (k, len(v), sum(v)/len(v)/1e3, v[-1]/1e3))
PY
