#!/bin/bash
# A/B a list of env-var settings on the bench: scripts/gpu_ab.sh <tag> "VAR=a VAR=b ..." [pytest: 0|1]
TAG=$1; shift
SETTINGS=$1; shift
mkdir -p gpurun_out
if [ "${1:-0}" = "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
fi
i=0
for S in $SETTINGS; do
  i=$((i+1))
  env $(echo $S | tr ',' ' ') timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu --no-autoreset > gpurun_out/${TAG}_bench_$i.json 2> gpurun_out/${TAG}_bench_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_$i.json")); r=d["roofline"]
    print("$S: value %.0f e2e %.0f flow_ms %.4f frac %.3f fin_ms %.4f" % (d["value"], d["e2e"]["value"], r["ms_per_launch"], r["frac"], r["finish_kernel_ms"]))
except Exception as e:
    print("$S: no bench json", e)
PY
  tail -2 gpurun_out/${TAG}_bench_$i.err
done
