#!/bin/bash
# Run a list of bench variants on the GPU box: scripts/gpu_multi.sh <tag> <spec file> [pytest: 0|1|expr]
# spec file lines:  name | ENV=val,ENV2=val (or -) | bench.py arguments
TAG=$1; SPEC=$2; PYT=${3:-0}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
lscpu | grep -E 'Model name|^CPU\(s\)' >> gpurun_out/${TAG}_gpu.txt
if [ "$PYT" != "0" ]; then
  if [ "$PYT" = "1" ]; then SEL=""; else SEL="-k $PYT"; fi
  timeout 1500 python -m pytest tests -m gpu -x -q $SEL > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -15 gpurun_out/${TAG}_pytest.log
fi
while IFS='|' read -r NAME ENVS ARGS; do
  NAME=$(echo $NAME | xargs); ENVS=$(echo $ENVS | xargs); [ -z "$NAME" ] && continue
  case "$NAME" in \#*) continue;; esac
  [ "$ENVS" = "-" ] && ENVS=""
  env $(echo $ENVS | tr ',' ' ') timeout 900 python bench.py $ARGS > gpurun_out/${TAG}_${NAME}.json 2> gpurun_out/${TAG}_${NAME}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_${NAME}.json")); r=d["roofline"]
    print("${NAME}: value %.0f e2e %.0f (%.0f%%) ms/step %.4f flow_ms %.4f frac %.3f fin_ms %.4f live %.1f" % (d["value"], d["e2e"]["value"], 100*d["e2e"]["value"]/d["value"], d["ms_per_step"], r["ms_per_launch"], r["frac"], r["finish_kernel_ms"], r["live_stations_per_env_farm"]))
    if "with_autoreset" in d: print("   autoreset: %.0f" % d["with_autoreset"]["value"], d["with_autoreset"]["pool"])
    for k,v in d.get("configs",{}).items(): print("   %s: value %.0f e2e %.0f frac %.3f flow_ms %.4f fin_ms %.4f" % (k, v["value"], v["e2e"], v["roofline_frac"], v["ms_per_launch"], v["finish_kernel_ms"]))
except Exception as e:
    print("${NAME}: failed", e)
PY
  tail -3 gpurun_out/${TAG}_${NAME}.err
done < $SPEC
