#!/bin/bash
# Quick GPU iteration: parity tests, bench per flow-kernel variant, one full ncu capture of the flow kernel per variant.
# usage: scripts/gpu_iter.sh <tag> "<variants>" [ncu: 0|1]
TAG=${1:-it}
VARS=${2:-"0 1"}
NCU=${3:-1}
mkdir -p gpurun_out
for V in $VARS; do
  export WG_FLOW_VARIANT=$V
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_v${V}.log 2>&1
  echo "pytest variant=$V exit $?" >> gpurun_out/${TAG}_pytest_v${V}.log
  tail -4 gpurun_out/${TAG}_pytest_v${V}.log
  timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/${TAG}_bench_v${V}.json 2> gpurun_out/${TAG}_bench_v${V}.err
  echo "bench variant=$V exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_v${V}.json"))
    r=d["roofline"]
    print("value %.0f e2e %.0f flow_ms %.4f frac %.3f fin_ms %.4f live %.1f" % (d["value"], d["e2e"]["value"], r["ms_per_launch"], r["frac"], r["finish_kernel_ms"], r["live_stations_per_env_farm"]))
except Exception as e:
    print("no bench json", e)
PY
  tail -3 gpurun_out/${TAG}_bench_v${V}.err
  if [ "$NCU" = "1" ]; then
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:wg_flow_kernel -s 30 -c 1 \
      -f -o gpurun_out/${TAG}_flow_v${V} python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full_v${V}.log 2>&1
    echo "ncu full exit $?"
  fi
done
