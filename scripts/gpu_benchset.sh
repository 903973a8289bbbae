#!/bin/bash
# Secondary bench lines for DESIGN.md: scripts/gpu_benchset.sh <tag>
TAG=${1:-set}
mkdir -p gpurun_out
run() { # name args...
  local name=$1; shift
  timeout 900 python bench.py --steps 60 --warmup 6 --no-cpu --no-autoreset "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_${name}.json")); r=d["roofline"]
    print("${name}: value %.0f e2e %.0f ms/step %.4f flow_ms %.4f frac %.3f fin_ms %.4f live %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["ms_per_launch"], r["frac"], r["finish_kernel_ms"], r["live_stations_per_env_farm"]))
except Exception as e:
    print("${name}: failed", e)
PY
  tail -2 gpurun_out/${TAG}_${name}.err
}
run cfg2_power_avg
run cfg2_baseline --reward Baseline
run cfg3_512envs --envs 512
run cfg4_8x8_1024 --nx 8 --ny 8 --envs 1024
run cfg5_4x2_2048 --nx 4 --ny 2 --envs 2048
run cfg2_mann --turbtype Mann
