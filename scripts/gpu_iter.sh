#!/bin/bash
# Quick GPU iteration: parity tests, bench for both tile-buffer modes, one full ncu capture of the flow kernel.
# usage: scripts/gpu_iter.sh <tag> [ncu: 0|1]
TAG=${1:-it}
NCU=${2:-1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
RC=$?
echo "pytest exit $RC" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
for ST in 1 2; do
  WG_FLOW_STAGES=$ST timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/${TAG}_bench_s${ST}.json 2> gpurun_out/${TAG}_bench_s${ST}.err
  echo "bench stages=$ST exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_s${ST}.json"))
    r=d["roofline"]
    print("value %.0f e2e %.0f flow_ms %.4f frac %.3f fin_ms %.4f live %.1f" % (d["value"], d["e2e"]["value"], r["ms_per_launch"], r["frac"], r["finish_kernel_ms"], r["live_stations_per_env_farm"]))
except Exception as e:
    print("no bench json", e)
PY
  tail -3 gpurun_out/${TAG}_bench_s${ST}.err
done
if [ "$NCU" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:wg_flow_kernel -s 30 -c 1 \
    -f -o gpurun_out/${TAG}_flow python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
