#!/bin/bash
# Run on the GPU box via gpurun: parity tests, bench, ncu launch list + one full capture of the flow kernel.
# usage: scripts/gpu_check.sh <tag> [skip_tests]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
lscpu | grep -E 'Model name|^CPU\(s\)' >> gpurun_out/${TAG}_gpu.txt
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -5 gpurun_out/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 60 --warmup 5 --reward Baseline --no-cpu --no-autoreset > gpurun_out/${TAG}_bench_baseline.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_baseline.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-autoreset > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wg_flow_kernel -s 30 -c 2 \
  -f -o gpurun_out/${TAG}_flow python bench.py --steps 8 --warmup 3 --no-cpu --no-autoreset > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
