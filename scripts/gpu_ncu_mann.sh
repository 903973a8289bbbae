#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wg_flow_kernel -s 40 -c 1 -f -o gpurun_out/${TAG}_flow python bench.py --steps 8 --warmup 3 --no-cpu --no-autoreset --turbtype Mann > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
