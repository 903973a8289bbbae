#!/bin/bash
# ncu --set full capture of one launch of a named kernel inside a short bench run:
#   scripts/gpu_ncu_kernel.sh <tag> <kernel-regex> [skip]
TAG=$1; KRN=$2; SKIP=${3:-30}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRN -s $SKIP -c 1 -f -o gpurun_out/${TAG} python bench.py --steps 8 --warmup 3 --no-cpu --no-autoreset > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?"
