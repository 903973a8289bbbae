#!/bin/bash
# Profile set for profiles/: scripts/gpu_profiles.sh <tag>   (run through gpurun, ONE GPU)
TAG=$1
mkdir -p gpurun_out
B="--steps 8 --warmup 3 --no-cpu --no-autoreset --no-extras"
# 1. launch list of the default workload (per-launch durations are cold-cache / serialised under ncu)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $B > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "launch list exit $?"
# 2. full captures of the flow kernel: default workload, 512 envs (split farms), Mann reference box
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wg_flow_kernel -s 30 -c 1 -f -o gpurun_out/${TAG}_flow python bench.py $B > gpurun_out/${TAG}_ncu_flow.log 2>&1
echo "flow exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wg_flow_kernel -s 30 -c 1 -f -o gpurun_out/${TAG}_flow512 python bench.py $B --envs 512 > gpurun_out/${TAG}_ncu_flow512.log 2>&1
echo "flow512 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wg_flow_kernel -s 40 -c 1 -f -o gpurun_out/${TAG}_flowmann python bench.py $B --turbtype Mann > gpurun_out/${TAG}_ncu_flowmann.log 2>&1
echo "flowmann exit $?"
# 3. the finish kernel at 512 envs
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wg_finish_kernel -s 40 -c 1 -f -o gpurun_out/${TAG}_finish512 python bench.py $B --envs 512 > gpurun_out/${TAG}_ncu_finish512.log 2>&1
echo "finish512 exit $?"
ls -la gpurun_out/${TAG}_*
