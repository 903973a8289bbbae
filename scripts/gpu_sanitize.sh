#!/bin/bash
# compute-sanitizer on the round-2 kernels: scripts/gpu_sanitize.sh <tag>
TAG=$1
mkdir -p gpurun_out
# memcheck: split farms + work table + PDL + pair finish (parity), device pool (adapters), bricks + Random (turbulence)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adapters.py tests/test_gpu_turbulence.py -m gpu -q -x \
  -k "step_parity_power_avg or work_table or free_running or device_pool or step_host or brick or random_white or mannload" > gpurun_out/${TAG}_compute_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; tail -4 gpurun_out/${TAG}_compute_sanitizer_memcheck.log
# racecheck (shared-memory hazards): the flow kernel with parts / fixed-point atomics and the two-warp finish kernel
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "step_parity_power_avg or work_table" > gpurun_out/${TAG}_compute_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?"; tail -4 gpurun_out/${TAG}_compute_sanitizer_racecheck.log
