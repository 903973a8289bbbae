mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r05b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r05b_pytest.log; tail -25 gpurun_out/r05b_pytest.log
echo "=== trace 512 split"; bash scripts/gpu_trace.sh r05b_512_split 512 2>&1 | tail -30
echo "=== trace 512 nosplit"; WG_NO_SPLIT=1 bash scripts/gpu_trace.sh r05b_512_nosplit 512 2>&1 | tail -30
