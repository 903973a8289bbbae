mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r05r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r05r_pytest.log; tail -22 gpurun_out/r05r_pytest.log
