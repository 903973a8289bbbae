mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_adapters.py -m gpu -q -x > gpurun_out/r05h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r05h_pytest.log; tail -8 gpurun_out/r05h_pytest.log
cat > /tmp/spec.txt <<'EOS'
cfg3_512 | - | --envs 512 --steps 200 --warmup 10 --no-cpu --no-extras
cfg2 | - | --steps 100 --warmup 10 --no-cpu --no-extras
EOS
bash scripts/gpu_multi.sh r05h /tmp/spec.txt 0
bash scripts/gpu_profiles.sh r05h
