mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r05k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r05k_pytest.log; tail -5 gpurun_out/r05k_pytest.log
python scripts/e2e_probe.py 512; python scripts/e2e_probe.py 4096
WG_NO_ZEROCOPY=1 python scripts/e2e_probe.py 512
cat > /tmp/spec.txt <<'EOS'
cfg3_1024 | - | --envs 1024 --steps 200 --warmup 10 --no-cpu --no-extras --no-autoreset
cfg3_1024_no2w | WG_NO_TWOWAVE=1 | --envs 1024 --steps 200 --warmup 10 --no-cpu --no-extras --no-autoreset
cfg3_1536 | - | --envs 1536 --steps 200 --warmup 10 --no-cpu --no-extras --no-autoreset
cfg3_1536_no2w | WG_NO_TWOWAVE=1 | --envs 1536 --steps 200 --warmup 10 --no-cpu --no-extras --no-autoreset
EOS
bash scripts/gpu_multi.sh r05k /tmp/spec.txt 0
