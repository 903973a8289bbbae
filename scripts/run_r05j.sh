mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r05j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r05j_pytest.log; tail -8 gpurun_out/r05j_pytest.log
cat > /tmp/spec.txt <<'EOS'
cfg3_512 | - | --envs 512 --steps 200 --warmup 10 --no-cpu --no-extras --no-autoreset
cfg3_1024 | - | --envs 1024 --steps 200 --warmup 10 --no-cpu --no-extras --no-autoreset
cfg3_2048 | - | --envs 2048 --steps 200 --warmup 10 --no-cpu --no-extras --no-autoreset
cfg5_256 | - | --envs 256 --nx 4 --ny 2 --steps 200 --warmup 10 --no-cpu --no-extras --no-autoreset
cfg2 | - | --steps 100 --warmup 10 --no-cpu --no-extras --no-autoreset
EOS
bash scripts/gpu_multi.sh r05j /tmp/spec.txt 0
