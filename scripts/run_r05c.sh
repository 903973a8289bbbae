mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r05c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r05c_pytest.log; tail -25 gpurun_out/r05c_pytest.log
echo "=== trace 512 split"; bash scripts/gpu_trace.sh r05c_512_split 512 2>&1 | tail -24
echo "=== trace 2048 4x2"; bash scripts/gpu_trace.sh r05c_cfg5 2048 4 2 2>&1 | tail -24
cat > /tmp/spec.txt <<'EOS'
cfg3_512 | - | --envs 512 --steps 200 --warmup 10 --no-cpu --no-autoreset --no-extras
cfg3_256 | - | --envs 256 --steps 200 --warmup 10 --no-cpu --no-autoreset --no-extras
cfg3_1024 | - | --envs 1024 --steps 200 --warmup 10 --no-cpu --no-autoreset --no-extras
cfg5_2048 | - | --envs 2048 --nx 4 --ny 2 --steps 100 --warmup 10 --no-cpu --no-autoreset --no-extras
cfg5_2048_nosplit | WG_NO_SPLIT=1 | --envs 2048 --nx 4 --ny 2 --steps 100 --warmup 10 --no-cpu --no-autoreset --no-extras
cfg5_256 | - | --envs 256 --nx 4 --ny 2 --steps 100 --warmup 10 --no-cpu --no-autoreset --no-extras
cfg2 | - | --steps 100 --warmup 10 --no-cpu --no-autoreset --no-extras
EOS
bash scripts/gpu_multi.sh r05c /tmp/spec.txt 0
