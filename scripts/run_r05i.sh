mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r05i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r05i_pytest.log; tail -8 gpurun_out/r05i_pytest.log
cat > /tmp/spec.txt <<'EOS'
cfg3_512 | - | --envs 512 --steps 200 --warmup 10 --no-cpu --no-extras
full | - | --steps 100 --warmup 10
EOS
bash scripts/gpu_multi.sh r05i /tmp/spec.txt 0
