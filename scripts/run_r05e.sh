mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r05e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r05e_pytest.log; tail -30 gpurun_out/r05e_pytest.log
cat > /tmp/spec.txt <<'EOS'
cfg2_auto | - | --steps 100 --warmup 10 --no-cpu --no-extras
cfg3_512_auto | - | --envs 512 --steps 200 --warmup 10 --no-cpu --no-extras
EOS
bash scripts/gpu_multi.sh r05e /tmp/spec.txt 0
