mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r05l_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r05l_pytest.log; tail -5 gpurun_out/r05l_pytest.log
python scripts/e2e_probe.py 512; python scripts/e2e_probe.py 4096
cat > /tmp/spec.txt <<'EOS'
cfg3_512 | - | --envs 512 --steps 200 --warmup 10 --no-cpu --no-extras --no-autoreset
cfg2 | - | --steps 100 --warmup 10 --no-cpu --no-extras --no-autoreset
EOS
bash scripts/gpu_multi.sh r05l /tmp/spec.txt 0
