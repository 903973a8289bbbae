mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r05o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r05o_pytest.log; tail -5 gpurun_out/r05o_pytest.log
python scripts/e2e_probe.py 512; python scripts/e2e_probe.py 4096
cat > /tmp/spec.txt <<'EOS'
mann | - | --turbtype Mann --steps 40 --warmup 5 --no-cpu --no-extras --no-autoreset
mann_test | - | --turbtype Mann --mann-box test --steps 40 --warmup 5 --no-cpu --no-extras --no-autoreset
cfg3_512 | - | --envs 512 --steps 200 --warmup 10 --no-cpu --no-extras --no-autoreset
cfg2 | - | --steps 100 --warmup 10 --no-cpu --no-extras --no-autoreset
EOS
bash scripts/gpu_multi.sh r05o /tmp/spec.txt 0
