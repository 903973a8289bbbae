mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r05d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r05d_pytest.log; tail -15 gpurun_out/r05d_pytest.log
cat > /tmp/spec.txt <<'EOS'
cfg3_512 | - | --envs 512 --steps 200 --warmup 10 --no-cpu --no-autoreset --no-extras
cfg3_512_nopdl | WG_NO_PDL=1 | --envs 512 --steps 200 --warmup 10 --no-cpu --no-autoreset --no-extras
cfg3_256 | - | --envs 256 --steps 200 --warmup 10 --no-cpu --no-autoreset --no-extras
mann_ref | - | --turbtype Mann --steps 40 --warmup 5 --no-cpu --no-autoreset --no-extras
mann_ref_nobricks | WG_NO_BRICKS=1 | --turbtype Mann --steps 40 --warmup 5 --no-cpu --no-autoreset --no-extras
mann_test | - | --turbtype Mann --mann-box test --steps 40 --warmup 5 --no-cpu --no-autoreset --no-extras
mann_test_bricks | WG_FORCE_BRICKS=1 | --turbtype Mann --mann-box test --steps 40 --warmup 5 --no-cpu --no-autoreset --no-extras
cfg2 | - | --steps 100 --warmup 10 --no-cpu --no-autoreset --no-extras
EOS
bash scripts/gpu_multi.sh r05d /tmp/spec.txt 0
