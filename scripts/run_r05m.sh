mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r05m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r05m_pytest.log; tail -5 gpurun_out/r05m_pytest.log
cat > /tmp/spec.txt <<'EOS'
cfg2 | - | --steps 100 --warmup 10 --no-cpu --no-extras --no-autoreset
cfg3_512 | - | --envs 512 --steps 200 --warmup 10 --no-cpu --no-extras --no-autoreset
cfg4 | - | --envs 1024 --nx 8 --ny 8 --steps 40 --warmup 5 --no-cpu --no-extras --no-autoreset
cfg4_nomax | WG_MAX_PART_TILES=0 | --envs 1024 --nx 8 --ny 8 --steps 40 --warmup 5 --no-cpu --no-extras --no-autoreset
cfg4_max24 | WG_MAX_PART_TILES=24 | --envs 1024 --nx 8 --ny 8 --steps 40 --warmup 5 --no-cpu --no-extras --no-autoreset
cfg4_nomax_no2w | WG_MAX_PART_TILES=0,WG_NO_TWOWAVE=1 | --envs 1024 --nx 8 --ny 8 --steps 40 --warmup 5 --no-cpu --no-extras --no-autoreset
cfg5 | - | --envs 2048 --nx 4 --ny 2 --steps 100 --warmup 10 --no-cpu --no-extras --no-autoreset
mann | - | --turbtype Mann --steps 40 --warmup 5 --no-cpu --no-extras --no-autoreset
EOS
bash scripts/gpu_multi.sh r05m /tmp/spec.txt 0
