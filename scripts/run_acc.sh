cp windgym_b200/lib/libwindgym_b200.so /tmp/orig.so
for L in base acc_rcp acc_misc acc_both; do
  if [ $L != base ]; then cp scratch_libs/$L.so windgym_b200/lib/libwindgym_b200.so; touch windgym_b200/lib/libwindgym_b200.so; fi
  echo "== $L"; python scripts/acc_probe.py 2>&1 | tail -2
  python bench.py --steps 40 --warmup 5 --no-cpu --no-extras --no-autoreset 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   cfg2 flow_ms %.4f frac %.3f'%(d['roofline']['ms_per_launch'], d['roofline']['frac']))"
  cp /tmp/orig.so windgym_b200/lib/libwindgym_b200.so
done
