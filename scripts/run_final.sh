mkdir -p gpurun_out
TAG=${1:-r06}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${TAG}_smoke.log
( time python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -4 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json")); r=d["roofline"]
print("value %.0f e2e %.0f (%.0f%%) ms/step %.4f flow_ms %.4f frac %.3f fin_ms %.4f launches %d scaling %s" % (d["value"], d["e2e"]["value"], 100*d["e2e"]["value"]/d["value"], d["ms_per_step"], r["ms_per_launch"], r["frac"], r["finish_kernel_ms"], d["gpu_launches"], d["scaling"]))
print("autoreset %.0f host pool %.0f" % (d["with_autoreset"]["value"], d["with_autoreset_host_pool"]["value"]), d["with_autoreset"]["pool"])
for k,v in d.get("configs",{}).items(): print("   %s: value %.0f e2e %.0f frac %.3f flow_ms %.4f fin_ms %.4f" % (k, v["value"], v["e2e"], v["roofline_frac"], v["ms_per_launch"], v["finish_kernel_ms"]))
print(d["cpu_baseline"]); print(d["clocks"])
PY
( time python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/${TAG}_reference_arm.json 2> gpurun_out/${TAG}_reference_arm.err; echo "ref exit $?"; cat gpurun_out/${TAG}_reference_arm.json | cut -c1-400; tail -4 gpurun_out/${TAG}_reference_arm.err
