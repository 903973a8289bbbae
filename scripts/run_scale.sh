#!/bin/bash
# Strong-scaling run of the named metric on N GPUs of one box: scripts/run_scale.sh <tag> <N>   (gpurun --gpus N)
TAG=$1; N=$2
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "exit $?"; tail -3 gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n$N.json")); r=d["roofline"]
print("N=$N value %.0f e2e %.0f ms/step %.4f flow_ms %.4f frac %.3f scaling %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["ms_per_launch"], r["frac"], d["scaling"]))
if "with_autoreset" in d: print("autoreset %.0f" % d["with_autoreset"]["value"])
for k,v in d.get("configs",{}).items(): print("   %s: value %.0f e2e %.0f frac %.3f" % (k, v["value"], v["e2e"], v["roofline_frac"]))
PY
