#!/bin/bash
# A/B prebuilt library variants under scratch_libs/ on one bench command: scripts/gpu_libab.sh <tag> "<bench args>" lib1.so lib2.so ...
TAG=$1; shift; ARGS=$1; shift
mkdir -p gpurun_out
cp windgym_b200/lib/libwindgym_b200.so /tmp/orig.so
for L in "$@"; do
  cp scratch_libs/$L windgym_b200/lib/libwindgym_b200.so; touch windgym_b200/lib/libwindgym_b200.so
  timeout 600 python bench.py --steps 60 --warmup 6 --no-cpu --no-autoreset $ARGS > gpurun_out/${TAG}_$L.json 2> gpurun_out/${TAG}_$L.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$L.json")); r=d["roofline"]
    print("$L: value %.0f flow_ms %.4f frac %.3f" % (d["value"], r["ms_per_launch"], r["frac"]))
except Exception as e:
    print("$L failed", e)
PY
  tail -1 gpurun_out/${TAG}_$L.err
done
cp /tmp/orig.so windgym_b200/lib/libwindgym_b200.so
