#!/bin/bash
# Per-CTA timeline of one wg_step flow launch (diagnostic build -DWG_TRACE in scratch_libs/trace.so, built with
# scripts/build_variant.sh trace -DWG_TRACE): scripts/gpu_trace.sh <tag> [bench-style args for trace_run.py]
TAG=$1; shift
mkdir -p gpurun_out
cp windgym_b200/lib/libwindgym_b200.so /tmp/orig.so
cp scratch_libs/trace.so windgym_b200/lib/libwindgym_b200.so
timeout 600 python scripts/trace_run.py gpurun_out/${TAG}_trace.npz "$@"
cp /tmp/orig.so windgym_b200/lib/libwindgym_b200.so
