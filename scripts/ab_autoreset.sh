#!/bin/bash
# A/B library variants under scratch_libs/ on the device-pool auto-reset loop: scripts/ab_autoreset.sh lib1.so lib2.so ... (3 runs each, interleaved)
cp windgym_b200/lib/libwindgym_b200.so /tmp/orig.so
for R in 1 2 3; do
  for L in "$@"; do
    cp scratch_libs/$L windgym_b200/lib/libwindgym_b200.so
    echo -n "$L run $R: "; python scripts/autoreset_probe.py 4096 device 900 2>&1 | grep "device pool" | cut -c1-120
  done
done
cp /tmp/orig.so windgym_b200/lib/libwindgym_b200.so
