#!/usr/bin/env python
"""Group ncu per-line counters of wg_flow_kernel by code region (function / phase). usage: ncu_groups.py <rep> [sym]"""
import subprocess, sys, os
here = os.path.dirname(os.path.abspath(__file__))
rep = sys.argv[1]; sym = sys.argv[2] if len(sys.argv) > 2 else "wg_flow_kernel"
out = subprocess.run([sys.executable, os.path.join(here, "ncu_by_line.py"), rep, sym, "0.0"], capture_output=True, text=True).stdout
src = open(os.path.join(here, "..", "windgym_b200", "csrc", "flow.cu")).read().splitlines()
def find(pat):
    return next(i + 1 for i, l in enumerate(src) if pat in l)
marks = sorted([("ptx helpers", 1), ("tmem helpers", find("TMEM scratch\n") if False else find("__device__ __forceinline__ void tmem_alloc")),
         ("physics helpers (filters/interp/moved)", find("__device__ __forceinline__ float f1_filter")),
         ("march", find("#define WG_NODE_FWD")), ("count_below/locate/segments", find("__device__ __forceinline__ int count_below")),
         ("flush_hits", find("struct RotorAcc")), ("prologue", find("__global__ void __launch_bounds__")),
         ("substep head (controller/retire/prefix)", find("for (int sub = 0; sub < nsteps")),
         ("tile pipeline (load/prefetch/store)", find("warp-private tile pipeline")),
         ("bracket detection", find("// ---- superposition")), ("round end + turbine epilogue", find("const bool more = sub + 1 < nsteps;")),
         ("particle release", find("// release one particle per turbine")), ("tail", find("d.yaw[bf * T + tid] = sh.yaw[tid];"))], key=lambda m: m[1])
agg = {}
for l in out.splitlines()[1:]:
    parts = l.split()
    if len(parts) < 6: continue
    try:
        f = parts[0].rstrip(":"); ln = int(parts[1]); i = float(parts[3].rstrip("%")); s = float(parts[5].rstrip("%"))
    except ValueError:
        continue
    g = "cuda headers (shuffles etc.)" if f != "flow.cu" else [m[0] for m in marks if m[1] <= ln][-1]
    a = agg.setdefault(g, [0.0, 0.0]); a[0] += i; a[1] += s
print(out.splitlines()[0])
for g, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-42s inst %6.2f%%  stall samples %6.2f%%" % (g, i, s))
