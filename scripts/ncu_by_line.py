#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line.

usage: scripts/ncu_by_line.py <report.ncu-rep> <kernel-symbol-substring> [min_pct]
Needs the library built with -lineinfo (nvdisasm -g gives address -> file:line; inlined frames are attributed to
the innermost line).  Prints executed warp instructions and stall samples per source line.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "windgym_b200", "lib", "libwindgym_b200.so")


def line_map(symbol_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, cur_line, active = None, None, False
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur_fn = m.group(1)
                active = symbol_sub in cur_fn
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out[int(m.group(1), 16)] = (cur_line, m.group(2))
    return out


def main():
    rep, sym = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
    lm = line_map(sym)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    hdr = rows[hi]
    ix = {h: k for k, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {}
    base = None
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            if r and r[0] == "Kernel Name":
                break
            continue
        addr = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]])
        if base is None:
            base = addr
        key, _ = lm.get(addr - base, (None, None))
        a = agg.setdefault(key, {"inst": 0.0, "samp": 0.0, "stalls": {}})
        a["inst"] += float(r[ix["Instructions Executed"]] or 0)
        a["samp"] += float(r[ix["# Samples"]] or 0)
        for sc in stall_cols:
            v = float(r[ix[sc]] or 0)
            if v:
                a["stalls"][sc] = a["stalls"].get(sc, 0.0) + v
    ti = sum(a["inst"] for a in agg.values())
    ts = sum(a["samp"] for a in agg.values())
    print(f"total warp instructions {ti:.0f}, samples {ts:.0f}")
    src = {}
    for key in sorted(k for k in agg if k):
        a = agg[key]
        if 100 * a["inst"] / ti < min_pct and 100 * a["samp"] / ts < min_pct:
            continue
        if key[0] not in src:
            p = os.path.join(ROOT, "windgym_b200", "csrc", key[0])
            src[key[0]] = open(p).read().splitlines() if os.path.isfile(p) else []
        text = src[key[0]][key[1] - 1].strip() if key[1] - 1 < len(src[key[0]]) else ""
        top = sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3]
        tops = " ".join(f"{k[6:]}:{100 * v / max(a['samp'], 1):.0f}%" for k, v in top)
        print(f"{key[0]}:{key[1]:4d} inst {100 * a['inst'] / ti:5.2f}% samp {100 * a['samp'] / ts:5.2f}%  [{tops}]  {text[:90]}")


if __name__ == "__main__":
    main()
