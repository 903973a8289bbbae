"""Per-CTA timeline of one wg_step flow launch (needs the -DWG_TRACE build of the library; see scripts/gpu_trace.sh).
Writes start / end (ns, globaltimer), SM id, env and live stations of every CTA of the last launch to an .npz and
prints a summary: tail length, CTA duration against its tile count."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from windgym_b200 import V80, VecWindFarmEnv, _lib  # noqa: E402


def main():
    out = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    nx = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    ny = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    T = nx * ny
    cfg = bench.workload_config(nx, ny, "Power_avg")
    ws, ti, wd, yaw0 = bench.sample_conditions(cfg, np.arange(B), T)
    env = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", n_passthrough=bench.n_passthrough_for(200, cfg), seed=0)
    env.reset(wind=(ws, ti, wd), yaw0=yaw0)
    acts = (torch.rand((40, B, T), generator=torch.Generator().manual_seed(1)) * 2 - 1).cuda()
    lib = _lib.load()
    phb = np.zeros(8, dtype=np.uint64)
    for i in range(40):
        if i == 39:
            torch.cuda.synchronize()
            lib.wg_debug_phase_read(phb.ctypes.data_as(C.c_void_p), C.c_int(1))
        env.step(acts[i])
    torch.cuda.synchronize()
    lib.wg_debug_phase_read(phb.ctypes.data_as(C.c_void_p), C.c_int(0))
    tot = float(phb.sum())
    for name, v in zip(("outside the tile loop", "tile set-up (segments, load issue, scalars, prefetch)", "wait for the tile",
                        "move + march", "store issue + superposition", "wait for the store to release the buffer",
                        "after the tile loop (barrier, epilogue, release, write-back)"), phb):
        print(f"  warp time: {name:55s} {float(v) / tot * 100:5.1f} %")
    n = max(B * env.n_farms * 2, 2048)          # CTAs of the launch: one per work-table entry (parts of split farms)
    buf = np.zeros((n, 8), dtype=np.uint64)
    rc = lib.wg_debug_trace_read(buf.ctypes.data_as(C.c_void_p), C.c_int(n))
    assert rc == 0, rc
    buf = buf[buf[:, 1] > 0]                     # unused table entries never stamp
    n = len(buf)
    t0 = buf[:, 0].astype(np.int64); t1 = buf[:, 1].astype(np.int64)
    sm = (buf[:, 2] & np.uint64(0xffff)).astype(np.int64)
    nparts = ((buf[:, 2] >> np.uint64(16)) & np.uint64(0xff)).astype(np.int64)
    envb = (buf[:, 3] >> np.uint64(32)).astype(np.int64)
    nst = (buf[:, 3] & np.uint64(0xffffffff)).astype(np.int64)
    ph = buf[:, 4:8].astype(np.int64)
    np.savez(out, t0=t0, t1=t1, sm=sm, env=envb, stations=nst, phases=ph, nparts=nparts)
    seg = np.stack([ph[:, 0] - t0, ph[:, 1] - ph[:, 0], ph[:, 2] - ph[:, 1], ph[:, 3] - ph[:, 2], t1 - ph[:, 3]], 1) / 1e3
    for name, col in zip(("prologue (to first barrier)", "substep head (retire, prefix)", "tile loop (warp 0)",
                          "loop end -> epilogue barrier", "turbine epilogue + release + write-back"), seg.T):
        print(f"  {name:42s} mean {col.mean():6.2f} us  p10 {np.percentile(col, 10):6.2f}  p90 {np.percentile(col, 90):6.2f}")
    base = t0.min(); span = t1.max() - base
    dur = (t1 - t0) / 1e3
    tiles = np.ceil(nst / 32); rounds = np.maximum(np.ceil(tiles / (4 * np.maximum(nparts, 1))), 1)
    print(f"{n} CTAs; parts per farm: " + ", ".join(f"{k}: {int((nparts == k).sum())} CTAs" for k in np.unique(nparts)))
    print(f"kernel span {span / 1e3:.1f} us; CTA duration mean {dur.mean():.1f} us min {dur.min():.1f} max {dur.max():.1f}")
    A = np.vstack([rounds, np.ones_like(rounds)]).T
    coef, *_ = np.linalg.lstsq(A, dur, rcond=None)
    print(f"duration ~ {coef[0]:.2f} us x rounds + {coef[1]:.2f} us  (rounds mean {rounds.mean():.2f})")
    # resident CTAs over time
    ev = np.concatenate([np.stack([t0 - base, np.ones_like(t0)], 1), np.stack([t1 - base, -np.ones_like(t1)], 1)])
    ev = ev[np.argsort(ev[:, 0], kind="stable")]
    res = np.cumsum(ev[:, 1])
    tgrid = np.linspace(0, span, 21)
    occ = [res[np.searchsorted(ev[:, 0], t, side="right") - 1] if t > 0 else 0 for t in tgrid]
    print("resident CTAs at 5% marks:", [int(o) for o in occ])
    area = np.sum(np.diff(ev[:, 0]) * res[:-1])
    print(f"mean resident CTAs {area / span:.1f} of {res.max():.0f} peak -> {area / span / res.max() * 100:.1f}% slot utilisation")
    last = np.array([t1[sm == s].max() - base for s in np.unique(sm)]) / 1e3
    print(f"SM finish times: min {last.min():.1f} median {np.median(last):.1f} max {last.max():.1f} us")
    # speed of a CTA as a function of co-residency: us per round in the first half vs tail
    mid = (t0 + t1) / 2 - base
    for lo, hi in ((0, 0.5), (0.5, 0.8), (0.8, 1.0)):
        m = (mid >= lo * span) & (mid < hi * span)
        if m.any():
            print(f"CTAs centred in [{lo:.1f},{hi:.1f}] of the span: {m.sum()} CTAs, {np.mean(dur[m] / rounds[m]):.2f} us per round")


if __name__ == "__main__":
    main()
