"""CPU: the C-ABI shared library builds, loads and exports every symbol include/windgym_b200.h declares; argument
validation works without a GPU; the product path fails loudly (no CPU fallback) when no B200 is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "windgym_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wg_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_exported(built_lib):
    names = _declared_functions()
    assert len(names) >= 13
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/windgym_b200.h but not exported"


def test_ctypes_table_matches_header(built_lib):
    from windgym_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _declared_functions()


def test_version(built_lib):
    assert built_lib.wg_version() == 100


def _config(**over):
    from windgym_b200 import _lib, V80
    t = V80()
    tabs = [np.ascontiguousarray(a, dtype=np.float32) for a in (t.ws_table, t.power_table_w, t.ct_table)]
    xy = [np.array([0.0, 640.0]), np.array([0.0, 0.0])]
    ch = _lib.MesChannel(0, 1, 1, 10, 10)
    mes = _lib.MesConfig(ws=ch, wd=ch, yaw=ch, power=ch, turb_ws=1, ws_min=2.0, ws_max=25.0, wd_min=250.0,
                         wd_max=290.0, yaw_min=-45.0, yaw_max=45.0, ti_min=0.0, ti_max=0.5, power_max=2e6)
    fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
    kw = dict(n_envs=2, n_turb=2, n_farms=1, p_cap=64, substeps=1, dt=1.0, diameter=80.0, hub_height=70.0,
              d_particle=0.2, n_tab=len(tabs[0]), tab_ws=tabs[0].ctypes.data_as(fp),
              tab_power=tabs[1].ctypes.data_as(fp), tab_ct=tabs[2].ctypes.data_as(fp),
              x_pos=xy[0].ctypes.data_as(dp), y_pos=xy[1].ctypes.data_as(dp), action_method=0, yaw_min=-45.0,
              yaw_max=45.0, yaw_step=1.0, base_controller=0, power_reward=2, power_avg=10, power_scaling=1.0,
              action_penalty=0.0, action_penalty_type=0, steps_on_reset=10, mes=mes)
    kw.update(over)
    cfg = _lib.Config(**kw)
    cfg._keep = (tabs, xy)
    return cfg


@pytest.mark.parametrize("over,code,msg", [
    (dict(n_envs=0), -1, "n_envs"),
    (dict(n_turb=65), -1, "n_turb"),
    (dict(p_cap=10), -1, "p_cap"),
    (dict(action_method=2), -2, "absolute method is not implemented"),          # Wind_Farm_Env.py:861
    (dict(action_method=7), -1, "ActionMethod must be yaw, wind or absolute"),    # :864
    (dict(power_reward=9), -1, "Power_reward must be either"),                    # :192
    (dict(power_reward=3, power_avg=39), -1, "larger then 40"),                   # :186-190
    (dict(power_reward=1, n_farms=1), -1, "Baseline reward needs"),
    (dict(n_farms=2, base_controller=5), -1, "BaseController must be either"),    # :314
    (dict(substeps=0), -1, "dt_env must be a multiple"),                          # :107
    (dict(steps_on_reset=0), -1, "fill_window"),                                  # :240
])
def test_create_rejects_bad_config(built_lib, over, code, msg):
    h = C.c_void_p()
    cfg = _config(**over)
    rc = built_lib.wg_create(C.byref(cfg), C.byref(h))
    assert rc == code
    assert msg in built_lib.wg_last_error().decode()
    assert not h.value


def test_null_arguments(built_lib):
    assert built_lib.wg_create(None, None) == -1
    n = C.c_size_t()
    assert built_lib.wg_state_bytes(None, C.byref(n)) == -1
    assert built_lib.wg_step(None, None, None, None, None, None, None) == -1
    assert built_lib.wg_profile_enable(None, 1) == -1
    assert built_lib.wg_result_bytes(None, C.byref(n)) == -1
    assert built_lib.wg_step_host(None, None, None, None, None, None, 0, None) == -1
    assert "null argument" in built_lib.wg_last_error().decode()


def test_no_cpu_fallback(built_lib):
    """Without a B200 the library refuses to create a handle; the Python layer raises (never routes to the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    h = C.c_void_p()
    cfg = _config()
    assert built_lib.wg_create(C.byref(cfg), C.byref(h)) == -4
    assert "no CPU fallback" in built_lib.wg_last_error().decode()
    from windgym_b200 import V80, VecWindFarmEnv, WgError
    from tests.helpers import small_config
    with pytest.raises(WgError):
        VecWindFarmEnv(V80(), 2, config=small_config(2, 1), device="cpu")
    with pytest.raises(Exception):
        VecWindFarmEnv(V80(), 2, config=small_config(2, 1), device="cuda:0")


def test_product_never_imports_oracle():
    """The product package must not reference the oracle (test infrastructure) anywhere."""
    pkg = os.path.join(ROOT, "windgym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f"{f} imports the oracle"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from windgym_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.WgError, match="no CPU fallback"):
        _lib.load()
