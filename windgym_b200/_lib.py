"""ctypes binding of libwindgym_b200.so (C-ABI declared in include/windgym_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) into ``windgym_b200/lib``.
There is no CPU fallback: a missing library or a failing call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libwindgym_b200.so")


class WgError(RuntimeError):
    pass


class MesChannel(C.Structure):
    _fields_ = [("current", C.c_int32), ("rolling_mean", C.c_int32), ("history_N", C.c_int32),
                ("history_length", C.c_int32), ("window_length", C.c_int32)]


class MesConfig(C.Structure):
    _fields_ = [("ws", MesChannel), ("wd", MesChannel), ("yaw", MesChannel), ("power", MesChannel),
                ("turb_ws", C.c_int32), ("turb_wd", C.c_int32), ("turb_TI", C.c_int32), ("turb_power", C.c_int32),
                ("farm_ws", C.c_int32), ("farm_wd", C.c_int32), ("farm_TI", C.c_int32), ("farm_power", C.c_int32),
                ("ws_min", C.c_double), ("ws_max", C.c_double), ("wd_min", C.c_double), ("wd_max", C.c_double),
                ("yaw_min", C.c_double), ("yaw_max", C.c_double), ("ti_min", C.c_double), ("ti_max", C.c_double),
                ("power_max", C.c_double),
                ("noise", C.c_int32), ("noise_std", C.c_float * 4), ("noise_seed", C.c_uint64),
                ("multi_agent", C.c_int32)]


class Config(C.Structure):
    _fields_ = [("n_envs", C.c_int32), ("n_turb", C.c_int32), ("n_farms", C.c_int32), ("p_cap", C.c_int32),
                ("substeps", C.c_int32), ("dt", C.c_float), ("diameter", C.c_float), ("hub_height", C.c_float),
                ("d_particle", C.c_float), ("n_tab", C.c_int32),
                ("tab_ws", C.POINTER(C.c_float)), ("tab_power", C.POINTER(C.c_float)), ("tab_ct", C.POINTER(C.c_float)),
                ("x_pos", C.POINTER(C.c_double)), ("y_pos", C.POINTER(C.c_double)),
                ("action_method", C.c_int32), ("yaw_min", C.c_float), ("yaw_max", C.c_float), ("yaw_step", C.c_float),
                ("base_controller", C.c_int32), ("power_reward", C.c_int32), ("power_avg", C.c_int32),
                ("power_scaling", C.c_float), ("action_penalty", C.c_float), ("action_penalty_type", C.c_int32),
                ("steps_on_reset", C.c_int32), ("mes", MesConfig), ("act_var", C.c_int32), ("derate_min", C.c_float)]


class ResetArgs(C.Structure):
    _fields_ = [("mask", C.c_void_p), ("ws", C.c_void_p), ("ti_flow", C.c_void_p), ("wd", C.c_void_p),
                ("yaw0", C.c_void_p), ("rated_power", C.c_void_p), ("k_emit", C.c_void_p),
                ("t_developed", C.c_void_p), ("time_max", C.c_void_p), ("tb_offset", C.c_void_p),
                ("tb_scale", C.c_void_p)]


class PoolDraw(C.Structure):
    _fields_ = [("ws_min", C.c_double), ("ws_max", C.c_double), ("ti_min", C.c_double), ("ti_max", C.c_double),
                ("wd_min", C.c_double), ("wd_max", C.c_double), ("yaw_start", C.c_double), ("n_passthrough", C.c_double),
                ("tb_std_u", C.c_double), ("yaw_const", C.c_float), ("yaw_random", C.c_int32), ("eval_mode", C.c_int32),
                ("seed", C.c_uint64)]


# every symbol include/windgym_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "wg_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "wg_destroy": (None, [C.c_void_p]),
    "wg_last_error": (C.c_char_p, []),
    "wg_version": (C.c_int, []),
    "wg_state_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "wg_obs_dim": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "wg_state_field": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_size_t), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "wg_state_field_name": (C.c_char_p, [C.c_void_p, C.c_int32]),
    "wg_reset": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(ResetArgs), C.c_void_p, C.c_void_p]),
    "wg_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wg_result_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "wg_step_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                               C.c_void_p]),
    "wg_flow_steps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "wg_mes_push_extract": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "wg_set_active": (C.c_int, [C.c_void_p, C.c_int32]),
    "wg_set_slot_share": (C.c_int, [C.c_void_p, C.c_float]),
    "wg_copy_envs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "wg_pool_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "wg_pool_refill": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(PoolDraw), C.c_void_p, C.c_int32, C.c_void_p]),
    "wg_pool_swap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wg_pool_need": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "wg_pool_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]),
    "wg_set_turbulence": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                    C.c_float, C.c_float]),
    "wg_set_added_turbulence": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                          C.c_float, C.c_float, C.c_float]),
    "wg_flow_field": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                C.c_float, C.c_void_p, C.c_void_p]),
    "wg_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "wg_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "wg_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
}

_lib = None


def load():
    """Load the shared library (once) and type every entry point.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise WgError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(windgym_b200 has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(status):
    if status != 0:
        raise WgError(f"[{status}] {load().wg_last_error().decode()}")
