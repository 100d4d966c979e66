"""ctypes binding of libditree.so (include/ditree.h).

There is NO fallback: if the shared library is missing or a CUDA device is absent the import of
the symbols / creation of a context raises.  ``load()`` only dlopens (works on a CPU-only box, used
by the symbol-export test); creating a :class:`Context` needs a B200.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# DITREE_LIB selects another build of the same library (kernel A/B tuning); there is still no fallback
LIB_PATH = os.environ.get("DITREE_LIB") or os.path.join(HERE, "libditree.so")

c_i64 = C.c_int64
c_p = C.c_void_p


class TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", c_p), ("ndim", C.c_int), ("shape", c_i64 * 4)]


class ModelCfg(C.Structure):
    _fields_ = [("action_dim", C.c_int), ("horizon", C.c_int), ("cond_dim", C.c_int), ("emb_dim", C.c_int),
                ("map_size", C.c_int), ("down_dims", C.c_int * 3), ("max_batch", C.c_int)]


class PlanCfg(C.Structure):
    _fields_ = [("unit_slots", C.c_int32), ("edge_slots", C.c_int32), ("node_cap", C.c_int32), ("action_horizon", C.c_int32),
                ("n_sched", C.c_int32), ("sched_chunks", C.c_int32 * 8), ("iteration_cap", C.c_int32),
                ("ode_steps", C.c_int32), ("max_units", C.c_int32), ("max_path", C.c_int32),
                ("goal_sample_rate", C.c_float), ("goal_conditioning_bias", C.c_float),
                ("local_map_scale", C.c_double), ("norm", C.c_double * 16), ("run_type", C.c_int32), ("reserved", C.c_int32)]


class PlanUnit(C.Structure):
    _fields_ = [("start", C.c_float * 6), ("goal", C.c_float * 2), ("half_w", C.c_float), ("half_h", C.c_float),
                ("map_slot", C.c_int32), ("seed", C.c_uint32), ("unit_id", C.c_int32), ("cdf_slot", C.c_int32)]


class PlanResult(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("unit_id", "goal_reached", "has_path", "n_states", "n_actions", "n_nodes",
                                         "iterations", "first_pass", "last_pass", "collisions", "chunks", "error")]


# name -> (restype, argtypes); every symbol include/ditree.h declares
SIGNATURES = {
    "dt_ctx_create": (C.c_int, [C.c_int, C.POINTER(c_p)]),
    "dt_ctx_destroy": (None, [c_p]),
    "dt_last_error": (C.c_char_p, [c_p]),
    "dt_version": (C.c_char_p, []),
    "dt_set_option": (C.c_int, [c_p, C.c_char_p, C.c_int]),
    "dt_set_map": (C.c_int, [c_p, c_p, C.c_int, C.c_int, C.c_float, c_p]),
    "dt_set_map_slot": (C.c_int, [c_p, C.c_int, c_p, C.c_int, C.c_int, C.c_float, c_p]),
    "dt_local_map_slots": (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, C.c_int, C.c_double, c_p, C.c_int, c_p, c_p]),
    "dt_plan_create": (C.c_int, [c_p, C.POINTER(PlanCfg), C.POINTER(c_p)]),
    "dt_plan_destroy": (None, [c_p]),
    "dt_plan_push": (C.c_int, [c_p, C.POINTER(PlanUnit), C.c_int, c_p]),
    "dt_plan_pass": (C.c_int, [c_p, c_p]),
    "dt_plan_set_cdf": (C.c_int, [c_p, C.c_int, c_p, C.c_int, c_p]),
    "dt_plan_peek_slots": (C.c_int, [c_p, c_p, c_p, c_p]),
    "dt_plan_counters": (C.c_int, [c_p, c_i64, C.c_int, C.POINTER(C.c_int32)]),
    "dt_plan_fetch": (C.c_int, [c_p, C.c_int, C.POINTER(PlanResult), c_p, c_p, C.c_int, c_p]),
    "dt_plan_peek_tree": (C.c_int, [c_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), c_p, c_p, C.c_int, c_p]),
    "dt_collide_car": (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_p, c_p]),
    "dt_collide_points": (C.c_int, [c_p, c_p, c_p, c_i64, c_i64, C.c_double, C.c_double, c_p, c_p]),
    "dt_collide_ant": (C.c_int, [c_p, c_p, c_i64, c_i64, C.c_double, c_p, c_p]),
    "dt_local_map": (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, C.c_int, C.c_double, C.c_int, c_p, c_p]),
    "dt_edt_prior": (C.c_int, [c_p, c_p, c_p]),
    "dt_prob_map": (C.c_int, [c_p, C.c_int, C.c_int, c_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                              c_p, c_p, c_p]),
    "dt_log_blend": (C.c_int, [c_p, c_p, c_p, c_p, C.c_int, C.c_double, C.c_double, c_p, c_p]),
    "dt_sample_cells": (C.c_int, [c_p, c_p, C.c_int, c_p, c_i64, c_p, c_p]),
    "dt_ray_probe": (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_p, c_p]),
    "dt_path_first_obstacle": (C.c_int, [c_p, c_p, c_p, c_i64, c_i64, c_p, c_p]),
    "dt_path_first_obstacle_grid": (C.c_int, [c_p, c_p, C.c_int, C.c_int, c_p, c_p, c_i64, c_i64, c_p, c_p]),
    "dt_lidar_scan": (C.c_int, [c_p, c_p, c_i64, c_p, c_p, c_p, c_p]),
    "dt_propagate_collide": (C.c_int, [c_p, c_p, c_i64, c_i64, c_p, c_i64, c_i64, c_i64, c_i64, C.c_int, C.c_float,
                                       C.c_float, c_p, c_i64, c_i64, c_i64, c_p, c_p, c_p, C.c_int, c_p]),
    "dt_build_cond_car": (C.c_int, [c_p, c_p, c_i64, c_i64, c_p, c_p, C.c_int, c_i64, c_p, C.c_double, c_p, c_p]),
    "dt_build_cond_ant": (C.c_int, [c_p, c_p, C.c_int, C.c_int, c_p, c_p, C.c_int, c_i64, c_p, C.c_double, c_p, c_p]),
    "dt_nearest": (C.c_int, [c_p, c_p, c_p, c_i64, c_p, c_p, c_i64, c_i64, c_p, c_p]),
    "dt_nearest_k": (C.c_int, [c_p, c_p, c_p, c_i64, c_p, c_p, c_i64, c_i64, C.c_int, c_p, c_p]),
    "dt_goal_cost_argmin": (C.c_int, [c_p, c_p, c_p, c_i64, C.c_float, C.c_float, c_p, c_p, c_p]),
    "dt_mppi_reduce": (C.c_int, [c_p, c_p, c_p, c_i64, C.c_int, C.c_float, c_p, c_p, c_p, c_p]),
    "dt_mppi_rollout_cost": (C.c_int, [c_p, c_p, c_p, c_p, c_i64, C.c_int, c_p, C.c_int, C.c_int, C.c_float, C.c_float,
                                       C.c_float, C.c_float, c_p, c_p, c_p]),
    "dt_mppi_shift": (C.c_int, [c_p, c_p, C.c_int, C.c_int, c_p, c_p]),
    "dt_load_denoiser": (C.c_int, [c_p, C.POINTER(TensorDesc), C.c_int, C.POINTER(ModelCfg), c_p]),
    "dt_fm_sample": (C.c_int, [c_p, c_p, c_p, c_p, c_i64, C.c_int, C.c_double, c_p, c_p, c_p]),
    "dt_encode_map": (C.c_int, [c_p, c_p, c_i64, c_p, c_p]),
    "dt_unet_forward": (C.c_int, [c_p, c_p, c_p, c_p, c_i64, C.c_float, c_p, c_p]),
    "dt_gemm_bf16": (C.c_int, [c_p, c_p, c_p, c_i64, C.c_int, C.c_int, c_p, c_p]),
    "dt_conv2d_gn_bf16": (C.c_int, [c_p, c_p, c_i64, C.c_int, C.c_int, C.c_int, c_p, C.c_int, C.c_int, C.c_int, C.c_int, c_p, c_p,
                                    c_p, C.c_int, c_p, c_p]),
    "dt_profile_begin": (C.c_int, [c_p]),
    "dt_profile_end": (C.c_int, [c_p, C.POINTER(C.c_double), C.POINTER(c_i64)]),
    "dt_profile_csv": (C.c_int, [c_p, C.c_char_p]),
    "dt_launch_count": (c_i64, [c_p]),
    "dt_sync_status": (C.c_int, [c_p, c_p]),
}

DT_OK, DT_E_CUDA, DT_E_ARG, DT_E_NOMAP, DT_E_NOMODEL, DT_E_INDEX, DT_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
DT_PROP_STOP_ON_COLLISION = 1
DT_F32, DT_BF16 = 0, 1

_lib = None


def load():
    """dlopen libditree.so and declare every prototype.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m ditreeonlineplanner_b200.build` "
            "(there is no CPU or PyTorch fallback for the DiTree hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class DitreeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libditree error {code}: {msg}")
        self.code = code
