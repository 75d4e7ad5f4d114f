"""ctypes binding of librlca.so (the C ABI declared in include/rlca.h).

PyTorch is plumbing here: it owns device memory and streams; the library gets
raw pointers.  There is NO fallback: if the shared library is missing or a
call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'librlca.so')


class RlcaError(RuntimeError):
    pass


class EnvConfig(C.Structure):
    """Mirror of rlca_env_config (include/rlca.h)."""
    _fields_ = [
        ('robots_per_world', C.c_int32), ('num_worlds', C.c_int32),
        ('beams', C.c_int32), ('raw_beams', C.c_int32),
        ('grid_w', C.c_int32), ('grid_h', C.c_int32),
        ('origin_cx', C.c_int32), ('origin_cy', C.c_int32),
        ('resolution', C.c_float), ('ppm', C.c_float),
        ('dt', C.c_float), ('inv_dt', C.c_float),
        ('range_max', C.c_float), ('range_cells', C.c_float),
        ('fov', C.c_float),
        ('half_len', C.c_float), ('half_wid', C.c_float),
        ('goal_radius', C.c_float), ('reward_arrive', C.c_float),
        ('reward_collision', C.c_float), ('progress_gain', C.c_float),
        ('w_threshold', C.c_float), ('w_penalty', C.c_float),
        ('v_min', C.c_float), ('v_max', C.c_float), ('w_min', C.c_float), ('w_max', C.c_float),
        ('timeout', C.c_int32), ('pre_distance_zero', C.c_int32),
        ('scenario', C.c_int32), ('auto_reset', C.c_int32),
        ('max_reject', C.c_int32), ('world_offset', C.c_int32),
        ('seed', C.c_uint64),
    ]


class EnvState(C.Structure):
    _fields_ = [('pose_dev', C.c_void_p), ('goal_dev', C.c_void_p),
                ('acc_dev', C.c_void_p), ('meta_dev', C.c_void_p)]


class StepIO(C.Structure):
    _fields_ = [('action_dev', C.c_void_p), ('live_dev', C.c_void_p), ('obs_dev', C.c_void_p),
                ('reward_dev', C.c_void_p), ('flags_dev', C.c_void_p), ('gs_dev', C.c_void_p),
                ('eplog_dev', C.c_void_p), ('stack_in_dev', C.c_void_p), ('stack_out_dev', C.c_void_p)]


# every symbol include/rlca.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    'rlca_walk_tables_host': (C.c_int, [C.c_float, _P, _P, _P, _P, _P, _P, _P]),
    'rlca_env_create': (C.c_int, [C.POINTER(EnvConfig), C.POINTER(_P)]),
    'rlca_env_destroy': (C.c_int, [_P]),
    'rlca_env_set_map': (C.c_int, [_P, _P, C.c_int32, C.c_int32]),
    'rlca_env_set_tables': (C.c_int, [_P, _P, _P]),
    'rlca_env_reset': (C.c_int, [_P, C.POINTER(EnvState), _P, C.c_int32, _P]),
    'rlca_env_observe': (C.c_int, [_P, C.POINTER(EnvState), C.POINTER(StepIO), _P]),
    'rlca_env_step': (C.c_int, [_P, C.POINTER(EnvState), C.POINTER(EnvState), C.POINTER(StepIO), _P]),
    'rlca_env_step_host': (C.c_int, [_P, C.POINTER(EnvState), C.POINTER(EnvState), C.POINTER(StepIO),
                                     _P, _P, _P, _P, _P, _P]),
    'rlca_raycast': (C.c_int, [_P, _P, _P, C.c_int32, _P]),
    'rlca_env_set_ctas_per_world': (C.c_int, [_P, C.c_int32]),
    'rlca_env_set_host_chunks': (C.c_int, [_P, C.c_int32]),
    'rlca_env_set_host_zero_copy': (C.c_int, [_P, C.c_int32]),
    'rlca_env_launch_count': (C.c_int64, [_P]),
    'rlca_sizeof_env_config': (C.c_int, []),
    'rlca_policy_param_offset': (C.c_int64, [C.c_int32]),
    'rlca_policy_param_size': (C.c_int64, [C.c_int32]),
    'rlca_policy_launch_count': (C.c_int64, [_P]),
    'rlca_policy_set_grad_event': (C.c_int, [_P, _P]),
    'rlca_adam_step_allreduce': (C.c_int, [_P, _P, _P, _P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32,
                                           C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_float, C.c_int32, _P]),
    'rlca_policy_set_tensor_cores': (C.c_int, [_P, C.c_int32]),
    'rlca_policy_weights_changed': (C.c_int, [_P]),
    'rlca_policy_features': (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P]),
    'rlca_policy_create': (C.c_int, [C.c_int32, C.POINTER(_P)]),
    'rlca_policy_destroy': (C.c_int, [_P]),
    'rlca_policy_forward': (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, _P, _P]),
    'rlca_policy_sample': (C.c_int, [_P, _P, C.c_int32, C.c_uint64, C.c_uint64, C.c_int32, _P, _P, _P, _P]),
    'rlca_ppo_loss_fwd_bwd': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_float, C.c_float, C.c_float,
                                        _P, _P]),
    'rlca_ppo_loss_fwd_bwd_weighted': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_float, C.c_float,
                                                 C.c_float, C.c_float, _P, _P]),
    'rlca_policy_backward': (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, _P]),
    'rlca_adam_step': (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32,
                                 C.c_float, _P]),
    'rlca_policy_adam_step': (C.c_int, [_P, _P, _P, _P, _P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32,
                                        C.c_float, _P]),
    'rlca_gae': (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_float, C.c_float, _P, _P, _P]),
    'rlca_adv_moments': (C.c_int, [_P, C.c_int64, _P, _P]),
    'rlca_adv_apply': (C.c_int, [_P, C.c_int64, _P, _P, _P]),
    'rlca_gather_rows': (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P]),
    'rlca_gather_minibatch': (C.c_int, [_P, _P, C.c_int32, _P, C.c_int32, _P, _P]),
    'rlca_obs_stack_push': (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, _P]),
    'rlca_last_error': (C.c_char_p, []),
    'rlca_version': (C.c_char_p, []),
}

_lib = None


def load():
    """Load librlca.so (built in-tree by __graft_entry__.build()). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RlcaError(f'{LIB_PATH} not found: run `python -c "import __graft_entry__ as g; g.build()"` '
                        '(there is no CPU fallback)')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().rlca_last_error()
        raise RlcaError(f'librlca error {rc}: {msg.decode() if msg else "?"}')
