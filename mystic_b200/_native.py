"""ctypes binding of include/mystic_b200.h.  No fallback: importing this module without the
built CUDA library raises, and every non-zero return code becomes a Python exception."""
import ctypes
import os

from .build import LIBPATH

# a tuning variant built with `python -m mystic_b200.build --out PATH -D...` (still this CUDA library)
LIBPATH = os.environ.get('MB2_LIBRARY') or LIBPATH

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_int64_p = ctypes.POINTER(ctypes.c_int64)

EXPORTS = [
    'mb2_abi_version', 'mb2_last_error', 'mb2_create', 'mb2_destroy', 'mb2_bind_buffers', 'mb2_current',
    'mb2_set_strategy', 'mb2_set_bounds', 'mb2_set_cost', 'mb2_set_penalty', 'mb2_set_rng_philox',
    'mb2_set_rng_replay', 'mb2_init_population', 'mb2_upload_population', 'mb2_set_capture',
    'mb2_eval_initial', 'mb2_step', 'mb2_pack_best', 'mb2_merge_best', 'mb2_read_result', 'mb2_eval_cost',
    'mb2_launch_info', 'mb2_launch_count', 'mb2_kernel_family', 'mb2_selftest_division', 'mb2_selftest_cosine', 'mb2_restore_state', 'mb2_create_ws', 'mb2_workspace_bytes',
    'mb2_set_termination', 'mb2_run', 'mb2_run_result', 'mb2_set_islands', 'mb2_read_islands', 'mb2_exchange_open', 'mb2_exchange_connect', 'mb2_exchange_connected', 'mb2_exchange_reset', 'mb2_exchange_best',
    'mb2_probe_fp64_peak', 'mb2_set_penalty_ex', 'mb2_set_island_groups', 'mb2_run_result_islands',
    'mb2_set_commit_mode', 'mb2_last_commit_in_place', 'mb2_probe_pool_geometry',
]


class Result(ctypes.Structure):
    """struct mb2_result"""
    _fields_ = [('best_energy', ctypes.c_double), ('best_index', ctypes.c_int64), ('n_inf', ctypes.c_int64),
                ('n_accepted', ctypes.c_int64), ('generation', ctypes.c_int64)]


class Termination(ctypes.Structure):
    """struct mb2_termination"""
    _fields_ = [('kind', ctypes.c_int32), ('lookback', ctypes.c_int32), ('tolerance', ctypes.c_double),
                ('target', ctypes.c_double), ('gtol', ctypes.c_double), ('term_generations', ctypes.c_int64),
                ('term_evaluations', ctypes.c_int64), ('max_generations', ctypes.c_int64),
                ('max_evaluations', ctypes.c_int64), ('generations', ctypes.c_int64), ('evaluations', ctypes.c_int64)]


class NativeError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIBPATH):
        raise ImportError(
            'mystic_b200: CUDA library %s is missing. Build it with `python -m mystic_b200.build` '
            '(nvcc, sm_100a). There is no CPU fallback.' % LIBPATH)
    lib = ctypes.CDLL(LIBPATH)
    vp, i32, i64, u64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double
    sig = {
        'mb2_abi_version': (ctypes.c_int, []),
        'mb2_last_error': (ctypes.c_char_p, []),
        'mb2_create': (ctypes.c_int, [ctypes.c_int, i64, i32, i64, i64, ctypes.POINTER(vp)]),
        'mb2_destroy': (ctypes.c_int, [vp]),
        'mb2_workspace_bytes': (i64, [i32, c_int64_p]),
        'mb2_create_ws': (ctypes.c_int, [ctypes.c_int, i64, i32, i64, i64, vp, i64, vp, i64, ctypes.POINTER(vp)]),
        'mb2_bind_buffers': (ctypes.c_int, [vp, vp, vp, vp, vp]),
        'mb2_current': (ctypes.c_int, [vp, ctypes.POINTER(vp), ctypes.POINTER(vp)]),
        'mb2_set_strategy': (ctypes.c_int, [vp, ctypes.c_int, dbl, dbl]),
        'mb2_set_bounds': (ctypes.c_int, [vp, ctypes.c_int, c_double_p, c_double_p]),
        'mb2_set_cost': (ctypes.c_int, [vp, ctypes.c_int, c_double_p, i32]),
        'mb2_set_penalty': (ctypes.c_int, [vp, i32, c_int32_p, c_double_p, c_double_p, c_double_p]),
        'mb2_set_penalty_ex': (ctypes.c_int, [vp, i32, c_int32_p, c_double_p, c_double_p, c_double_p, c_double_p]),
        'mb2_set_rng_philox': (ctypes.c_int, [vp, u64]),
        'mb2_set_rng_replay': (ctypes.c_int, [vp, vp, vp, vp]),
        'mb2_init_population': (ctypes.c_int, [vp, u64, c_double_p, c_double_p, vp]),
        'mb2_upload_population': (ctypes.c_int, [vp, vp, vp]),
        'mb2_set_capture': (ctypes.c_int, [vp, vp, vp]),
        'mb2_eval_initial': (ctypes.c_int, [vp, vp]),
        'mb2_step': (ctypes.c_int, [vp, vp]),
        'mb2_pack_best': (ctypes.c_int, [vp, vp, vp]),
        'mb2_merge_best': (ctypes.c_int, [vp, vp, i32, vp]),
        'mb2_read_result': (ctypes.c_int, [vp, ctypes.POINTER(Result), c_double_p, vp]),
        'mb2_eval_cost': (ctypes.c_int, [vp, vp, i64, vp, vp]),
        'mb2_launch_info': (ctypes.c_int, [vp, c_int32_p, c_int32_p, c_int32_p]),
        'mb2_probe_pool_geometry': (ctypes.c_int, [i32, c_int32_p]),
        'mb2_launch_count': (i64, [vp]),
        'mb2_kernel_family': (ctypes.c_int, [vp]),
        'mb2_selftest_division': (ctypes.c_int, [vp, i64, u64, dbl, c_int64_p]),
        'mb2_set_islands': (ctypes.c_int, [vp, i32]),
        'mb2_read_islands': (ctypes.c_int, [vp, ctypes.c_void_p, c_double_p, vp]),
        'mb2_set_island_groups': (ctypes.c_int, [vp, i32]),
        'mb2_run_result_islands': (ctypes.c_int, [vp, c_int32_p, c_int32_p, vp]),
        'mb2_exchange_open': (ctypes.c_int, [vp, i32, i32, vp, i32]),
        'mb2_exchange_connect': (ctypes.c_int, [vp, vp, i32]),
        'mb2_exchange_best': (ctypes.c_int, [vp, vp]),
        'mb2_exchange_connected': (ctypes.c_int, [vp]),
        'mb2_exchange_reset': (ctypes.c_int, [vp]),
        'mb2_selftest_cosine': (ctypes.c_int, [vp, i64, u64, dbl, c_double_p]),
        'mb2_restore_state': (ctypes.c_int, [vp, c_double_p, dbl, i64, c_double_p, i64]),
        'mb2_set_termination': (ctypes.c_int, [vp, ctypes.POINTER(Termination), c_double_p, i64]),
        'mb2_run': (ctypes.c_int, [vp, i32, vp, vp]),
        'mb2_run_result': (ctypes.c_int, [vp, c_int32_p, c_int32_p, vp]),
        'mb2_probe_fp64_peak': (ctypes.c_int, [ctypes.c_int, i32, c_double_p, c_double_p, vp]),
        'mb2_set_commit_mode': (ctypes.c_int, [vp, ctypes.c_int]),
        'mb2_last_commit_in_place': (ctypes.c_int, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.mb2_abi_version() != 2:
        raise ImportError('mystic_b200: ABI version mismatch in %s' % LIBPATH)
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise NativeError('mystic_b200 native call failed (%d): %s' % (rc, lib.mb2_last_error().decode()))
