"""ctypes binding of ``libqmcb200.so`` (C ABI in ``include/qmcb200.h``).

There is no CPU fallback: if the shared library is missing, or no CUDA device
is present when an engine is created, the call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# QMCB_LIB selects a tuning variant built by scripts/ (development aid)
LIB_PATH = os.environ.get('QMCB_LIB') or os.path.join(_HERE, 'libqmcb200.so')

# Every symbol include/qmcb200.h declares (tests check the export table).
SYMBOLS = [
    'qmcb_create', 'qmcb_destroy', 'qmcb_last_error', 'qmcb_version',
    'qmcb_model_eval', 'qmcb_model_eval_device', 'qmcb_fourier_density',
    'qmcb_dmc_init', 'qmcb_dmc_set_state', 'qmcb_dmc_run_block',
    'qmcb_dmc_get_state', 'qmcb_dmc_get_next', 'qmcb_set_profiling',
    'qmcb_last_block_stats', 'qmcb_measure_fp64_peak', 'qmcb_comm_unique_id', 'qmcb_comm_init',
    'qmcb_dmc_rebalance', 'qmcb_vmc_init', 'qmcb_vmc_run_block',
    'qmcb_vmc_get_state', 'qmcb_measure_fp64_sustained', 'qmcb_stream',
    'qmcb_host_alloc', 'qmcb_host_free', 'qmcb_rebalance_plan',
    'qmcb_one_body_density', 'qmcb_one_body_density_device',
    'qmcb_fourier_density_k', 'qmcb_set_model_params', 'qmcb_cs_load',
    'qmcb_cs_variance', 'qmcb_dmc_reblock_reset', 'qmcb_dmc_reblock_get',
    'qmcb_vmc_run_chain', 'qmcb_vmc_one_body_density',
]


class EngineError(RuntimeError):
    """A qmcb_* call returned a negative status."""


class ModelParams(C.Structure):
    _fields_ = [('model', C.c_double * 12), ('obf', C.c_double * 7),
                ('tbf', C.c_double * 6)]


class DMCParams(C.Structure):
    _fields_ = [
        ('time_step', C.c_double), ('nwc_factor', C.c_double),
        ('lower_bound', C.c_double), ('upper_bound', C.c_double),
        ('max_num_walkers', C.c_int64), ('target_num_walkers', C.c_int64),
        ('rng_seed', C.c_uint64), ('energy_mode', C.c_int32),
        ('ssf_num_modes', C.c_int32), ('ssf_as_pure', C.c_int32),
        ('density_num_bins', C.c_int32), ('density_as_pure', C.c_int32),
        ('reserved0', C.c_int32), ('ssf_pfw_nts', C.c_int64),
        ('density_pfw_nts', C.c_int64), ('local_capacity', C.c_int64)]


class VMCParams(C.Structure):
    _fields_ = [
        ('move_spread', C.c_double), ('lower_bound', C.c_double),
        ('upper_bound', C.c_double), ('rng_seed', C.c_uint64),
        ('chain_offset', C.c_int64), ('ssf_num_modes', C.c_int32),
        ('proposal', C.c_int32)]


class StateScalars(C.Structure):
    _fields_ = [
        ('energy', C.c_double), ('weight', C.c_double),
        ('ref_energy', C.c_double), ('accum_energy', C.c_double),
        ('total_energy', C.c_double), ('total_weight', C.c_double),
        ('num_walkers', C.c_int64), ('max_num_walkers', C.c_int64),
        ('step', C.c_int64), ('capacity_hits', C.c_int64)]


_lib = None


def load():
    """Load the engine library; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f'{LIB_PATH} not found: build it with '
            f'`python -m phd_qmclib_b200.build` (nvcc, sm_100a). '
            f'There is no CPU fallback.')
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    L.qmcb_create.argtypes = [C.POINTER(ModelParams), C.c_int,
                              C.POINTER(vp)]
    L.qmcb_destroy.argtypes = [vp]
    L.qmcb_destroy.restype = None
    L.qmcb_last_error.argtypes = [vp]
    L.qmcb_last_error.restype = C.c_char_p
    L.qmcb_version.restype = C.c_char_p
    L.qmcb_model_eval.argtypes = [vp, vp, i64, vp, vp, vp]
    L.qmcb_model_eval_device.argtypes = [vp, vp, i64, vp, vp, vp]
    L.qmcb_fourier_density.argtypes = [vp, vp, i64, i32, vp]
    L.qmcb_one_body_density.argtypes = [vp, vp, i64, vp, i32, vp]
    L.qmcb_one_body_density_device.argtypes = [vp, vp, i64, vp, i32, vp]
    L.qmcb_fourier_density_k.argtypes = [vp, vp, i64, vp, i32, vp]
    L.qmcb_set_model_params.argtypes = [vp, C.POINTER(ModelParams)]
    L.qmcb_cs_load.argtypes = [vp, vp, i64, vp]
    L.qmcb_cs_variance.argtypes = [vp, C.POINTER(ModelParams),
                                   C.POINTER(dbl), C.POINTER(dbl), vp, vp]
    L.qmcb_dmc_init.argtypes = [vp, C.POINTER(DMCParams), vp, i64, dbl, i64]
    L.qmcb_dmc_set_state.argtypes = [vp, C.POINTER(DMCParams), vp, vp, vp,
                                     vp, C.POINTER(StateScalars), i64]
    L.qmcb_dmc_run_block.argtypes = [vp, i64, i32, vp, vp, vp, vp, vp, vp,
                                     vp]
    L.qmcb_dmc_get_state.argtypes = [vp, vp, vp, vp, vp, vp,
                                     C.POINTER(StateScalars)]
    L.qmcb_dmc_get_next.argtypes = [vp, vp, vp, vp, vp,
                                    C.POINTER(StateScalars)]
    L.qmcb_set_profiling.argtypes = [vp, i32]
    L.qmcb_dmc_reblock_reset.argtypes = [vp, i32]
    L.qmcb_dmc_reblock_get.argtypes = [vp, vp, vp, vp]
    L.qmcb_last_block_stats.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl),
                                        C.POINTER(i64)]
    L.qmcb_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(dbl), C.POINTER(dbl)]
    L.qmcb_measure_fp64_sustained.argtypes = [C.c_int, dbl, C.POINTER(dbl)]
    L.qmcb_stream.argtypes = [vp]
    L.qmcb_stream.restype = vp
    L.qmcb_host_alloc.argtypes = [i64]
    L.qmcb_host_alloc.restype = vp
    L.qmcb_host_free.argtypes = [vp]
    L.qmcb_host_free.restype = None
    L.qmcb_comm_unique_id.argtypes = [vp]
    L.qmcb_comm_init.argtypes = [vp, vp, i32, i32]
    L.qmcb_dmc_rebalance.argtypes = [vp, C.POINTER(i64)]
    L.qmcb_rebalance_plan.argtypes = [vp, i32, i32, vp, vp, C.POINTER(i64)]
    L.qmcb_vmc_init.argtypes = [vp, C.POINTER(VMCParams), vp, i64]
    L.qmcb_vmc_run_block.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp]
    L.qmcb_vmc_run_chain.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    L.qmcb_vmc_one_body_density.argtypes = [vp, vp, i32, vp]
    L.qmcb_vmc_get_state.argtypes = [vp, vp, vp]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ('qmcb_destroy', 'qmcb_last_error', 'qmcb_version',
                        'qmcb_stream', 'qmcb_host_alloc', 'qmcb_host_free'):
            fn.restype = C.c_int
    _lib = L
    return L


def ptr(a):
    """Pointer to a C-contiguous numpy array (or NULL)."""
    if a is None:
        return None
    assert a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(C.c_void_p)


def model_params_struct(block) -> ModelParams:
    block = np.ascontiguousarray(block, dtype=np.float64)
    if block.shape != (25,):
        raise ValueError('parameter block must hold 25 doubles')
    mp = ModelParams()
    mp.model[:] = block[:12].tolist()
    mp.obf[:] = block[12:19].tolist()
    mp.tbf[:] = block[19:].tolist()
    return mp
