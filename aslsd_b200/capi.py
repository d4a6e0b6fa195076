"""ctypes binding of ``libaslsd_b200.so`` (the C ABI of include/aslsd_b200.h).

The library is built in-tree by ``aslsd_b200._build.build()`` (nvcc, sm_100a).
There is no CPU fallback: if the library is missing or no CUDA device can be
opened, the calls raise.
"""
import ctypes
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_lib',
                        'libaslsd_b200.so')

ASLSD_F = 1
ASLSD_DIFF = 2
ASLSD_GRAD_ALLOC = 4
OPT_LAG_QUADRATURE = 1
OPT_WEIGHT_CACHE = 2
OPT_ASYNC_EVENTS = 3
OPT_DEVICE_ASSEMBLY = 4

BASIS_EXPONENTIAL = 0
BASIS_GAUSSIAN = 1

ERR_UNSUPPORTED = -3

c_i64_p = ctypes.POINTER(ctypes.c_int64)
c_i32_p = ctypes.POINTER(ctypes.c_int32)
c_f64_p = ctypes.POINTER(ctypes.c_double)


class Basis(ctypes.Structure):
    _fields_ = [('type', ctypes.c_int32), ('free_mask', ctypes.c_int32),
                ('fixed', ctypes.c_double * 3)]


class Pcg64(ctypes.Structure):
    _fields_ = [('state_hi', ctypes.c_uint64), ('state_lo', ctypes.c_uint64),
                ('inc_hi', ctypes.c_uint64), ('inc_lo', ctypes.c_uint64),
                ('has_uint32', ctypes.c_int32), ('uinteger', ctypes.c_uint32)]


class SampleTerm(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('i', ctypes.c_int32), ('j', ctypes.c_int32),
                ('n_strata', ctypes.c_int32), ('sat_t', ctypes.c_double), ('v_sat', ctypes.c_double)]


# every symbol declared in include/aslsd_b200.h: name -> (restype, argtypes)
SYMBOLS = {
    'aslsd_create': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    'aslsd_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'aslsd_last_error': (ctypes.c_char_p, [ctypes.c_void_p]),
    'aslsd_set_stream': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'aslsd_last_timing': (ctypes.c_int, [ctypes.c_void_p, c_f64_p, c_i64_p]),
    'aslsd_last_moment_timing': (ctypes.c_int, [ctypes.c_void_p, c_f64_p]),
    'aslsd_exact_raw_used': (ctypes.c_int, [ctypes.c_void_p, c_i64_p]),
    'aslsd_exact_max_lag': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i32_p, ctypes.POINTER(c_f64_p),
                                           ctypes.c_int, c_f64_p]),
    'aslsd_measure_fp64_peak': (ctypes.c_int, [ctypes.c_void_p, c_f64_p]),
    'aslsd_set_events': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i64_p,
                                        ctypes.POINTER(c_f64_p), ctypes.c_double]),
    'aslsd_set_horizon': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double]),
    'aslsd_index_map_sizes': (ctypes.c_int, [ctypes.c_void_p, c_i64_p, c_i64_p]),
    'aslsd_index_maps': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(c_i64_p),
                                        ctypes.POINTER(c_i64_p), ctypes.POINTER(c_i64_p)]),
    'aslsd_set_model': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i32_p,
                                       ctypes.POINTER(Basis)]),
    'aslsd_model_n_param': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i32_p, c_i32_p]),
    'aslsd_exact_out_len': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i64_p, c_i64_p]),
    'aslsd_exact_sums': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_f64_p, ctypes.c_int,
                                        c_i64_p, c_i64_p, c_f64_p, c_i64_p]),
    'aslsd_point_sums': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_f64_p, ctypes.c_int, c_f64_p]),
    'aslsd_exact_plan_rows': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, c_i64_p, c_i32_p, c_f64_p]),
    'aslsd_rng_choice': (ctypes.c_int, [ctypes.POINTER(Pcg64), ctypes.c_int64, ctypes.c_int64,
                                        ctypes.c_int, c_i64_p]),
    'aslsd_sample_single_ix': (ctypes.c_int, [ctypes.POINTER(Pcg64), ctypes.c_int64, c_i64_p, c_i64_p, c_i64_p,
                                              c_i64_p]),
    'aslsd_sample_double_ix': (ctypes.c_int, [ctypes.POINTER(Pcg64), ctypes.c_int64, c_i64_p, c_i64_p,
                                              c_i64_p, c_i64_p, c_i64_p, ctypes.c_int64, c_i64_p, c_i64_p]),
    'aslsd_set_weights': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i32_p,
                                         ctypes.POINTER(c_f64_p)]),
    'aslsd_exact_sums_weighted': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_f64_p, ctypes.c_int,
                                                 c_i64_p, c_i64_p, c_i32_p,
                                                 ctypes.POINTER(ctypes.c_uint8), c_f64_p, c_i64_p]),
    'aslsd_exact_raw_len': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i64_p]),
    'aslsd_exact_raw_device': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_f64_p, ctypes.c_int,
                                              c_i64_p, c_i64_p, ctypes.c_void_p, c_i64_p]),
    'aslsd_exact_assemble': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_f64_p, ctypes.c_int,
                                            c_f64_p, c_f64_p]),
    'aslsd_last_eval_counts': (ctypes.c_int, [ctypes.c_void_p, c_i64_p, c_i64_p]),
    'aslsd_last_work_counts': (ctypes.c_int, [ctypes.c_void_p, c_i64_p, ctypes.c_int]),
    'aslsd_set_option': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    'aslsd_exact_sums_batch': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i32_p,
                                              ctypes.POINTER(c_f64_p), ctypes.c_int, c_i64_p, c_i64_p,
                                              ctypes.POINTER(c_f64_p), ctypes.POINTER(c_i64_p)]),
    'aslsd_exact_raw_len_batch': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i64_p]),
    'aslsd_exact_raw_device_batch': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i32_p,
                                                    ctypes.POINTER(c_f64_p), ctypes.c_int, c_i64_p,
                                                    c_i64_p, ctypes.c_void_p]),
    'aslsd_exact_assemble_batch': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i32_p,
                                                  ctypes.POINTER(c_f64_p), ctypes.c_int, c_f64_p,
                                                  ctypes.POINTER(c_f64_p)]),
    'aslsd_exact_packed_len_batch': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i32_p, c_i64_p]),
    'aslsd_exact_packed_device_batch': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i32_p,
                                                       ctypes.POINTER(c_f64_p), ctypes.c_int, c_i64_p,
                                                       c_i64_p, ctypes.c_void_p]),
    'aslsd_exact_assemble_map_batch': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i32_p,
                                                      ctypes.POINTER(c_f64_p), ctypes.c_int, c_f64_p,
                                                      ctypes.POINTER(c_f64_p)]),
    'aslsd_simulate': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i32_p, ctypes.POINTER(Basis), c_f64_p,
                                      ctypes.c_double, ctypes.c_uint64, c_i64_p]),
    'aslsd_simulate_fetch': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(c_f64_p)]),
    'aslsd_synchronize': (ctypes.c_int, [ctypes.c_void_p]),
    'aslsd_sampled_n_out': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, c_i32_p]),
    'aslsd_sampled_sums': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_f64_p, ctypes.c_int,
                                          ctypes.c_int, ctypes.POINTER(SampleTerm), c_i64_p,
                                          c_i64_p, c_i64_p, c_f64_p, c_f64_p, c_f64_p, c_f64_p,
                                          c_f64_p, c_f64_p]),
    'aslsd_sampled_sums_weighted': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_f64_p, ctypes.c_int,
                                                   ctypes.c_int, ctypes.POINTER(SampleTerm), c_i64_p,
                                                   c_i64_p, c_i64_p, c_f64_p, c_f64_p, c_i32_p, c_f64_p,
                                                   c_f64_p, c_f64_p, c_f64_p]),
    'aslsd_update_allocation': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_f64_p, c_f64_p,
                                               c_f64_p, c_f64_p]),
}

_lib = None


def load_library():
    """Load the shared library and declare every prototype.  Raises if the
    extension has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'aslsd_b200: %s is missing - build it with '
            '`python -c "import __graft_entry__ as g; g.build()"` '
            '(there is no CPU fallback)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def as_f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def ptr_f64(a):
    return a.ctypes.data_as(c_f64_p)


def ptr_i64(a):
    return a.ctypes.data_as(c_i64_p)
