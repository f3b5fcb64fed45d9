"""ctypes binding of libbm.so (include/bm.h).  There is no CPU fallback: if the library is missing or a
call fails, an exception is raised."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('BM_LIB_PATH') or os.path.join(HERE, '_lib', 'libbm.so')   # BM_LIB_PATH: tuning variants

BM_OK, BM_EINVAL, BM_ECUDA, BM_ECUSOLVER, BM_EZERO_PSI, BM_ENONFINITE, BM_ENOMEM = range(7)
RENORM = {'max': 0, 'l2': 1}
SVD = {'gesvdj': 0, 'gesvd': 1, 'gram': 2, 'gram_syevd': 3}
RULE = {'batched': 0, 'single': 1}

# every symbol include/bm.h declares: (name, restype, argtypes)
_P = C.c_void_p
_SIGS = [
    ('bm_version', C.c_int, []),
    ('bm_create', C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    ('bm_destroy', C.c_int, [_P]),
    ('bm_last_error', C.c_char_p, [_P]),
    ('bm_set_svd', C.c_int, [_P, C.c_int]),
    ('bm_eig_sym', C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, _P, _P, _P]),
    ('bm_eig_stats', C.c_int, [_P, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_int)]),
    ('bm_eig_fallback_reasons', C.c_int, [_P, _P]),
    ('bm_set_precision', C.c_int, [_P, C.c_int]),
    ('bm_set_data', C.c_int, [_P, _P, C.c_int]),
    ('bm_set_site', C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int]),
    ('bm_get_site', C.c_int, [_P, C.c_int, _P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ('bm_site_dims', C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ('bm_env_build', C.c_int, [_P, C.c_int]),
    ('bm_env_advance', C.c_int, [_P, C.c_int, C.c_int]),
    ('bm_env_get', C.c_int, [_P, C.c_int, C.c_int, _P, C.POINTER(C.c_int)]),
    ('bm_grad_dims', C.c_int, [_P, C.c_int] + [C.POINTER(C.c_int)] * 5),
    ('bm_step_grad', C.c_int, [_P, C.c_int, _P, C.c_int, _P]),
    ('bm_step_update', C.c_int, [_P, C.c_int, _P, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double,
                                 C.c_double, _P, _P]),
    ('bm_step', C.c_int, [_P, C.c_int, _P, C.c_int, _P, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double,
                          C.c_double, C.c_int, _P, _P]),
    ('bm_sweep', C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int, _P, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, _P, _P, _P]),
    ('bm_split_bond', C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_double, _P, _P]),
    ('bm_get_merged', C.c_int, [_P, _P]),
    ('bm_ref_to_engine', C.c_int, [_P, C.c_int, _P, _P]),
    ('bm_engine_to_ref', C.c_int, [_P, C.c_int, _P, _P]),
    ('bm_logpsi', C.c_int, [_P, _P, C.c_int, _P, _P]),
    ('bm_last_logpsi', C.c_int, [_P, _P]),
    ('bm_sample', C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_uint64, _P, _P]),
    ('bm_reconstruct_one', C.c_int, [_P, _P, _P, C.c_int, _P]),
    ('bm_reconstruct_batch', C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P]),
    ('bm_sample_device', C.c_int, [_P, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_float)]),
    ('bm_peer_export', C.c_int, [_P, _P]),
    ('bm_peer_import', C.c_int, [_P, _P, C.c_int, C.c_int]),
    ('bm_set_timing', C.c_int, [_P, C.c_int]),
    ('bm_get_timing', C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    ('bm_launch_count', C.c_longlong, [_P]),
    ('bm_stream', _P, [_P]),
]
EXPORTS = [s[0] for s in _SIGS]

_lib = None


class BornEngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f'libbm error {code}: {msg}')
        self.code = code


class ZeroPsiError(BornEngineError):
    """psi == 0 for some sample: the reference skips the update (TNutils.py:1244-1254)."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f'{LIB_PATH} is missing: build it with `python -m mps_born_machines_b200.build` '
                              '(there is no CPU fallback)')
        _lib = C.CDLL(LIB_PATH)
        for name, res, args in _SIGS:
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


def check(ctx, code):
    if code != BM_OK:
        msg = lib().bm_last_error(ctx)
        msg = msg.decode() if msg else ''
        raise (ZeroPsiError if code == BM_EZERO_PSI else BornEngineError)(code, msg)
