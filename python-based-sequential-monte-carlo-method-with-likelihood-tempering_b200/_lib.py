"""ctypes binding of libsmcb200.so (the C-ABI declared in include/smcb200.h).

There is no CPU fallback: if the library is missing it is built with nvcc; if that fails, or if
no CUDA device is present when a handle is created, a RuntimeError is raised.
"""
import ctypes as C
import os

from . import _build

_LIB = None

c_i64, c_u64, c_u32, c_int, c_dbl = C.c_int64, C.c_uint64, C.c_uint32, C.c_int, C.c_double
p_void = C.c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPE)
_SIGNATURES = {
    "smcb_version": [],
    "smcb_create": [c_int, C.POINTER(p_void)],
    "smcb_destroy": [p_void],
    "smcb_last_error": [p_void],
    "smcb_reserve": [p_void, c_i64, c_int],
    "smcb_launch_count": [p_void],
    "smcb_set_data_mm_progress": [p_void, p_void, p_void, p_void, c_int, c_int],
    "smcb_set_data_mm_rate": [p_void, p_void, p_void, c_i64, c_int],
    "smcb_set_data_mm_rate_sufficient": [p_void, p_void, p_void, c_i64, c_dbl, c_dbl],
    "smcb_set_data_kinetic": [p_void, p_void, p_void, c_int, p_void, c_int, p_void, c_int, c_int],
    "smcb_loglik": [p_void, c_int, p_void, c_i64, c_i64, c_int, p_void, p_void, p_void],
    "smcb_loglik_bounded": [p_void, c_int, p_void, c_i64, c_i64, c_int, p_void, p_void, p_void, p_void],
    "smcb_set_param": [p_void, c_int, c_dbl],
    "smcb_profile_read": [p_void, p_void],
    "smcb_predict_mm_progress": [p_void, p_void, c_i64, c_i64, p_void, p_void],
    "smcb_loglik_stats": [p_void, p_void],
    "smcb_lk_max": [p_void, p_void, c_i64, p_void, p_void],
    "smcb_temper_sums": [p_void, p_void, c_i64, p_void, p_void, c_int, p_void, p_void],
    "smcb_weights": [p_void, p_void, c_i64, p_void, c_dbl, p_void, p_void, p_void],
    "smcb_resample_counts": [p_void, p_void, c_i64, c_i64, c_dbl, c_int, p_void, c_u64, c_i64, p_void, p_void,
                             p_void],
    "smcb_resample_totals": [p_void, p_void, c_i64, c_i64, p_void, p_void],
    "smcb_ancestors": [p_void, p_void, c_i64, c_i64, p_void, p_void, p_void],
    "smcb_resample_fused": [p_void, p_void, p_void, c_i64, c_i64, c_u64, c_i64, c_i64, p_void, c_dbl, p_void, c_dbl, p_void,
                            c_i64, c_int, p_void, c_i64, p_void, p_void, p_void, p_void],
    "smcb_gather": [p_void, p_void, c_i64, p_void, c_i64, c_int, p_void, c_i64, p_void],
    "smcb_colsum": [p_void, p_void, c_i64, c_i64, c_int, p_void, p_void],
    "smcb_centered_moments": [p_void, p_void, c_i64, c_i64, c_int, p_void, p_void, p_void],
    "smcb_mh_propose": [p_void, p_void, c_i64, c_i64, c_int, p_void, c_dbl, p_void, p_void, p_void, c_u64, c_u64,
                        c_u32, c_u32, p_void, c_i64, p_void, p_void],
    "smcb_mh_accept": [p_void, p_void, c_i64, p_void, p_void, c_i64, p_void, p_void, c_i64, c_int, c_dbl, p_void,
                       p_void, c_u64, c_u64, c_u32, c_u32, p_void, p_void, p_void],
    "smcb_prior_logratio": [p_void, p_void, c_i64, p_void, c_i64, c_i64, c_int, p_void, p_void, p_void, p_void, p_void],
    "smcb_mh_threshold": [p_void, p_void, p_void, c_i64, c_dbl, p_void, p_void, c_u64, c_u64, c_u32, c_u32, p_void,
                          p_void],
    "smcb_mh_fused": [p_void, c_int, p_void, c_i64, p_void, c_i64, c_int, p_void, c_dbl, p_void, p_void, c_dbl,
                      c_int, c_u64, c_u64, c_u32, c_u32, p_void, p_void, p_void],
    "smcb_mh_sweeps": [p_void, c_int, p_void, c_i64, p_void, c_i64, c_int, c_i64, p_void, c_dbl, p_void, p_void, c_dbl,
                       c_int, c_int, c_u64, c_u64, c_u32, c_u32, p_void, c_i64, p_void, p_void, p_void, p_void, p_void,
                       p_void, p_void],
    "smcb_philox_draws": [p_void, c_i64, c_int, c_u64, c_u64, c_u32, c_u32, p_void, p_void, p_void],
    "smcb_sample_uniform_box": [p_void, p_void, c_i64, c_i64, c_int, p_void, p_void, c_u64, c_u64, p_void],
    "smcb_measure_fma_peak": [p_void, p_void],
    "smcb_set_user_likelihood": [p_void, p_void, p_void],
    "smcb_zero": [p_void, p_void, c_i64, p_void],
    "smcb_copy_rows": [p_void, p_void, c_i64, p_void, c_i64, c_i64, c_int, p_void],
    "smcb_temper_eval": [p_void, p_void, c_i64, p_void, c_int, p_void, p_void],
    "smcb_mh_propose_dev": [p_void, p_void, c_i64, c_i64, c_int, p_void, c_dbl, p_void, p_void, p_void, c_u64, c_u64,
                            c_u32, c_u32, p_void, c_i64, p_void, p_void],
    "smcb_moments_merged": [p_void, p_void, c_i64, c_i64, c_int, c_i64, p_void, p_void, p_void, p_void],
    "smcb_comm_unique_id": [p_void, c_int],
    "smcb_comm_init": [p_void, p_void, c_int, c_int, c_int],
    "smcb_comm_destroy": [p_void],
    "smcb_comm_rank": [p_void],
    "smcb_comm_world": [p_void],
    "smcb_comm_all_gather": [p_void, p_void, p_void, c_i64, p_void],
    "smcb_comm_all_reduce_f64": [p_void, p_void, c_i64, c_int, p_void],
    "smcb_comm_broadcast": [p_void, p_void, c_i64, c_int, p_void],
    "smcb_comm_all_to_all_v": [p_void, p_void, p_void, p_void, p_void, p_void],
    "smcb_comm_exchange_rows": [p_void, p_void, c_i64, p_void, p_void, c_i64, p_void, c_int, p_void],
    "smcb_collective_count": [p_void],
}
# host callback of a user-supplied likelihood (smcb_user_loglik_fn)
USER_LOGLIK_FN = C.CFUNCTYPE(c_int, p_void, p_void, c_i64, c_i64, c_int, p_void, p_void, p_void)
_RESTYPE = {"smcb_last_error": C.c_char_p, "smcb_launch_count": c_i64, "smcb_collective_count": c_i64}

EXPORTS = tuple(_SIGNATURES)

MODEL_MM_PROGRESS, MODEL_MM_RATE, MODEL_KINETIC_RK, MODEL_KINETIC_DAE, MODEL_USER = 1, 2, 3, 4, 5
OP_SUM, OP_MAX, COMM_ID_BYTES = 0, 1, 128
SCAN_SEQUENTIAL, SCAN_FIXED = 0, 1
MAX_DIM, MAX_CAND = 32, 16
KIN_NCOND_FIELDS = 10
PARAM_MM_BUDGET, PARAM_MM_REFILL_MIN, PARAM_MM_PATIENCE, PARAM_PROFILE, PARAM_MM_CHUNK, PARAM_MM_TAIL_WARPS = 1, 2, 3, 4, 5, 6
PARAM_MM_INTEGRATOR, MM_RK45_SCIPY, MM_EXACT = 7, 0, 1
N_STATS = 24


def library_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load (building first if needed) and return the ctypes library object."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if build_if_missing:
        # no-op when the library is newer than every source and the header; a stale library behind an edited
        # header would corrupt memory through ctypes instead of raising.  Where there is no nvcc (a box that
        # received the built library) an existing library is used as it is.
        try:
            _build.build(force=os.environ.get("SMCB_REBUILD") == "1")
        except RuntimeError:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing and could not be built; the CUDA library is required "
                           "(there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, c_int)
    if lib.smcb_version() != _build.header_version():
        raise RuntimeError(f"{path} is version {lib.smcb_version()}, include/smcb200.h declares "
                           f"{_build.header_version()}: rebuild with `python -m smcb200._build --force`")
    _LIB = lib
    return lib


class SmcbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsmcb200 error {code}: {msg}")
        self.code = code


def check(handle, rc):
    if rc != 0:
        msg = load().smcb_last_error(handle)
        raise SmcbError(rc, msg.decode() if msg else "?")
