"""ctypes binding of libehmm_b200.so (the C ABI declared in include/ehmm.h).

There is no fallback: if the shared library is missing or fails to load, importing
this module raises.  Build it with ``python epigrahmm_b200/build.py``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libehmm_b200.so")

EHMM_OK, EHMM_ERR_ARG, EHMM_ERR_CUDA, EHMM_ERR_STATE, EHMM_ERR_NOMEM, EHMM_ERR_RNG = range(6)


class EhmmError(RuntimeError):
    """Raised for any non-zero status from the C ABI (the Rcpp glue would call Rcpp::stop)."""

    def __init__(self, code: int, message: str):
        super().__init__("ehmm error %d: %s" % (code, message))
        self.code = code


if not os.path.exists(SO_PATH):
    raise ImportError(
        "%s not found: the CUDA library has not been built (python epigrahmm_b200/build.py). "
        "epigrahmm_b200 has no CPU fallback." % SO_PATH)

lib = C.CDLL(SO_PATH)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_up = C.POINTER(C.c_uint32)
_bp = C.POINTER(C.c_uint8)
_lp = C.POINTER(C.c_int64)
_sess = C.c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPE)
PROTOTYPES = {
    "ehmm_last_error": [],
    "ehmm_version": [],
    "ehmm_device_count": [C.POINTER(C.c_int)],
    "ehmm_session_open": [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.POINTER(_sess)],
    "ehmm_session_find": [C.c_char_p, C.POINTER(_sess)],
    "ehmm_session_close": [_sess],
    "ehmm_session_stream": [_sess, C.POINTER(C.c_void_p)],
    "ehmm_session_sync": [_sess],
    "ehmm_session_set_block": [_sess, C.c_int64, C.c_int64],
    "ehmm_set_data": [_sess, C.c_int, _ip, _dp, _dp],
    "ehmm_set_data_device": [_sess, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p],
    "ehmm_controls_first": [_sess, _dp],
    "ehmm_set_controls_first": [_sess, C.c_double],
    "ehmm_set_groups": [_sess, _ip, C.c_int64, _dp, _dp, _dp, _bp],
    "ehmm_build_groups": [_sess, _lp],
    "ehmm_groups_fetch": [_sess, _ip, _dp, _dp, _dp, _bp],
    "ehmm_groups_first_cell": [_sess, _lp],
    "ehmm_groups_remap": [_sess, _ip, C.c_int64, _dp, _dp, _dp, _bp],
    "ehmm_set_data_encoded": [_sess, C.c_int, _ip, C.c_int64, _dp, _dp, _dp, _bp],
    "ehmm_set_data_encoded16": [_sess, C.c_int, C.POINTER(C.c_uint16), C.c_int64, _dp, _dp, _dp, _bp],
    "ehmm_prefetch_groups": [_sess, C.c_int, C.c_void_p, C.c_int],
    "ehmm_set_design_n": [_sess, C.c_int, C.c_int, _bp],
    "ehmm_set_design": [_sess, C.c_int, _bp],
    "ehmm_loglik_consensus": [_sess, _dp, _dp, C.c_int, C.c_double],
    "ehmm_loglik_consensus_zinb": [_sess, _dp, _dp, C.c_int, C.c_double],
    "ehmm_loglik_init": [_sess, _dp, _dp],
    "ehmm_loglik_differential_hmm": [_sess, _dp, _dp, C.c_double],
    "ehmm_loglik_differential_hmm_zinb": [_sess, _dp, _dp, C.c_double],
    "ehmm_inner_expstep": [_sess, _dp, _dp, C.c_double],
    "ehmm_inner_expstep_zinb": [_sess, _dp, _dp, C.c_double],
    "ehmm_expstep": [_sess, _dp, _dp, _dp],
    "ehmm_expstep_device": [_sess, _dp, _dp, C.c_void_p],
    "ehmm_loglik_value": [_sess, _dp],
    "ehmm_maxstepprob": [_sess, _dp, _dp],
    "ehmm_qfunction": [_sess, _dp, _dp, _dp],
    "ehmm_bic": [_sess, C.c_int, C.c_int, _dp],
    "ehmm_inner_maxstepprob": [_sess, _dp],
    "ehmm_marginal_probability": [_sess, _dp],
    "ehmm_viterbi": [_sess, _dp, _dp, _dp],
    "ehmm_call_peaks": [_sess, C.c_double, _bp, _lp, _dp],
    "ehmm_call_peaks_viterbi": [_sess, _bp, _lp],
    "ehmm_call_peaks_exact": [_sess, C.c_int],
    "ehmm_call_peaks_path": [_sess, C.POINTER(C.c_int)],
    "ehmm_peak_runs": [_sess, _lp, C.c_int, C.c_int64, _lp, _lp, _lp],
    "ehmm_flush": [_sess, _dp, _dp, _dp, _dp, _dp, _dp, _dp],
    "ehmm_expstep_mode": [_sess, C.c_int],
    "ehmm_materialize": [_sess],
    "ehmm_viterbi_mode": [_sess, C.c_int],
    "ehmm_viterbi_info": [_sess, C.POINTER(C.c_int64)],
    "ehmm_viterbi_forward": [_sess, _dp, _dp, _dp, _dp],
    "ehmm_viterbi_backtrace": [_sess, C.c_int, C.POINTER(C.c_int), _dp],
    "ehmm_expstep_reduce": [_sess, _dp, _dp, _dp, _dp],
    "ehmm_expstep_apply": [_sess, _dp, _dp, _dp],
    "ehmm_expstep_apply_ops": [_sess, _dp, C.c_int, C.c_int],
    "ehmm_estep_stats": [_sess, _dp],
    "ehmm_estep_row0": [_sess, _dp],
    "ehmm_estep_set_stats": [_sess, _dp, _dp, C.c_double],
    "ehmm_rng_set_state": [_sess, _up],
    "ehmm_rng_get_state": [_sess, _up],
    "ehmm_rc_consensus": [_sess, C.c_double, _dp, C.c_int64, _lp],
    "ehmm_rc_differential": [_sess, C.c_double, C.c_int, _dp, C.c_int64, _lp],
    "ehmm_rc_count": [_sess, C.c_int, C.c_double, _lp],
    "ehmm_rc_apply_at": [_sess, C.c_int, C.c_double, _lp, C.c_int64],
    "ehmm_rc_buckets_device": [_sess, C.POINTER(C.c_void_p), _lp, C.POINTER(C.c_int)],
    "ehmm_rc_finalize": [_sess],
    "ehmm_rc_hint": [_sess, C.c_double],
    "ehmm_inner_mstep_sums": [_sess, _dp],
    "ehmm_rc_table_rows": [_sess, C.c_int, _lp],
    "ehmm_rc_table_fetch": [_sess, C.c_int, _dp, _dp],
    "ehmm_rc_buckets_fetch": [_sess, C.c_int, _dp],
    "ehmm_reweight": [C.c_int, _dp, C.c_int64, C.c_double, _dp, C.c_int64, _lp],
    "ehmm_aggregate": [C.c_int, _dp, C.c_int64, _ip, C.c_int64, _dp],
    "ehmm_glm_nb": [_sess, C.c_int, _dp, C.c_int, _dp, _dp],
    "ehmm_glm_zinb": [_sess, C.c_int, _dp, C.c_int, C.c_double, _dp, _dp],
    "ehmm_glm_nb_batch": [_sess, C.c_int, C.POINTER(C.c_int), _dp, C.c_int, _dp, _dp],
    "ehmm_glm_mix_nb": [_sess, _dp, C.c_double, _dp, _dp],
    "ehmm_glm_mix_zinb": [_sess, _dp, C.c_double, _dp, _dp],
    "ehmm_profile_kernels": [],
    "ehmm_profile_name": [C.c_int],
    "ehmm_profile": [_sess, C.c_int],
    "ehmm_profile_read": [_sess, _lp, _dp],
    "ehmm_dataset_dims": [_sess, C.c_char_p, _lp, _lp],
    "ehmm_fetch": [_sess, C.c_char_p, _dp, C.c_int64],
    "ehmm_device_ptr": [_sess, C.c_char_p, C.POINTER(C.c_void_p)],
    "ehmm_store": [_sess, C.c_char_p, _dp, C.c_int64, C.c_int64],
    "ehmm_comm_open": [C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_double, C.POINTER(C.c_void_p)],
    "ehmm_comm_allgather": [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p],
    "ehmm_comm_close": [C.c_void_p],
}
_RESTYPE = {"ehmm_last_error": C.c_char_p, "ehmm_profile_name": C.c_char_p}

for _name, _args in PROTOTYPES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header and library out of sync
    _fn.argtypes = _args
    _fn.restype = _RESTYPE.get(_name, C.c_int)


def last_error() -> str:
    return lib.ehmm_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != EHMM_OK:
        raise EhmmError(rc, last_error())


def dptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def iptr(a):
    return None if a is None else a.ctypes.data_as(_ip)


def bptr(a):
    return None if a is None else a.ctypes.data_as(_bp)
