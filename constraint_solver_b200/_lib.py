"""ctypes binding of libcs_b200.so (the C ABI declared in include/cs_b200.h).

This is the only way the Python host layer reaches the device: there is no CPU fallback
and no other backend.  Importing works without a GPU (so the symbol table can be checked);
creating a handle without a GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CS_B200_LIB: an alternative build of the SAME library (A/B experiments with compile-time knobs)
LIB_PATH = os.environ.get("CS_B200_LIB") or os.path.join(_HERE, "libcs_b200.so")

CS_OK = 0
CS_ERR_INVALID_ARG = -1
CS_ERR_CUDA = -2
CS_ERR_NO_DEVICE = -3
CS_ERR_OOM = -4
CS_ERR_STATE = -5
CS_ERR_UNSUPPORTED = -6

CS_NQ_SWAP, CS_NQ_CHANGE = 0, 1
CS_NQ_MAX_N_SMEM = 16384
CS_NQ_MAX_N = 1_000_000
CS_NQ_FLAG_GLOBAL = 1
CS_NQ_FLAG_SCALAR = 2
CS_NQ_FLAG_REFERENCE_PROPOSER = 4
CS_ES_FLAG_REFERENCE_PROPOSER = 1
CHAIN_RUNNING, CHAIN_BEST, CHAIN_STALLED, CHAIN_EMPTY = 0, 1, 2, 3
PHILOX_INIT, PHILOX_PERTURB, PHILOX_LS, PHILOX_HOLIDAYS = 0, 1, 2, 3


class CsError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        msg = f"{where}: status {status}"
        if detail:
            msg += f" ({detail})"
        super().__init__(msg)


class CsMove(C.Structure):
    _fields_ = [("a", C.c_uint32), ("b", C.c_uint32)]


class CsStepStats(C.Structure):
    _fields_ = [
        ("moves_scored", C.c_uint64),
        ("steps_accepted", C.c_uint64),
        ("best_score", C.c_int64),
        ("best_chain", C.c_uint32),
        ("chains_at_best", C.c_uint32),
        ("device_ms", C.c_float),
        ("kernel_launches", C.c_uint32),
    ]


class CsNqConfig(C.Structure):
    _fields_ = [
        ("n", C.c_uint32),
        ("n_chains", C.c_uint32),
        ("chain_offset", C.c_uint32),
        ("trace_capacity", C.c_uint32),
        ("seed", C.c_uint64),
        ("device", C.c_int32),
        ("neighbourhood", C.c_uint32),
        ("flags", C.c_uint32),
    ]


class CsEsConfig(C.Structure):
    _fields_ = [
        ("n_days", C.c_uint32),
        ("n_employees", C.c_uint32),
        ("start_weekday", C.c_uint32),
        ("n_chains", C.c_uint32),
        ("chain_offset", C.c_uint32),
        ("trace_capacity", C.c_uint32),
        ("seed", C.c_uint64),
        ("device", C.c_int32),
        ("flags", C.c_uint32),
    ]


class CsEsMove(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("a", C.c_uint32), ("b", C.c_uint32)]


class CsEsStepStats(C.Structure):
    _fields_ = [
        ("moves_scored", C.c_uint64),
        ("steps_accepted", C.c_uint64),
        ("best_hard", C.c_int64),
        ("best_soft", C.c_int64),
        ("best_chain", C.c_uint32),
        ("chains_at_best", C.c_uint32),
        ("chains_feasible", C.c_uint32),
        ("device_ms", C.c_float),
        ("kernel_launches", C.c_uint32),
    ]


class CsIlsStats(C.Structure):
    _fields_ = [
        ("moves_scored", C.c_uint64),
        ("ls_steps", C.c_uint64),
        ("best_key", C.c_int64),
        ("best_chain", C.c_uint32),
        ("chains_done", C.c_uint32),
        ("rounds_run", C.c_uint32),
        ("device_ms", C.c_float),
        ("kernel_launches", C.c_uint32),
    ]


CS_ES_CHANGE, CS_ES_SWAP = 0, 1
CS_ES_MAX_SLOTS = 192
CS_ES_MAX_DAYS = CS_ES_MAX_SLOTS
CS_ES_MAX_SHIFTS = 3

_P = C.POINTER
_VP = C.c_void_p

# name -> (restype, argtypes); every symbol include/cs_b200.h declares
SIGNATURES = {
    "cs_abi_version": (C.c_int32, []),
    "cs_device_count": (C.c_int32, []),
    "cs_philox4x32_10": (None, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, _P(C.c_uint32)]),
    "cs_status_string": (C.c_char_p, [C.c_int32]),
    "cs_microbench": (C.c_int32, [C.c_int32, C.c_uint32, _P(C.c_double), _P(C.c_double)]),
    "cs_nq_create": (C.c_int32, [_P(CsNqConfig), _P(_VP)]),
    "cs_nq_destroy": (C.c_int32, [_VP]),
    "cs_nq_last_error": (C.c_char_p, [_VP]),
    "cs_nq_set_stream": (C.c_int32, [_VP, _VP]),
    "cs_nq_init_random": (C.c_int32, [_VP]),
    "cs_nq_set_chains": (C.c_int32, [_VP, C.c_uint32, C.c_uint32, _VP]),
    "cs_nq_get_chains": (C.c_int32, [_VP, C.c_uint32, C.c_uint32, _VP]),
    "cs_nq_get_scores": (C.c_int32, [_VP, _VP]),
    "cs_nq_get_status": (C.c_int32, [_VP, _VP]),
    "cs_nq_score_full": (C.c_int32, [_VP, C.c_uint32, _P(C.c_int64)]),
    "cs_nq_eval_moves": (C.c_int32, [_VP, C.c_uint32, C.c_uint32, _VP, C.c_uint64, _VP]),
    "cs_nq_enumerate": (C.c_int32, [_VP, C.c_uint32, _VP, C.c_uint64, _P(C.c_uint64)]),
    "cs_nq_neighbourhood_deltas": (C.c_int32, [_VP, C.c_uint32, _VP, C.c_uint64, _P(C.c_uint64)]),
    "cs_nq_band_deltas": (C.c_int32, [_VP, C.c_uint32, C.c_uint32, C.c_uint32, _VP, C.c_uint64, _P(C.c_uint64)]),
    "cs_nq_set_window": (C.c_int32, [_VP, C.c_uint64]),
    "cs_nq_set_chains_async": (C.c_int32, [_VP, C.c_uint32, C.c_uint32, _VP]),
    "cs_nq_commit_chains": (C.c_int32, [_VP]),
    "cs_nq_step": (C.c_int32, [_VP, C.c_uint32, _P(CsStepStats)]),
    "cs_nq_step_enqueue": (C.c_int32, [_VP, C.c_uint32]),
    "cs_nq_step_wait": (C.c_int32, [_VP, _P(CsStepStats)]),
    "cs_nq_exchange_select": (C.c_int32, [_VP, _VP, _VP, C.c_uint32]),
    "cs_nq_local_search": (C.c_int32, [_VP, C.c_uint64, C.c_uint64, _P(CsStepStats)]),
    "cs_nq_get_best_chains": (C.c_int32, [_VP, C.c_uint32, C.c_uint32, _VP, _VP]),
    "cs_nq_local_search_one": (C.c_int32, [_VP, _VP, C.c_uint64, C.c_uint64, _VP, _P(C.c_int64)]),
    "cs_nq_get_trace": (C.c_int32, [_VP, C.c_uint32, _VP, _VP, C.c_uint64, _P(C.c_uint64)]),
    "cs_nq_best": (C.c_int32, [_VP, _VP, _P(C.c_int64), _P(C.c_uint32)]),
    "cs_nq_best_key_device_ptr": (C.c_int32, [_VP, _P(_VP)]),
    "cs_nq_set_chain_u16_device": (C.c_int32, [_VP, C.c_uint32, _VP]),
    "cs_nq_chain_device_ptr": (C.c_int32, [_VP, C.c_uint32, _P(_VP), _P(C.c_uint32)]),
    "cs_nq_set_partition": (C.c_int32, [_VP, C.c_uint32, C.c_uint32]),
    "cs_nq_part_scan": (C.c_int32, [_VP]),
    "cs_nq_part_key_device_ptr": (C.c_int32, [_VP, _P(_VP)]),
    "cs_nq_part_apply": (C.c_int32, [_VP, _P(CsStepStats)]),
    "cs_nq_ils_init": (C.c_int32, [_VP, C.c_uint32, C.c_uint32]),
    "cs_nq_ils_run": (C.c_int32, [_VP, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, _P(CsIlsStats)]),
    "cs_nq_ils_get_best": (C.c_int32, [_VP, C.c_uint32, _VP, _P(C.c_int64)]),
    "cs_nq_ils_get_log": (C.c_int32, [_VP, C.c_uint32, _VP, _VP, C.c_uint64, _P(C.c_uint64)]),
    "cs_es_ils_init": (C.c_int32, [_VP, C.c_uint32, C.c_uint32]),
    "cs_es_ils_run": (C.c_int32, [_VP, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, _P(CsIlsStats)]),
    "cs_es_ils_get_best": (C.c_int32, [_VP, C.c_uint32, _VP, _P(C.c_int64), _P(C.c_int64)]),
    "cs_es_ils_get_log": (C.c_int32, [_VP, C.c_uint32, _VP, _VP, C.c_uint64, _P(C.c_uint64)]),
    "cs_es_create": (C.c_int32, [_P(CsEsConfig), _VP, _VP, _VP, C.c_uint64, _P(_VP)]),
    "cs_es_create_ex": (C.c_int32, [_P(CsEsConfig), _VP, _VP, _VP, C.c_uint64, C.c_uint32, _VP, _P(_VP)]),
    "cs_es_get_dims": (C.c_int32, [_VP, _P(C.c_uint32), _P(C.c_uint32), _P(C.c_uint32)]),
    "cs_es_score_full_ex": (C.c_int32, [_VP, C.c_uint32, _P(C.c_int64), _P(C.c_int64), _P(C.c_int64)]),
    "cs_es_destroy": (C.c_int32, [_VP]),
    "cs_es_last_error": (C.c_char_p, [_VP]),
    "cs_es_set_stream": (C.c_int32, [_VP, _VP]),
    "cs_es_set_window": (C.c_int32, [_VP, C.c_uint64]),
    "cs_es_init_random": (C.c_int32, [_VP]),
    "cs_es_set_chains": (C.c_int32, [_VP, C.c_uint32, C.c_uint32, _VP]),
    "cs_es_set_chains_async": (C.c_int32, [_VP, C.c_uint32, C.c_uint32, _VP]),
    "cs_es_commit_chains": (C.c_int32, [_VP]),
    "cs_es_get_chains": (C.c_int32, [_VP, C.c_uint32, C.c_uint32, _VP]),
    "cs_es_get_scores": (C.c_int32, [_VP, _VP, _VP]),
    "cs_es_get_status": (C.c_int32, [_VP, _VP]),
    "cs_es_score_full": (C.c_int32, [_VP, C.c_uint32, _P(C.c_int64), _P(C.c_int64), _P(C.c_int64)]),
    "cs_es_eval_moves": (C.c_int32, [_VP, C.c_uint32, _VP, C.c_uint64, _VP, _VP]),
    "cs_es_enumerate": (C.c_int32, [_VP, C.c_uint32, _VP, C.c_uint64, _P(C.c_uint64)]),
    "cs_es_neighbourhood_deltas": (C.c_int32, [_VP, C.c_uint32, _VP, _VP, C.c_uint64, _P(C.c_uint64)]),
    "cs_es_step": (C.c_int32, [_VP, C.c_uint32, _P(CsEsStepStats)]),
    "cs_es_step_enqueue": (C.c_int32, [_VP, C.c_uint32]),
    "cs_es_step_wait": (C.c_int32, [_VP, _P(CsEsStepStats)]),
    "cs_es_exchange_select": (C.c_int32, [_VP, _VP, _VP, C.c_uint32]),
    "cs_es_local_search": (C.c_int32, [_VP, C.c_uint64, C.c_uint64, _P(CsEsStepStats)]),
    "cs_es_get_best_chains": (C.c_int32, [_VP, C.c_uint32, C.c_uint32, _VP, _VP, _VP]),
    "cs_es_local_search_one": (C.c_int32, [_VP, _VP, C.c_uint64, C.c_uint64, _VP, _P(C.c_int64), _P(C.c_int64)]),
    "cs_es_get_trace": (C.c_int32, [_VP, C.c_uint32, _VP, _VP, _VP, C.c_uint64, _P(C.c_uint64)]),
    "cs_es_best": (C.c_int32, [_VP, _VP, _P(C.c_int64), _P(C.c_int64), _P(C.c_uint32)]),
    "cs_es_best_key_device_ptr": (C.c_int32, [_VP, _P(_VP)]),
    "cs_es_chain_device_ptr": (C.c_int32, [_VP, C.c_uint32, _P(_VP), _P(C.c_uint32)]),
}

_lib = None


def load():
    """Load the CUDA extension; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or make -C constraint_solver_b200/csrc). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def status_string(status: int) -> str:
    return load().cs_status_string(status).decode()
