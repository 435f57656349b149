"""constraint_solver_b200 -- B200-native (sm_100a) move-evaluation hot path of
asimihsan/constraint-solver's local-search crate, behind a C ABI (include/cs_b200.h).

(The directory is spelled with an underscore so it is importable; the task text calls the
package `constraint-solver_b200`.)
"""
from . import _lib
from ._lib import CsError, load
from .nqueens import (CHANGE, SWAP, NQueensChains, NQueensInitialSolutionGenerator,
                      NQueensMoveProposer, NQueensScore, NQueensSolution,
                      NQueensSolutionScoreCalculator, ScoredSolution, StepStats)
from .local_search import LocalSearch
from .scheduling import (ScheduleChains, ScheduleMoveProposer, ScheduleScore,
                         ScheduleSolutionScoreCalculator)


def philox4x32_10(seed: int, chain: int, purpose: int, counter: int):
    import ctypes as C
    out = (C.c_uint32 * 4)()
    load().cs_philox4x32_10(seed, chain, purpose, counter, out)
    return [int(x) for x in out]


def device_count() -> int:
    return int(load().cs_device_count())


ES_MAX_SLOTS = _lib.CS_ES_MAX_SLOTS  # scored slots (days x shifts per day) the scheduling kernels take
MICROBENCH_SMEM_LDS32, MICROBENCH_SMEM_LDS128, MICROBENCH_L2_READ = 0, 1, 2


def microbench(which: int, device: int = -1):
    """Measured on-chip bandwidth (GB/s, rated SM MHz): the roofline denominators (cs_microbench)."""
    import ctypes as C
    gbs, mhz = C.c_double(), C.c_double()
    rc = load().cs_microbench(device, which, C.byref(gbs), C.byref(mhz))
    if rc != 0:
        raise CsError(rc, "cs_microbench", _lib.status_string(rc))
    return float(gbs.value), float(mhz.value)
