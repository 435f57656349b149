"""Host-side mirror of the reference's n-queens plug-in over the C ABI.

Mirrors examples/nqueens/src/lib.rs (NQueensSolution :18-21, NQueensScore :63-71,
NQueensSolutionScoreCalculator :126-140, NQueensInitialSolutionGenerator :152-161,
NQueensMoveProposer :173-256) with the same names and argument meaning, plus the batch
engine `NQueensChains` (thousands of restart chains per GPU) the reference has no
equivalent for.  Every computation happens on the device through libcs_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterator, Optional

import numpy as np

from . import _lib as L

SWAP, CHANGE = L.CS_NQ_SWAP, L.CS_NQ_CHANGE
INT64_MAX = np.iinfo(np.int64).max


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


@dataclass(frozen=True, order=True)
class NQueensScore:
    """examples/nqueens/src/lib.rs:63-71"""
    value: int

    def is_best(self) -> bool:
        return self.value == 0


class NQueensSolution:
    """examples/nqueens/src/lib.rs:18-21: rows[col] = row of the queen in that column."""

    def __init__(self, rows):
        self.rows = np.ascontiguousarray(np.asarray(rows, dtype=np.int64))

    def __eq__(self, other):
        return isinstance(other, NQueensSolution) and np.array_equal(self.rows, other.rows)

    def __lt__(self, other):  # derived Ord: lexicographic over rows
        return tuple(self.rows.tolist()) < tuple(other.rows.tolist())

    def __hash__(self):
        return hash(self.rows.tobytes())

    def __repr__(self):  # board pretty-printer, lib.rs:26-60
        n = len(self.rows)
        bar = "-" * (4 * n + 1)
        lines = [bar]
        for r in range(n):
            lines.append("".join("| Q " if self.rows[c] == r else "|   " for c in range(n)) + "|")
            lines.append(bar)
        return "\n".join(lines)


@dataclass(frozen=True)
class ScoredSolution:
    """local-search/src/local_search.rs:29-47; ordered by (score, solution)."""
    score: NQueensScore
    solution: NQueensSolution

    def __lt__(self, other):
        if self.score != other.score:
            return self.score < other.score
        return self.solution < other.solution


@dataclass
class StepStats:
    moves_scored: int
    steps_accepted: int
    best_score: int
    best_chain: int
    chains_at_best: int
    device_ms: float
    kernel_launches: int


class NQueensChains:
    """Batch engine: `n_chains` independent restart chains of an n-queens board on one GPU.

    chain_offset is the global id of local chain 0 (its Philox stream), so ranks can shard
    chains with no data-path collective.
    """

    def __init__(self, n: int, n_chains: int = 1, *, seed: int = 42, chain_offset: int = 0,
                 neighbourhood: int = SWAP, trace_capacity: int = 0, device: int = -1,
                 force_global: bool = False, force_scalar: bool = False,
                 reference_proposer: bool = False):
        self._lib = L.load()
        self.n, self.n_chains = int(n), int(n_chains)
        self.neighbourhood = neighbourhood
        self.trace_capacity = trace_capacity
        self.chain_offset = chain_offset
        cfg = L.CsNqConfig(n=n, n_chains=n_chains, chain_offset=chain_offset,
                           trace_capacity=trace_capacity, seed=seed, device=device,
                           neighbourhood=neighbourhood,
                           flags=(L.CS_NQ_FLAG_GLOBAL if force_global else 0)
                           | (L.CS_NQ_FLAG_SCALAR if force_scalar else 0)
                           | (L.CS_NQ_FLAG_REFERENCE_PROPOSER if reference_proposer else 0))
        h = C.c_void_p()
        rc = self._lib.cs_nq_create(C.byref(cfg), C.byref(h))
        if rc != L.CS_OK:
            raise L.CsError(rc, "cs_nq_create", L.status_string(rc))
        self._h = h

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.cs_nq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc, where):
        if rc != L.CS_OK:
            raise L.CsError(rc, where, self._lib.cs_nq_last_error(self._h).decode())

    # -- state
    def set_stream(self, cuda_stream: int):
        self._check(self._lib.cs_nq_set_stream(self._h, C.c_void_p(cuda_stream)), "cs_nq_set_stream")

    def init_random(self):
        self._check(self._lib.cs_nq_init_random(self._h), "cs_nq_init_random")

    def set_chains(self, rows, first_chain: int = 0):
        rows = np.ascontiguousarray(np.asarray(rows, dtype=np.int64))
        if rows.ndim == 1:
            rows = rows[None, :]
        if rows.shape[1] != self.n:
            raise ValueError("rows must have shape [count, n]")
        self._check(self._lib.cs_nq_set_chains(self._h, first_chain, rows.shape[0], _ptr(rows)),
                    "cs_nq_set_chains")

    def set_chains_ptr(self, host_ptr: int, count: int, first_chain: int = 0):
        """Same as set_chains from a raw host pointer (e.g. pinned memory) of int64 [count][n]."""
        self._check(self._lib.cs_nq_set_chains(self._h, first_chain, count, C.c_void_p(host_ptr)),
                    "cs_nq_set_chains")

    def set_chains_async_ptr(self, host_ptr: int, count: int, first_chain: int = 0):
        """start the H2D copy of `count` int64 boards at host_ptr (pinned) on the copy stream"""
        self._check(self._lib.cs_nq_set_chains_async(self._h, first_chain, count, C.c_void_p(host_ptr)),
                    "cs_nq_set_chains_async")

    def commit_chains(self):
        """wait for the pending upload, then validate / pack / reset / score like set_chains"""
        self._check(self._lib.cs_nq_commit_chains(self._h), "cs_nq_commit_chains")

    def get_chains(self, first_chain: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.n_chains - first_chain if count is None else count
        out = np.empty((count, self.n), dtype=np.int64)
        self._check(self._lib.cs_nq_get_chains(self._h, first_chain, count, _ptr(out)),
                    "cs_nq_get_chains")
        return out

    def get_best_chains(self, first_chain: int = 0, count: Optional[int] = None):
        count = self.n_chains - first_chain if count is None else count
        out = np.empty((count, self.n), dtype=np.int64)
        sc = np.empty(count, dtype=np.int64)
        self._check(self._lib.cs_nq_get_best_chains(self._h, first_chain, count, _ptr(out), _ptr(sc)),
                    "cs_nq_get_best_chains")
        return out, sc

    def scores(self) -> np.ndarray:
        out = np.empty(self.n_chains, dtype=np.int64)
        self._check(self._lib.cs_nq_get_scores(self._h, _ptr(out)), "cs_nq_get_scores")
        return out

    def status(self) -> np.ndarray:
        out = np.empty(self.n_chains, dtype=np.uint32)
        self._check(self._lib.cs_nq_get_status(self._h, _ptr(out)), "cs_nq_get_status")
        return out

    # -- scoring
    def score_full(self, chain: int = 0) -> int:
        s = C.c_int64()
        self._check(self._lib.cs_nq_score_full(self._h, chain, C.byref(s)), "cs_nq_score_full")
        return int(s.value)

    def eval_moves(self, a, b, chain: int = 0, kind: Optional[int] = None) -> np.ndarray:
        kind = self.neighbourhood if kind is None else kind
        a = np.asarray(a, dtype=np.uint32)
        b = np.asarray(b, dtype=np.uint32)
        mv = np.ascontiguousarray(np.stack([a, b], axis=1)) if len(a) else np.zeros((0, 2), np.uint32)
        out = np.empty(max(len(a), 1), dtype=np.int64)
        self._check(self._lib.cs_nq_eval_moves(self._h, chain, kind, _ptr(mv), len(a), _ptr(out)),
                    "cs_nq_eval_moves")
        return out[: len(a)]

    def enumerate(self, chain: int = 0) -> np.ndarray:
        n = C.c_uint64()
        self._check(self._lib.cs_nq_enumerate(self._h, chain, None, 0, C.byref(n)), "cs_nq_enumerate")
        mv = np.empty((max(n.value, 1), 2), dtype=np.uint32)
        self._check(self._lib.cs_nq_enumerate(self._h, chain, _ptr(mv), n.value, C.byref(n)),
                    "cs_nq_enumerate")
        return mv[: n.value]

    def neighbourhood_deltas(self, chain: int = 0) -> np.ndarray:
        n = C.c_uint64()
        self._check(self._lib.cs_nq_neighbourhood_deltas(self._h, chain, None, 0, C.byref(n)),
                    "cs_nq_neighbourhood_deltas")
        out = np.empty(max(n.value, 1), dtype=np.int64)
        self._check(self._lib.cs_nq_neighbourhood_deltas(self._h, chain, _ptr(out), n.value, C.byref(n)),
                    "cs_nq_neighbourhood_deltas")
        return out[: n.value]

    def band_deltas(self, i_begin: int, i_end: int, chain: int = 0) -> np.ndarray:
        """Every candidate delta of columns [i_begin, i_end) from the production scan (swap)."""
        n = C.c_uint64()
        self._check(self._lib.cs_nq_band_deltas(self._h, chain, i_begin, i_end, None, 0, C.byref(n)),
                    "cs_nq_band_deltas")
        out = np.empty(max(n.value, 1), dtype=np.int64)
        self._check(self._lib.cs_nq_band_deltas(self._h, chain, i_begin, i_end, _ptr(out), n.value, C.byref(n)),
                    "cs_nq_band_deltas")
        return out[: n.value]

    def set_window(self, window_size: int):
        self._check(self._lib.cs_nq_set_window(self._h, window_size), "cs_nq_set_window")

    # -- the hot path
    @staticmethod
    def _stats(s: L.CsStepStats) -> StepStats:
        return StepStats(int(s.moves_scored), int(s.steps_accepted), int(s.best_score),
                         int(s.best_chain), int(s.chains_at_best), float(s.device_ms),
                         int(s.kernel_launches))

    def step(self, n_steps: int = 1) -> StepStats:
        s = L.CsStepStats()
        self._check(self._lib.cs_nq_step(self._h, n_steps, C.byref(s)), "cs_nq_step")
        return self._stats(s)

    def step_enqueue(self, n_steps: int = 1) -> None:
        """Put n_steps chain-steps on the handle's stream and return without waiting."""
        self._check(self._lib.cs_nq_step_enqueue(self._h, n_steps), "cs_nq_step_enqueue")

    def step_wait(self) -> StepStats:
        """Wait for the stream; stats of the last enqueued launch."""
        s = L.CsStepStats()
        self._check(self._lib.cs_nq_step_wait(self._h, C.byref(s)), "cs_nq_step_wait")
        return self._stats(s)

    def exchange_select(self, key_device_ptr: int, elite_device_ptr: int, elite_len: int) -> None:
        """Owner-masked gather of the chain named by a DEVICE key (see cs_nq_exchange_select)."""
        self._check(self._lib.cs_nq_exchange_select(self._h, C.c_void_p(key_device_ptr),
                                                    C.c_void_p(elite_device_ptr), elite_len),
                    "cs_nq_exchange_select")

    def local_search(self, allow_no_improvement_for: int, max_iterations: int) -> StepStats:
        s = L.CsStepStats()
        self._check(self._lib.cs_nq_local_search(self._h, allow_no_improvement_for, max_iterations,
                                                 C.byref(s)), "cs_nq_local_search")
        return self._stats(s)

    def local_search_one(self, start, allow_no_improvement_for: int, max_iterations: int):
        start = np.ascontiguousarray(np.asarray(start, dtype=np.int64))
        if start.shape != (self.n,):
            raise ValueError("start must have n entries")
        best = np.empty(self.n, dtype=np.int64)
        sc = C.c_int64()
        self._check(self._lib.cs_nq_local_search_one(self._h, _ptr(start), allow_no_improvement_for,
                                                     max_iterations, _ptr(best), C.byref(sc)),
                    "cs_nq_local_search_one")
        return best, int(sc.value)

    def trace(self, chain: int = 0):
        n = C.c_uint64()
        cap = max(self.trace_capacity, 1)
        mv = np.empty((cap, 2), dtype=np.uint32)
        sc = np.empty(cap, dtype=np.int64)
        self._check(self._lib.cs_nq_get_trace(self._h, chain, _ptr(mv), _ptr(sc), cap, C.byref(n)),
                    "cs_nq_get_trace")
        k = min(int(n.value), self.trace_capacity)
        return mv[:k], sc[:k], int(n.value)

    def best(self):
        rows = np.empty(self.n, dtype=np.int64)
        sc, ch = C.c_int64(), C.c_uint32()
        self._check(self._lib.cs_nq_best(self._h, _ptr(rows), C.byref(sc), C.byref(ch)), "cs_nq_best")
        return rows, int(sc.value), int(ch.value)


    # -- iterated local search shell (iterated_local_search.rs:96-203)
    def ils_init(self, best_solutions_capacity: int = 32, log_capacity: int = 0):
        self._ils_log_cap = log_capacity
        self._check(self._lib.cs_nq_ils_init(self._h, best_solutions_capacity, log_capacity),
                    "cs_nq_ils_init")

    def ils_run(self, rounds: int, ls_max_iterations: int, allow_no_improvement_for: int,
                stop_when_any_best: bool = False):
        s = L.CsIlsStats()
        self._check(self._lib.cs_nq_ils_run(self._h, rounds, ls_max_iterations, allow_no_improvement_for,
                                            1 if stop_when_any_best else 0, C.byref(s)), "cs_nq_ils_run")
        return dict(moves_scored=int(s.moves_scored), ls_steps=int(s.ls_steps), best_key=int(s.best_key),
                    best_chain=int(s.best_chain), chains_done=int(s.chains_done),
                    rounds_run=int(s.rounds_run), device_ms=float(s.device_ms),
                    kernel_launches=int(s.kernel_launches))

    def ils_log(self, chain: int = 0):
        n = C.c_uint64()
        cap = max(getattr(self, "_ils_log_cap", 0), 1)
        key = np.empty(cap, dtype=np.int64)
        choice = np.empty(cap, dtype=np.uint32)
        self._check(self._lib.cs_nq_ils_get_log(self._h, chain, _ptr(key), _ptr(choice), cap, C.byref(n)),
                    "cs_nq_ils_get_log")
        k = min(int(n.value), getattr(self, "_ils_log_cap", 0))
        return key[:k], choice[:k], int(n.value)

    def ils_best(self, chain: int = 0):
        rows = np.empty(self.n, dtype=np.int64)
        sc = C.c_int64()
        self._check(self._lib.cs_nq_ils_get_best(self._h, chain, _ptr(rows), C.byref(sc)), "cs_nq_ils_get_best")
        return rows, int(sc.value)

    # -- one big instance, neighbourhood partitioned across handles / GPUs
    def set_partition(self, part: int, parts: int):
        self._check(self._lib.cs_nq_set_partition(self._h, part, parts), "cs_nq_set_partition")

    def part_scan(self):
        self._check(self._lib.cs_nq_part_scan(self._h), "cs_nq_part_scan")

    def part_key_device_ptr(self) -> int:
        p = C.c_void_p()
        self._check(self._lib.cs_nq_part_key_device_ptr(self._h, C.byref(p)), "cs_nq_part_key_device_ptr")
        return int(p.value)

    def part_apply(self) -> StepStats:
        s = L.CsStepStats()
        self._check(self._lib.cs_nq_part_apply(self._h, C.byref(s)), "cs_nq_part_apply")
        return self._stats(s)

    def best_key_device_ptr(self) -> int:
        p = C.c_void_p()
        self._check(self._lib.cs_nq_best_key_device_ptr(self._h, C.byref(p)), "cs_nq_best_key_device_ptr")
        return int(p.value)

    def chain_device_ptr(self, chain: int):
        p, stride = C.c_void_p(), C.c_uint32()
        self._check(self._lib.cs_nq_chain_device_ptr(self._h, chain, C.byref(p), C.byref(stride)),
                    "cs_nq_chain_device_ptr")
        return int(p.value), int(stride.value)

    def set_chain_from_device(self, chain: int, dptr: int):
        self._check(self._lib.cs_nq_set_chain_u16_device(self._h, chain, C.c_void_p(dptr)),
                    "cs_nq_set_chain_u16_device")


# ---------------------------------------------------------------- reference-shaped plug-in
class NQueensSolutionScoreCalculator:
    """examples/nqueens/src/lib.rs:122-140 -- full re-score on the device."""

    def __init__(self):
        self._engines = {}

    def _engine(self, n) -> NQueensChains:
        if n not in self._engines:
            self._engines[n] = NQueensChains(n, 1)
        return self._engines[n]

    def get_scored_solution(self, solution: NQueensSolution) -> ScoredSolution:
        e = self._engine(len(solution.rows))
        e.set_chains(solution.rows)
        return ScoredSolution(NQueensScore(e.score_full(0)), solution)


class NQueensInitialSolutionGenerator:
    """examples/nqueens/src/lib.rs:142-161; rng = (seed, chain) Philox stream id."""

    def __init__(self, board_size: int):
        self.board_size = board_size

    def generate_initial_solution(self, rng) -> NQueensSolution:
        seed, chain = rng
        with NQueensChains(self.board_size, 1, seed=seed, chain_offset=chain) as e:
            e.init_random()
            return NQueensSolution(e.get_chains()[0])


class NQueensMoveProposer:
    """examples/nqueens/src/lib.rs:163-256 shape; the neighbourhood is the FULL swap (or
    change) neighbourhood enumerated on the device side of the ABI, not the reference's
    sampled conflicted-column subset."""

    def __init__(self, board_size: int, neighbourhood: int = SWAP):
        self.board_size = board_size
        self.neighbourhood = neighbourhood

    def iter_local_moves(self, start: NQueensSolution, rng=None) -> Iterator[NQueensSolution]:
        with NQueensChains(self.board_size, 1, neighbourhood=self.neighbourhood) as e:
            e.set_chains(start.rows)
            moves = e.enumerate(0)
        for a, b in moves:
            rows = start.rows.copy()
            if self.neighbourhood == SWAP:
                rows[a], rows[b] = rows[b], rows[a]
            else:
                rows[a] = b
            yield NQueensSolution(rows)
