"""Host-side mirror of the reference's employee-scheduling plug-in over the C ABI.

Mirrors examples/employee-scheduling/src/lib.rs: Employee (:119-122), ScheduleSolution
(:127-146, `date_to_employee` with its phantom slot), ScheduleScore (:239-249),
ScheduleSolutionScoreCalculator (:251-375), ScheduleInitialSolutionGenerator (:377-420),
the change/swap move types (:422-491) and get_ils-style construction (:57-117).  Everything
is computed on the device through libcs_b200.so; dates are reduced to (start weekday, D).
"""
from __future__ import annotations

import ctypes as C
import datetime as _dt
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib as L

CHANGE, SWAP = L.CS_ES_CHANGE, L.CS_ES_SWAP
INT64_MAX = np.iinfo(np.int64).max


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


@dataclass(frozen=True, order=True)
class ScheduleScore:
    """lib.rs:239-249: ordered lexicographically hard then soft."""
    hard_score: float
    soft_score: float

    def is_best(self) -> bool:
        return self.hard_score == 0.0 and self.soft_score == 0.0


@dataclass
class EsStepStats:
    moves_scored: int
    steps_accepted: int
    best_hard: int
    best_soft: int
    best_chain: int
    chains_at_best: int
    chains_feasible: int
    device_ms: float
    kernel_launches: int


def weekday_of(date: _dt.date) -> int:
    return date.weekday()  # Monday = 0, like chrono's num_days_from_monday


class ScheduleChains:
    """Batch engine: n_chains independent rota chains on one GPU.

    employees: iterable of int ids.  holidays: iterable of (employee_id, day_index).
    A solution row has n_days + 1 entries (phantom last slot, lib.rs:405-412).
    """

    def __init__(self, n_days: int, employees: Sequence[int], *, start_weekday: int = 0,
                 holidays=(), n_chains: int = 1, seed: int = 42, chain_offset: int = 0,
                 trace_capacity: int = 0, device: int = -1, reference_proposer: bool = False,
                 shifts_per_day: int = 1, skills=None):
        """shifts_per_day > 1 / skills: the slot-generalised EXTENSION (not pinned by the reference):
        n_days * shifts_per_day scored slots (slot t = day t // S, shift t % S); skills[k] is a bit
        mask over shift kinds for employees[k] (None = everybody qualified for everything)."""
        self._lib = L.load()
        self.n_days = int(n_days)
        self.shifts_per_day = int(shifts_per_day)
        self.n_scored = self.n_days * self.shifts_per_day       # scored slots (move indices)
        self.n_slots = self.n_scored + 1                        # + the phantom slot
        self.employees = np.ascontiguousarray(sorted(int(e) for e in employees), dtype=np.int64)
        self.n_employees = len(self.employees)
        self.n_chains = int(n_chains)
        self.trace_capacity = trace_capacity
        self.chain_offset = chain_offset
        hol = np.asarray(list(holidays), dtype=np.int64).reshape(-1, 2)
        he = np.ascontiguousarray(hol[:, 0]) if len(hol) else np.zeros(1, np.int64)
        hd = np.ascontiguousarray(hol[:, 1]) if len(hol) else np.zeros(1, np.int64)
        cfg = L.CsEsConfig(n_days=n_days, n_employees=self.n_employees, start_weekday=start_weekday,
                           n_chains=n_chains, chain_offset=chain_offset,
                           trace_capacity=trace_capacity, seed=seed, device=device,
                           flags=L.CS_ES_FLAG_REFERENCE_PROPOSER if reference_proposer else 0)
        ids = np.ascontiguousarray(np.asarray(list(employees), dtype=np.int64))
        h = C.c_void_p()
        if self.shifts_per_day == 1 and skills is None:
            rc = self._lib.cs_es_create(C.byref(cfg), _ptr(ids), _ptr(he), _ptr(hd), len(hol), C.byref(h))
        else:
            sk = None
            if skills is not None:
                sk = np.ascontiguousarray(np.asarray(list(skills), dtype=np.uint32))
                if len(sk) != len(ids):
                    raise ValueError("skills must have one entry per employee")
            rc = self._lib.cs_es_create_ex(C.byref(cfg), _ptr(ids), _ptr(he), _ptr(hd), len(hol),
                                           self.shifts_per_day, _ptr(sk) if sk is not None else None, C.byref(h))
        if rc != L.CS_OK:
            raise L.CsError(rc, "cs_es_create", L.status_string(rc))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cs_es_destroy(self._h)
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
            raise L.CsError(rc, where, self._lib.cs_es_last_error(self._h).decode())

    def set_stream(self, cuda_stream: int):
        self._check(self._lib.cs_es_set_stream(self._h, C.c_void_p(cuda_stream)), "cs_es_set_stream")

    def set_window(self, window_size: int):
        """window_size of LocalSearch::new; only the reference proposer truncates (main.rs:26: 100)"""
        self._check(self._lib.cs_es_set_window(self._h, window_size), "cs_es_set_window")

    def init_random(self):
        self._check(self._lib.cs_es_init_random(self._h), "cs_es_init_random")

    def _rows(self, rows):
        rows = np.ascontiguousarray(np.asarray(rows, dtype=np.int64))
        if rows.ndim == 1:
            rows = rows[None, :]
        if rows.shape[1] == self.n_scored:  # no phantom supplied: repeat the last slot
            rows = np.ascontiguousarray(np.concatenate([rows, rows[:, -1:]], axis=1))
        if rows.shape[1] != self.n_slots:
            raise ValueError("rows must have n_days * shifts_per_day + 1 (or without the + 1) entries")
        return rows

    def set_chains(self, rows, first_chain: int = 0):
        rows = self._rows(rows)
        self._check(self._lib.cs_es_set_chains(self._h, first_chain, rows.shape[0], _ptr(rows)),
                    "cs_es_set_chains")

    def set_chains_ptr(self, host_ptr: int, count: int, first_chain: int = 0):
        self._check(self._lib.cs_es_set_chains(self._h, first_chain, count, C.c_void_p(host_ptr)),
                    "cs_es_set_chains")

    def set_chains_async_ptr(self, host_ptr: int, count: int, first_chain: int = 0):
        """start the H2D copy of `count` int64 rotas at host_ptr (pinned) on the copy stream"""
        self._check(self._lib.cs_es_set_chains_async(self._h, first_chain, count, C.c_void_p(host_ptr)),
                    "cs_es_set_chains_async")

    def commit_chains(self):
        """wait for the pending upload, then convert / reset / score like set_chains"""
        self._check(self._lib.cs_es_commit_chains(self._h), "cs_es_commit_chains")

    def get_chains(self, first_chain: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.n_chains - first_chain if count is None else count
        out = np.empty((count, self.n_slots), dtype=np.int64)
        self._check(self._lib.cs_es_get_chains(self._h, first_chain, count, _ptr(out)), "cs_es_get_chains")
        return out

    def get_best_chains(self, first_chain: int = 0, count: Optional[int] = None):
        count = self.n_chains - first_chain if count is None else count
        out = np.empty((count, self.n_slots), dtype=np.int64)
        bh = np.empty(count, dtype=np.int64)
        bs = np.empty(count, dtype=np.int64)
        self._check(self._lib.cs_es_get_best_chains(self._h, first_chain, count, _ptr(out), _ptr(bh), _ptr(bs)),
                    "cs_es_get_best_chains")
        return out, bh, bs

    def scores(self):
        hard = np.empty(self.n_chains, dtype=np.int64)
        soft = np.empty(self.n_chains, dtype=np.int64)
        self._check(self._lib.cs_es_get_scores(self._h, _ptr(hard), _ptr(soft)), "cs_es_get_scores")
        return hard, soft

    def status(self):
        out = np.empty(self.n_chains, dtype=np.uint32)
        self._check(self._lib.cs_es_get_status(self._h, _ptr(out)), "cs_es_get_status")
        return out

    def score_full(self, chain: int = 0):
        hard, soft = C.c_int64(), C.c_int64()
        terms = (C.c_int64 * 8)()
        self._check(self._lib.cs_es_score_full(self._h, chain, C.byref(hard), C.byref(soft), terms),
                    "cs_es_score_full")
        return int(hard.value), int(soft.value), [int(x) for x in terms]

    def score_full_ex(self, chain: int = 0):
        """(hard, soft, [H1..H4, S1..S4, X1 same-day overlap, X2 skill]) by the reference-loop kernel"""
        hard, soft = C.c_int64(), C.c_int64()
        terms = (C.c_int64 * 10)()
        self._check(self._lib.cs_es_score_full_ex(self._h, chain, C.byref(hard), C.byref(soft), terms),
                    "cs_es_score_full_ex")
        return int(hard.value), int(soft.value), [int(x) for x in terms]

    @staticmethod
    def _moves(kind, a, b):
        kind = np.broadcast_to(np.asarray(kind, dtype=np.uint32), np.shape(a))
        mv = np.zeros(len(a), dtype=[("kind", np.uint32), ("a", np.uint32), ("b", np.uint32)])
        mv["kind"], mv["a"], mv["b"] = kind, a, b
        return mv

    def eval_moves(self, kind, a, b, chain: int = 0):
        mv = self._moves(kind, np.asarray(a), np.asarray(b))
        dh = np.empty(max(len(mv), 1), dtype=np.int64)
        ds = np.empty(max(len(mv), 1), dtype=np.int64)
        self._check(self._lib.cs_es_eval_moves(self._h, chain, _ptr(mv), len(mv), _ptr(dh), _ptr(ds)),
                    "cs_es_eval_moves")
        return dh[: len(mv)], ds[: len(mv)]

    def enumerate(self, chain: int = 0):
        n = C.c_uint64()
        self._check(self._lib.cs_es_enumerate(self._h, chain, None, 0, C.byref(n)), "cs_es_enumerate")
        mv = np.zeros(max(n.value, 1), dtype=[("kind", np.uint32), ("a", np.uint32), ("b", np.uint32)])
        self._check(self._lib.cs_es_enumerate(self._h, chain, _ptr(mv), n.value, C.byref(n)), "cs_es_enumerate")
        return mv[: n.value]

    def neighbourhood_deltas(self, chain: int = 0):
        n = C.c_uint64()
        self._check(self._lib.cs_es_neighbourhood_deltas(self._h, chain, None, None, 0, C.byref(n)),
                    "cs_es_neighbourhood_deltas")
        dh = np.empty(max(n.value, 1), dtype=np.int64)
        ds = np.empty(max(n.value, 1), dtype=np.int64)
        self._check(self._lib.cs_es_neighbourhood_deltas(self._h, chain, _ptr(dh), _ptr(ds), n.value, C.byref(n)),
                    "cs_es_neighbourhood_deltas")
        return dh[: n.value], ds[: n.value]

    @staticmethod
    def _stats(s):
        return EsStepStats(int(s.moves_scored), int(s.steps_accepted), int(s.best_hard), int(s.best_soft),
                           int(s.best_chain), int(s.chains_at_best), int(s.chains_feasible),
                           float(s.device_ms), int(s.kernel_launches))

    def step(self, n_steps: int = 1) -> EsStepStats:
        s = L.CsEsStepStats()
        self._check(self._lib.cs_es_step(self._h, n_steps, C.byref(s)), "cs_es_step")
        return self._stats(s)

    def step_enqueue(self, n_steps: int = 1) -> None:
        """Put n_steps chain-steps on the handle's stream and return without waiting."""
        self._check(self._lib.cs_es_step_enqueue(self._h, n_steps), "cs_es_step_enqueue")

    def step_wait(self) -> EsStepStats:
        """Wait for the stream; stats of the last enqueued launch."""
        s = L.CsEsStepStats()
        self._check(self._lib.cs_es_step_wait(self._h, C.byref(s)), "cs_es_step_wait")
        return self._stats(s)

    def exchange_select(self, key_device_ptr: int, elite_device_ptr: int, elite_len: int) -> None:
        """Owner-masked gather of the chain named by a DEVICE key (see cs_es_exchange_select)."""
        self._check(self._lib.cs_es_exchange_select(self._h, C.c_void_p(key_device_ptr),
                                                    C.c_void_p(elite_device_ptr), elite_len),
                    "cs_es_exchange_select")

    def local_search(self, allow_no_improvement_for: int, max_iterations: int) -> EsStepStats:
        s = L.CsEsStepStats()
        self._check(self._lib.cs_es_local_search(self._h, allow_no_improvement_for, max_iterations, C.byref(s)),
                    "cs_es_local_search")
        return self._stats(s)

    def local_search_one(self, start, allow_no_improvement_for: int, max_iterations: int):
        start = self._rows(start)[0]
        best = np.empty(self.n_slots, dtype=np.int64)
        bh, bs = C.c_int64(), C.c_int64()
        self._check(self._lib.cs_es_local_search_one(self._h, _ptr(start), allow_no_improvement_for,
                                                     max_iterations, _ptr(best), C.byref(bh), C.byref(bs)),
                    "cs_es_local_search_one")
        return best, int(bh.value), int(bs.value)

    def trace(self, chain: int = 0):
        n = C.c_uint64()
        cap = max(self.trace_capacity, 1)
        mv = np.zeros(cap, dtype=[("kind", np.uint32), ("a", np.uint32), ("b", np.uint32)])
        hard = np.empty(cap, dtype=np.int64)
        soft = np.empty(cap, dtype=np.int64)
        self._check(self._lib.cs_es_get_trace(self._h, chain, _ptr(mv), _ptr(hard), _ptr(soft), cap, C.byref(n)),
                    "cs_es_get_trace")
        k = min(int(n.value), self.trace_capacity)
        return mv[:k], hard[:k], soft[:k], int(n.value)

    def best(self):
        rows = np.empty(self.n_slots, dtype=np.int64)
        h, s, c = C.c_int64(), C.c_int64(), C.c_uint32()
        self._check(self._lib.cs_es_best(self._h, _ptr(rows), C.byref(h), C.byref(s), C.byref(c)), "cs_es_best")
        return rows, int(h.value), int(s.value), int(c.value)


    # -- iterated local search shell (iterated_local_search.rs:96-203)
    def ils_init(self, best_solutions_capacity: int = 32, log_capacity: int = 0):
        self._ils_log_cap = log_capacity
        self._check(self._lib.cs_es_ils_init(self._h, best_solutions_capacity, log_capacity),
                    "cs_es_ils_init")

    def ils_run(self, rounds: int, ls_max_iterations: int, allow_no_improvement_for: int,
                stop_when_any_best: bool = False):
        s = L.CsIlsStats()
        self._check(self._lib.cs_es_ils_run(self._h, rounds, ls_max_iterations, allow_no_improvement_for,
                                            1 if stop_when_any_best else 0, C.byref(s)), "cs_es_ils_run")
        return dict(moves_scored=int(s.moves_scored), ls_steps=int(s.ls_steps), best_key=int(s.best_key),
                    best_chain=int(s.best_chain), chains_done=int(s.chains_done),
                    rounds_run=int(s.rounds_run), device_ms=float(s.device_ms),
                    kernel_launches=int(s.kernel_launches))

    def ils_log(self, chain: int = 0):
        n = C.c_uint64()
        cap = max(getattr(self, "_ils_log_cap", 0), 1)
        key = np.empty(cap, dtype=np.int64)
        choice = np.empty(cap, dtype=np.uint32)
        self._check(self._lib.cs_es_ils_get_log(self._h, chain, _ptr(key), _ptr(choice), cap, C.byref(n)),
                    "cs_es_ils_get_log")
        k = min(int(n.value), getattr(self, "_ils_log_cap", 0))
        return key[:k], choice[:k], int(n.value)

    def ils_best(self, chain: int = 0):
        rows = np.empty(self.n_slots, dtype=np.int64)
        h, s = C.c_int64(), C.c_int64()
        self._check(self._lib.cs_es_ils_get_best(self._h, chain, _ptr(rows), C.byref(h), C.byref(s)),
                    "cs_es_ils_get_best")
        return rows, int(h.value), int(s.value)

    def best_key_device_ptr(self) -> int:
        p = C.c_void_p()
        self._check(self._lib.cs_es_best_key_device_ptr(self._h, C.byref(p)), "cs_es_best_key_device_ptr")
        return int(p.value)

    def chain_device_ptr(self, chain: int):
        p, n = C.c_void_p(), C.c_uint32()
        self._check(self._lib.cs_es_chain_device_ptr(self._h, chain, C.byref(p), C.byref(n)),
                    "cs_es_chain_device_ptr")
        return int(p.value), int(n.value)


class ScheduleSolutionScoreCalculator:
    """lib.rs:251-375 shape: new(employee_to_holidays) + get_scored_solution(solution)."""

    def __init__(self, n_days, employees, start_weekday=0, holidays=()):
        self._e = ScheduleChains(n_days, employees, start_weekday=start_weekday, holidays=holidays)

    def get_scored_solution(self, date_to_employee):
        self._e.set_chains(date_to_employee)
        hard, soft, _ = self._e.score_full(0)
        return ScheduleScore(float(hard), float(soft)), np.asarray(date_to_employee)


class ScheduleMoveProposer:
    """lib.rs:493-559 shape (the exhaustive proposer) / :440-491 (`reference=True`: the random
    ChangeDay / SwapDays proposer get_ils installs): carries what the device needs to build the
    handle; the neighbourhood itself is enumerated and scored on the GPU."""

    def __init__(self, n_days, employees, start_weekday=0, holidays=(), reference=False):
        self.n_days, self.employees = int(n_days), list(employees)
        self.start_weekday, self.holidays, self.reference = start_weekday, list(holidays), reference
