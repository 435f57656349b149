"""Randomised GPU parity sweeps (fixed seeds): many small random instances of both problems,
every candidate delta of the production scans against the oracle's clone + full re-score, and
LocalSearch::execute trajectories against the oracle's restatement.  Complements the hand-picked
shapes of test_es_gpu.py / test_nq_gpu.py / test_nq_big_gpu.py."""
import numpy as np
import pytest

import constraint_solver_b200 as cs
from oracle import oracle as orc
from test_es_gpu import _oracle_deltas

pytestmark = pytest.mark.gpu


def test_scheduling_random_instances_every_delta_and_trajectory():
    rng = np.random.default_rng(20261018)
    for case in range(48):
        D = int(rng.integers(1, 65))
        E = int(rng.integers(1, 90)) if case % 3 else int(rng.integers(1, 6))   # crowded rotas too
        wd = int(rng.integers(0, 7))
        ids = np.sort(rng.choice(np.arange(0, 4 * E + 5), size=E, replace=False)).astype(np.int64)
        nh = int(rng.integers(0, 3 * E + 1))
        hol = sorted({(int(ids[rng.integers(0, E)]), int(rng.integers(0, D))) for _ in range(nh)})
        # skewed starts: a few employees hold most days (long windows, large spreads) or uniform
        if case % 2:
            p = rng.dirichlet(np.full(E, 0.3))
            start = ids[rng.choice(E, size=D + 1, p=p)]
        else:
            start = ids[rng.integers(0, E, size=D + 1)]
        with cs.ScheduleChains(D, ids, start_weekday=wd, holidays=hol, n_chains=2, trace_capacity=16) as e:
            e.set_chains(np.stack([start, start]))
            ref_h, ref_s = _oracle_deltas(start[:D], ids, wd, hol)
            dev_h, dev_s = e.neighbourhood_deltas(1)
            bad = np.nonzero((dev_h != ref_h) | (dev_s != ref_s))[0]
            assert bad.size == 0, (case, D, E, wd, bad[:5], dev_h[bad[:5]], ref_h[bad[:5]], dev_s[bad[:5]], ref_s[bad[:5]])
            hard, soft = e.scores()
            assert orc.es_score(start[:D], wd, hol) == (int(hard[0]), int(soft[0]))
            ref = orc.es_local_search(start[:D], ids, wd, hol, allow_no_improvement_for=3, max_iterations=8, trace_cap=16)
            st = e.local_search(3, 8)
            mv, th, ts, total = e.trace(0)
            assert total == ref["steps"], (case, D, E)
            assert np.array_equal(mv["kind"], ref["trace_kind"]) and np.array_equal(mv["a"], ref["trace_x"])
            assert np.array_equal(mv["b"], ref["trace_y"])
            assert np.array_equal(th, ref["trace_hard"]) and np.array_equal(ts, ref["trace_soft"])
            ident = int((ref_h == orc.INT64_MAX).sum())
            if ref["steps"] >= 1:   # the first step scored the whole non-identity neighbourhood
                assert st.moves_scored >= 2 * (len(ref_h) - ident)


def test_scheduling_reference_mode_random_instances():
    rng = np.random.default_rng(7)
    for case in range(16):
        D, E, wd = int(rng.integers(2, 65)), int(rng.integers(2, 40)), int(rng.integers(0, 7))
        ids = np.arange(E, dtype=np.int64) * 2 + 3
        hol = sorted({(int(ids[rng.integers(0, E)]), int(rng.integers(0, D))) for _ in range(int(rng.integers(0, 2 * E)))})
        window = int(rng.integers(1, 140))
        with cs.ScheduleChains(D, ids, start_weekday=wd, holidays=hol, n_chains=3, seed=case, trace_capacity=12,
                               reference_proposer=True) as e:
            e.set_window(window)
            e.init_random()
            start = e.get_chains()
            e.local_search(4, 10)
            for k in range(3):
                ref = orc.es_local_search_ref(start[k][:D], ids, case, k, wd, hol, 4, 10, window, trace_cap=12)
                mv, th, ts, total = e.trace(k)
                assert total == ref["steps"], (case, k)
                assert np.array_equal(mv["kind"], ref["trace_kind"]) and np.array_equal(mv["a"], ref["trace_x"])
                assert np.array_equal(mv["b"], ref["trace_y"]) and np.array_equal(th, ref["trace_hard"])
                assert np.array_equal(ts, ref["trace_soft"])


def test_nqueens_random_boards_all_three_scans():
    """shared-memory scalar + packed scans and the global packed scan on random boards"""
    rng = np.random.default_rng(99)
    for case in range(10):
        n = int(rng.integers(256, 700))
        rows = np.ascontiguousarray(rng.permutation(n), dtype=np.int64)
        for _ in range(int(rng.integers(0, 40))):       # a few crowded diagonals
            i = int(rng.integers(0, n - 1))
            j = int(np.where(rows == (rows[i] + 1) % n)[0][0])
            if j != i + 1:
                rows[i + 1], rows[j] = rows[j], rows[i + 1]
        ref = orc.nq_neighbourhood_deltas(rows, orc.SWAP)
        for kw in (dict(), dict(force_scalar=True), dict(force_global=True)):
            with cs.NQueensChains(n, 1, **kw) as e:
                e.set_chains(rows)
                dev = e.neighbourhood_deltas(0)
                bad = np.nonzero(dev != ref)[0]
                assert bad.size == 0, (case, n, kw, bad[:5], dev[bad[:5]], ref[bad[:5]])


def test_scheduling_one_employee_holds_every_day():
    """D = 64 with all days on one employee: total-days bin 64 has no bit in the 64-bit occupancy
    set (present == 1 there); moves out of and back into that state must still be exact."""
    for D, E in [(64, 2), (64, 5), (63, 3), (14, 2)]:
        ids = np.arange(E, dtype=np.int64) + 10
        start = np.full(D + 1, ids[0], dtype=np.int64)
        near = start.copy()
        near[D // 2] = ids[1]                      # one day away from the degenerate state
        for rows in (start, near):
            with cs.ScheduleChains(D, ids, start_weekday=3, holidays=[(int(ids[1]), 0)], trace_capacity=8) as e:
                e.set_chains(rows)
                ref_h, ref_s = _oracle_deltas(rows[:D], ids, 3, [(int(ids[1]), 0)])
                dev_h, dev_s = e.neighbourhood_deltas(0)
                assert np.array_equal(dev_h, ref_h) and np.array_equal(dev_s, ref_s), (D, E)
                ref = orc.es_local_search(rows[:D], ids, 3, [(int(ids[1]), 0)], allow_no_improvement_for=2,
                                          max_iterations=6, trace_cap=8)
                e.local_search(2, 6)
                mv, th, ts, total = e.trace(0)
                assert total == ref["steps"] and np.array_equal(th, ref["trace_hard"]) and np.array_equal(ts, ref["trace_soft"])
