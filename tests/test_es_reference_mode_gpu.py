"""Scheduling reference mode (CS_ES_FLAG_REFERENCE_PROPOSER): the reference's OWN proposer --
ScheduleRandomMoveProposer, examples/employee-scheduling/src/lib.rs:440-491 (random ChangeDay /
SwapDays stream from a CLONED rng) -- its window (.take(window_size), local_search.rs:321) and
its derived-Ord tie-break (score, then date_to_employee; :29-37,323) on the device, against the
oracle's literal restatement (full clones, full re-score, vector compare) over the same Philox
stream."""
import numpy as np
import pytest

import constraint_solver_b200 as cs
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

CASES = [
    # D, E, start weekday, holidays, window, allow, iterations
    (31, 7, 0, [], 100, 20, 60),                                  # the reference driver's rota (main.rs:11-31)
    (31, 7, 0, [(0, 0), (0, 1), (3, 3)], 100, 20, 60),
    (14, 3, 2, [(1, 5)], 40, 5, 30),
    (56, 40, 5, [(e, (7 * e) % 56) for e in range(40)], 100, 20, 25),
    (28, 50, 0, [(e, e % 28) for e in range(0, 50, 3)], 100, 3, 40),
    (9, 2, 6, [], 17, 4, 20),
    (1, 3, 0, [], 10, 3, 5),                                      # D = 1: swaps are skipped
]


def test_local_search_trajectory_equals_oracle():
    for D, E, wd, hol, window, allow, iters in CASES:
        ids = np.arange(E, dtype=np.int64) * 3 + 1                # non-dense employee ids
        hol_ids = [(int(ids[e]), d) for e, d in hol]
        chains = 6
        with cs.ScheduleChains(D, ids, start_weekday=wd, holidays=hol_ids, n_chains=chains, seed=11,
                               chain_offset=2, trace_capacity=iters + 4, reference_proposer=True) as e:
            e.set_window(window)
            e.init_random()
            start = e.get_chains()
            st = e.local_search(allow, iters)
            best, bh, bs = e.get_best_chains()
            cur = e.get_chains()
            scored = 0
            for k in range(chains):
                ref = orc.es_local_search_ref(start[k][:D], ids, 11, 2 + k, wd, hol_ids, allow, iters, window,
                                              trace_cap=iters + 4)
                mv, th, ts, total = e.trace(k)
                assert total == ref["steps"], (D, E, k, total, ref["steps"])
                assert np.array_equal(mv["kind"], ref["trace_kind"]), (D, E, k)
                assert np.array_equal(mv["a"], ref["trace_x"]) and np.array_equal(mv["b"], ref["trace_y"])
                assert np.array_equal(th, ref["trace_hard"]) and np.array_equal(ts, ref["trace_soft"])
                assert (int(bh[k]), int(bs[k])) == (ref["best_hard"], ref["best_soft"])
                assert np.array_equal(best[k][:D], ref["best"]) and np.array_equal(cur[k][:D], ref["current"])
                assert orc.es_score(best[k][:D], wd, hol_ids) == (ref["best_hard"], ref["best_soft"])
                assert best[k][D] == start[k][D]                  # the phantom slot is never moved
                scored += ref["scored"]
            assert st.moves_scored == scored


def test_reference_driver_instance_ils_equals_oracle():
    """get_ils with the reference driver's constants (main.rs:11-31: 7 employees, 31 days, LS 1000
    iterations, window 100, no-improvement 20, best-set 64), ILS rounds replayed by the oracle."""
    D, E, rounds = 31, 7, 30
    ids = np.arange(E, dtype=np.int64)
    chains = 4
    with cs.ScheduleChains(D, ids, n_chains=chains, seed=42, reference_proposer=True) as e:
        e.set_window(100)
        e.init_random()
        e.ils_init(64, log_capacity=rounds)
        st = e.ils_run(rounds, 1000, 20)
        for k in range(chains):
            ref = orc.es_ils_ref(42, k, D, ids, ls_max_iterations=1000, allow_no_improvement_for=20, rounds=rounds,
                                 best_cap=64, window_size=100)
            key, choice, total = e.ils_log(k)
            assert total == ref["rounds"]
            assert np.array_equal(key, ref["round_new_key"]) and np.array_equal(choice, ref["round_choice"])
            rows, hard, soft = e.ils_best(k)
            assert (hard, soft) == (ref["best_hard"], ref["best_soft"]) and np.array_equal(rows, ref["best"])
        assert st["rounds_run"] == rounds
        assert min(e.ils_best(k)[1] for k in range(chains)) == 0   # a feasible rota is found


def test_full_neighbourhood_dump_is_unchanged_by_the_flag_and_one_employee_is_empty():
    D, ids = 20, np.arange(5, dtype=np.int64)
    a = np.ascontiguousarray(np.random.default_rng(3).integers(0, 5, size=D + 1), dtype=np.int64)
    with cs.ScheduleChains(D, ids, reference_proposer=True) as r, cs.ScheduleChains(D, ids) as f:
        r.set_chains(a)
        f.set_chains(a)
        for x, y in zip(r.neighbourhood_deltas(0), f.neighbourhood_deltas(0)):
            assert np.array_equal(x, y)
        with pytest.raises(cs.CsError):
            r.set_window(0)
    with cs.ScheduleChains(6, [9], reference_proposer=True) as one:   # every candidate is tabu
        one.init_random()
        st = one.local_search(5, 10)
        assert st.moves_scored == 0 and int(one.status()[0]) == 3   # CS_CHAIN_EMPTY (the reference spins forever)
