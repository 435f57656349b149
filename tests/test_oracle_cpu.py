"""CPU suite: the oracle against the golden vectors (the reference's own KATs + SURVEY 8c)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc


def _load(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return json.load(f)


def test_philox_known_answers(golden_dir):
    for case in _load(golden_dir, "philox_kat.json"):
        assert orc.philox4x32_10(case["ctr"], case["key"]) == case["out"]


def test_nq_reference_kats(golden_dir):
    g = _load(golden_dir, "nq_kat.json")
    for case in g["reference"] + g["derived"]:
        assert orc.nq_score(case["rows"]) == case["score"]
        if "col_scores" in case:
            assert orc.nq_col_scores(case["rows"]).tolist() == case["col_scores"]


def test_nq_score_equals_line_counter_closed_form():
    # sum over lines k(k-1) == the reference pair loop, incl. non-permutations
    rng = np.random.default_rng(1)
    for n in (1, 2, 3, 7, 16, 33, 100):
        for _ in range(20):
            rows = rng.integers(0, n, size=n)
            cols = np.arange(n)
            tot = 0
            for key in (rows, cols - rows, cols + rows):
                _, cnt = np.unique(key, return_counts=True)
                tot += int((cnt * (cnt - 1)).sum())
            assert tot == orc.nq_score(rows)


def test_nq_selection_kats_reference_tiebreak(golden_dir):
    # best-of-window under the derived Ord (score, solution lexicographic)
    for case in _load(golden_dir, "nq_kat.json")["selection"]:
        rows = np.array(case["rows"], dtype=np.int64)
        n = len(rows)
        cands = []
        for c in case["cols"]:
            for v in range(n):
                if v == rows[c]:
                    continue
                r = rows.copy()
                r[c] = v
                cands.append((orc.nq_score(r), tuple(r.tolist())))
        cands.sort()
        assert list(cands[0][1]) == case["chosen"] and cands[0][0] == case["score"]


def test_nq_init_perm_is_permutation_and_deterministic():
    a = orc.nq_init_perm(42, 3, 1000)
    b = orc.nq_init_perm(42, 3, 1000)
    c = orc.nq_init_perm(42, 4, 1000)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert sorted(a.tolist()) == list(range(1000))


def test_nq_local_search_follows_execute_semantics():
    # local_search.rs:301-342: returns current when is_best; best = last improving neighbour;
    # identity candidates are the only tabu ones; empty neighbourhood breaks.
    res = orc.nq_local_search([1, 3, 0, 2], trace_cap=8)
    assert res["steps"] == 0 and res["best_score"] == 0
    res = orc.nq_local_search([0, 0, 0, 0], kind=orc.SWAP, trace_cap=8)  # all swaps identity
    assert res["steps"] == 0 and res["best_score"] == 12
    rows = orc.nq_init_perm(1, 0, 12)
    res = orc.nq_local_search(rows, allow_no_improvement_for=3, max_iterations=50, trace_cap=64)
    r = rows.copy()
    best_seen, cur = None, orc.nq_score(rows)
    for (i, j, s) in zip(res["trace_a"], res["trace_b"], res["trace_score"]):
        r[i], r[j] = r[j], r[i]
        assert orc.nq_score(r) == s
        if s < cur:
            best_seen = (s, r.copy())
        cur = s
    if best_seen is not None:
        assert res["best_score"] == best_seen[0] and np.array_equal(res["best"], best_seen[1])
    # max_iterations bounds the accepted steps
    res2 = orc.nq_local_search(rows, allow_no_improvement_for=1000, max_iterations=2, trace_cap=8)
    assert res2["steps"] <= 2


def test_nq_window_and_reference_tiebreak_modes():
    rows = orc.nq_init_perm(9, 0, 10)
    full = orc.nq_local_search(rows, kind=orc.CHANGE, tie=orc.TIE_REFERENCE, max_iterations=1,
                               allow_no_improvement_for=5, trace_cap=4)
    win = orc.nq_local_search(rows, kind=orc.CHANGE, tie=orc.TIE_REFERENCE, max_iterations=1,
                              allow_no_improvement_for=5, window_size=7, trace_cap=4)
    assert full["trace_score"][0] <= win["trace_score"][0]
    assert win["trace_a"][0] == 0  # first 7 non-identity candidates all live in column 0


def test_weekday_arithmetic():
    assert orc.weekday(2022, 5, 9) == 0  # Monday, examples/employee-scheduling/src/main.rs:11
    assert orc.weekday(1970, 1, 1) == 3
    assert orc.weekday(2000, 2, 29) == 1
    assert orc.days_from_civil(2022, 6, 8) - orc.days_from_civil(2022, 5, 9) == 30


def test_es_golden_vectors(golden_dir):
    g = _load(golden_dir, "es_kat.json")
    for case in g["cases"]:
        hard, soft = orc.es_score(case["a"], g["start_weekday"], case["holidays"])
        assert (hard, soft) == (case["hard"], case["soft"]), case
        assert orc.es_score_terms(case["a"], g["start_weekday"], case["holidays"]).tolist() == case["terms"]


def test_es_holiday_out_of_range_is_an_error():
    with pytest.raises(ValueError):
        orc.es_score([0, 1, 2], 0, [(0, 3)])


def test_es_phantom_slot_is_not_scored():
    # date_to_employee has D+1 entries (lib.rs:405-412); scoring reads only the first D
    a = [i % 7 for i in range(31)]
    assert orc.es_score(a, 0) == orc.es_score(np.array(a + [3])[:31], 0)


def test_es_local_search_trace_replays():
    rng = np.random.default_rng(5)
    emp = np.arange(5)
    a = rng.integers(0, 5, size=16)
    hol = [(0, 1), (2, 7), (4, 15)]
    res = orc.es_local_search(a, emp, 0, hol, allow_no_improvement_for=4, max_iterations=12,
                              trace_cap=32)
    cur = a.copy()
    for k, x, y, h, s in zip(res["trace_kind"], res["trace_x"], res["trace_y"],
                             res["trace_hard"], res["trace_soft"]):
        if k == orc.ES_CHANGE:
            cur[x] = emp[y]
        else:
            cur[x], cur[y] = cur[y], cur[x]
        assert orc.es_score(cur, 0, hol) == (h, s)
    assert np.array_equal(cur, res["current"])


def test_scheduling_reference_proposer_restatement_against_a_python_rewrite():
    """orc_es_local_search_ref (random ChangeDay / SwapDays stream from a cloned rng, window,
    derived-Ord tie-break; examples/employee-scheduling/src/lib.rs:440-491,
    local_search.rs:315-335) against an independent pure-Python rewrite of the same step."""
    D, ids, wd, hol, window, seed, chain = 12, np.array([2, 5, 9], dtype=np.int64), 3, [(5, 4)], 25, 77, 4
    start = ids[np.random.default_rng(0).integers(0, 3, size=D)]
    ref = orc.es_local_search_ref(start, ids, seed, chain, wd, hol, allow_no_improvement_for=4,
                                  max_iterations=6, window_size=window, trace_cap=8)
    cur = start.copy()
    score = orc.es_score(cur, wd, hol)
    steps, no_improve = 0, 0
    for _ in range(6):
        if score == (0, 0):
            break
        cands, k = [], 0
        while len(cands) < window and k < 1 << 16:
            u = [orc.philox_stream(seed, chain, 2, (3 * k + j) // 4)[(3 * k + j) % 4] for j in range(3)]
            k += 1
            c = cur.copy()
            if (u[0] * 5) >> 32 < 1:
                c[(u[1] * D) >> 32] = ids[(u[2] * len(ids)) >> 32]
            else:
                d1, d2 = (u[1] * D) >> 32, (u[2] * (D - 1)) >> 32
                d2 += d2 >= d1
                c[d1], c[d2] = c[d2], c[d1]
            if not np.array_equal(c, cur):
                cands.append((orc.es_score(c, wd, hol), c.tolist()))
        nb_score, nb = min(cands)                                  # (score, solution) derived Ord
        if nb_score < score:
            no_improve = 0
        else:
            no_improve += 1
            if no_improve >= 4:
                break
        cur, score = np.array(nb, dtype=np.int64), nb_score
        assert (int(ref["trace_hard"][steps]), int(ref["trace_soft"][steps])) == score
        steps += 1
    assert steps == ref["steps"] and np.array_equal(cur, ref["current"])


# ---------------------------------------------------------------- round 2: the fast checker and the second scorer
def test_fast_delta_scorer_equals_clone_and_full_rescore_on_every_candidate():
    """The O(1) counter/delta scorer (the checker used at n = 10 000 ... 10^6) is itself proven
    against the literal reference formulation -- clone + O(n^2) re-score of EVERY candidate --
    on permutations, arbitrary multisets (change moves and the perturbation break the
    permutation, nqueens lib.rs:228,311-312), all-equal and diagonal boards, n up to 400."""
    rng = np.random.default_rng(11)
    for n in (1, 2, 3, 4, 7, 16, 33, 64, 97, 150):
        boards = [rng.permutation(n), rng.integers(0, n, n), np.zeros(n, dtype=np.int64), np.arange(n),
                  np.arange(n)[::-1].copy()]
        if n >= 4:
            r = rng.permutation(n)
            r[1] = r[n - 1]
            boards.append(r)
        for rows in boards:
            for kind in (orc.SWAP, orc.CHANGE):
                slow = orc.nq_neighbourhood_deltas(rows, kind)
                assert np.array_equal(orc.nq_neighbourhood_deltas_mt(rows, kind), slow)
                fast = orc.nq_fast_band_deltas(rows, kind=kind)
                assert np.array_equal(fast, slow), (n, kind)
    for n, make in ((400, lambda: rng.permutation(400)), (301, lambda: rng.integers(0, 301, 301))):
        rows = make()
        slow = orc.nq_neighbourhood_deltas_mt(rows, orc.SWAP)      # 8e4 candidates x 8e4 pair tests
        assert np.array_equal(orc.nq_fast_band_deltas(rows), slow)
        # bands tile the enumeration
        cuts = [0, 1, 17, n // 2, n - 2, n - 1]
        parts = [orc.nq_fast_band_deltas(rows, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
        assert np.array_equal(np.concatenate(parts), slow)
        # argmin by (delta, i, j) == first minimum of the enumeration
        d, a, b, scored = orc.nq_fast_argmin(rows)
        k = int(np.argmin(slow))
        assert slow[k] == d and scored == int((slow != orc.INT64_MAX).sum())
        assert k == a * n - a * (a + 1) // 2 + (b - a - 1)


def test_cpu_delta_baseline_walks_the_oracle_trajectory():
    n, steps = 60, 4
    scored, chk = orc.nq_delta_baseline(5, n, 3, steps, threads=2)
    want_scored, want_sum = 0, 0
    for chain in range(3):
        res = orc.nq_local_search(orc.nq_init_perm(5, chain, n), allow_no_improvement_for=10**9,
                                  max_iterations=steps, trace_cap=steps)
        s0 = orc.nq_score(orc.nq_init_perm(5, chain, n))
        want_sum += int(res["trace_score"][-1]) - s0
        want_scored += steps * n * (n - 1) // 2
    assert (scored, chk) == (want_scored, want_sum)


def test_duplicate_holiday_entries_count_once():
    """The reference keeps holidays in a HashSet<Holiday> (lib.rs:255-259): a repeated (employee,
    day) entry is one holiday.  (the oracle used to count it twice.)"""
    a = [0, 1, 0, 1, 2, 2, 0]
    once = orc.es_score_terms(a, 0, [(0, 0), (1, 3), (2, 4)])
    twice = orc.es_score_terms(a, 0, [(0, 0), (0, 0), (1, 3), (2, 4), (1, 3), (0, 0)])
    assert once.tolist() == twice.tolist() and int(once[0]) == 3
    assert int(orc.es_score_terms(a, 0, [(0, 1), (0, 1)])[0]) == 0


def _second_scorer_case(rng, D, E, wd_shift, holiday_heavy):
    import datetime as dt

    from es_second_scorer import ScheduleSolution, get_scored_solution

    start = dt.date(2022, 5, 9) + dt.timedelta(days=int(wd_shift))   # 2022-05-09 is a Monday
    ids = np.sort(rng.choice(np.arange(0, 3 * E + 5), size=E, replace=False)).astype(np.int64)
    style = rng.integers(0, 4)
    if style == 0:
        a = ids[rng.integers(0, E, size=D + 1)]
    elif style == 1:
        a = ids[rng.integers(0, min(E, 3), size=D + 1)]               # few employees, long runs
    elif style == 2:
        a = ids[(np.arange(D + 1) + rng.integers(0, E)) % E]          # round robin
    else:
        a = ids[(rng.integers(0, E) + np.arange(D + 1) // int(rng.integers(1, 5))) % E]
    nh = int(rng.integers(0, 3 * D * min(E, 4) + 1)) if holiday_heavy else int(rng.integers(0, E + 2))
    hol = [(int(ids[rng.integers(0, E)]), int(rng.integers(0, D))) for _ in range(nh)]  # duplicates allowed
    table = {}
    for emp, day in hol:
        table.setdefault(emp, set()).add(start + dt.timedelta(days=day))
    sol = ScheduleSolution(start, start + dt.timedelta(days=D - 1), a.tolist())
    return get_scored_solution(sol, table), orc.es_score(a[:D], start.weekday(), hol), (D, E, a.tolist(), hol)


def test_c_oracle_against_the_independent_python_scorer_on_random_rotas():
    """>= 1000 random rotas incl. D < 7, D < 14, D >= 64 (beyond the first device limit), single
    employee, holiday-heavy and duplicate-holiday cases: (hard, soft) of the date-based Python
    rewrite (tests/es_second_scorer.py, written from lib.rs:261-375) == the C oracle."""
    rng = np.random.default_rng(20251018)
    n = 0
    shapes = [(1, 1), (2, 1), (3, 2), (6, 2), (7, 3), (8, 1), (9, 4), (13, 2), (14, 5), (15, 3), (21, 7),
              (28, 50), (31, 7), (56, 20), (64, 9), (90, 11), (168, 30)]
    for rep in range(62):
        for D, E in shapes:
            got, want, info = _second_scorer_case(rng, D, E, rng.integers(0, 7), rep % 3 == 0)
            assert got == want, info
            n += 1
    assert n >= 1000


def test_second_scorer_reproduces_the_committed_golden_vectors(golden_dir):
    import datetime as dt

    from es_second_scorer import ScheduleSolution, get_scored_solution

    g = _load(golden_dir, "es_kat.json")
    start = dt.date.fromisoformat(g["start_date"])
    for case in g["cases"]:
        table = {}
        for emp, day in case["holidays"]:
            table.setdefault(emp, set()).add(start + dt.timedelta(days=day))
        sol = ScheduleSolution(start, start + dt.timedelta(days=len(case["a"]) - 1), case["a"])
        assert get_scored_solution(sol, table) == (case["hard"], case["soft"])


# ---------------------------------------------------------------- slot-generalised extension (not pinned by the reference)
def test_extension_reduces_to_the_reference_restatement_at_one_shift_per_day():
    rng = np.random.default_rng(31)
    for rep in range(400):
        D, E, wd = int(rng.integers(1, 193)), int(rng.integers(1, 14)), int(rng.integers(0, 7))
        ids = np.sort(rng.choice(np.arange(0, 60), size=E, replace=False))
        a = ids[rng.integers(0, E if rep % 2 else min(E, 3), size=D)]
        hol = [(int(ids[rng.integers(0, E)]), int(rng.integers(0, D))) for _ in range(int(rng.integers(0, 3 * E)))]
        t8, t10 = orc.es_score_terms(a, wd, hol), orc.esx_score_terms(a, ids, D, 1, wd, hol)
        assert t8.tolist() == t10[:8].tolist() and t10[8] == 0 and t10[9] == 0, (D, E, wd)
        assert orc.esx_score_terms(a, ids, D, 1, wd, hol, [1] * E).tolist() == t10.tolist()   # everybody qualified
    D, E = 20, 5
    ids = np.arange(E) * 2
    a = ids[rng.integers(0, E, size=D)]
    hol = [(2, 3), (8, 19)]
    h1, s1 = orc.es_neighbourhood_deltas(a, ids, 3, hol)
    h2, s2 = orc.esx_neighbourhood_deltas(a, ids, D, 1, 3, hol)
    assert np.array_equal(h1, h2) and np.array_equal(s1, s2)
    r1 = orc.es_local_search(a, ids, 3, hol, allow_no_improvement_for=4, max_iterations=12, trace_cap=16)
    r2 = orc.esx_local_search(a, ids, D, 1, 3, hol, None, allow_no_improvement_for=4, max_iterations=12, trace_cap=16)
    assert r1["steps"] == r2["steps"] and (r1["best_hard"], r1["best_soft"]) == (r2["best_hard"], r2["best_soft"])
    for k in ("best", "current", "trace_kind", "trace_x", "trace_y", "trace_hard", "trace_soft"):
        assert np.array_equal(r1[k], r2[k]), k


def test_extension_known_answers():
    # 2 days x 3 shifts, employee 0 everywhere, on holiday on day 1, not qualified for shift 2
    t = orc.esx_score_terms([0] * 6, [0, 1], 2, 3, 0, [(0, 1)], [0b011, 0b111])
    assert t.tolist() == [3, 5, 0, 0, 0, 0, 0, 0, 6, 2]
    # 14 days x 2 shifts from a Monday, employees alternate by shift: each holds 14 slots in the one
    # 14-day window (> 3) and 7 per 7-day window (> 2, 8 windows x 2 employees)
    a = [0, 1] * 14
    t = orc.esx_score_terms(a, [0, 1], 14, 2, 0)
    assert t[1] == 0 and t[3] == 2 and t[4] == 16 and t[8] == 0 and t[6] == 0
    assert t[2] == 8        # Sat 5 / Sun 6 vs Sat 12 / Sun 13: the same employee per shift kind, 4 pairs x 2 shifts
    assert t[5] == 5 * 2    # every weekday: both employees hold 2 slots => min 2


def test_extension_oracle_against_the_independent_python_rewrite():
    import datetime as dt

    from es_second_scorer import get_scored_solution_slots

    rng = np.random.default_rng(777)
    n = 0
    for rep in range(40):
        for D, S, E in [(1, 2, 2), (3, 3, 2), (7, 2, 3), (9, 3, 4), (14, 3, 5), (16, 2, 9), (28, 3, 50), (40, 3, 6),
                        (64, 3, 12), (96, 2, 5)]:
            start = dt.date(2022, 5, 9) + dt.timedelta(days=int(rng.integers(0, 7)))
            ids = np.sort(rng.choice(np.arange(0, 3 * E + 5), size=E, replace=False)).astype(np.int64)
            T = D * S
            a = ids[rng.integers(0, E if rep % 2 else min(E, 3), size=T)]
            hol = [(int(ids[rng.integers(0, E)]), int(rng.integers(0, D))) for _ in range(int(rng.integers(0, 2 * E + 2)))]
            skills = [int(rng.integers(0, 1 << S)) for _ in range(E)]
            table = {}
            for emp, day in hol:
                table.setdefault(emp, set()).add(start + dt.timedelta(days=day))
            sk = {int(e): {s for s in range(S) if (skills[k] >> s) & 1} for k, e in enumerate(ids)}
            hard, soft, terms = get_scored_solution_slots(start, D, S, a.tolist(), table, sk)
            want = orc.esx_score_terms(a, ids, D, S, start.weekday(), hol, skills)
            assert terms == want.tolist(), (D, S, E, terms, want.tolist())
            assert (hard, soft) == orc.esx_score(a, ids, D, S, start.weekday(), hol, skills)
            n += 1
    assert n == 400
